// K2 — CSR x dense propagation over the node graph, batched over time.
//
// Replaces `x = adj @ x` of lib/sgp_preprocessing.py:200-203 (torch_sparse spmm_sum) and the
// torch.cat of lib/nn/encoders/sgp_spatial_encoder.py:35: every hop reads one feature block of
// the concatenated [Tc, N, D] buffer and writes the next block in place of the `res` list.
//
// Two operator formats:
//   * plain CSR (any graph, any F): one warp per (t, row); the row's (col, val) slice is staged
//     through a warp-private shared-memory window of 32 edges, every lane then gathers its
//     float4 slice of the source row (coalesced 512 B per warp-load) and accumulates with packed
//     FFMA2.  Each gathered row is used for ONE output row -> L1/L2-bandwidth bound.
//   * RBU (row-block-union, F % 128 == 0): R locality-grouped rows share the union of their
//     columns; each source row slice is loaded once per group and reused from registers for all
//     R output rows (dense [U, R] value slab, zeros where a row lacks the column).  Cuts the
//     L2->SM gather traffic by ~R*deg/U and is the path used on kNN sensor graphs.
// Bound: HBM nominally (bytes/hop = 8 nnz + 4(N+1) + 8 N F Tc), but at deg 100 / F 256 the fp32
// FMA pipe (2 nnz F flops) and the L2->SM gather traffic are the tighter limits; see DESIGN.md.
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"

namespace sgp {

constexpr int kSpmmWarps = 8;

// ---------------------------------------------------------------------------------------------
// CSR, vectorised: lane owns NV float4 (columns f0 + q*128 + 4*lane .. +3)
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(kSpmmWarps * 32)
spmm_csr_vec(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
             const float* __restrict__ val, const int32_t* __restrict__ row_order,
             const float* __restrict__ src, int64_t s_ts, int64_t s_ns,
             const float* __restrict__ src2, int64_t s2_ts, int64_t s2_ns, int n_split,
             float* __restrict__ dst, int64_t d_ts, int64_t d_ns,
             int n_rows, int F, long long total /* Tc * n_rows */) {
    __shared__ int2 stage[kSpmmWarps][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * kSpmmWarps + warp;
    if (gw >= total) return;
    const int t = (int)(gw / n_rows), r = (int)(gw % n_rows);
    const int i = row_order ? row_order[r] : r;
    const int f0 = blockIdx.y * (NV * 128) + lane * 4;
    const float* sp = src + (size_t)t * s_ts + f0;
    const float* sp2 = src2 ? src2 + (size_t)t * s2_ts + f0 : nullptr;
    // source row c lives in `src` when c < n_split, else in the halo buffer `src2`
    auto row_ptr = [&](int c) -> const float* {
        return (c < n_split) ? sp + (size_t)c * s_ns : sp2 + (size_t)(c - n_split) * s2_ns;
    };
    bool on[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) on[q] = (f0 + q * 128) < F;

    float2 acc[NV][2];
#pragma unroll
    for (int q = 0; q < NV; ++q) acc[q][0] = acc[q][1] = make_float2(0.f, 0.f);

    const int beg = rowptr[i], end = rowptr[i + 1];
    int2* win = stage[warp];
    for (int base = beg; base < end; base += 32) {
        const int e = base + lane;
        int2 cv = make_int2(0, 0);
        if (e < end) cv = make_int2(__ldg(col + e), __float_as_int(__ldg(val + e)));
        __syncwarp();
        win[lane] = cv;
        __syncwarp();
        const int cnt = min(32, end - base);
        int j = 0;
        for (; j + 4 <= cnt; j += 4) {
            int2 c[4];
            float4 xv[4][NV];
#pragma unroll
            for (int u = 0; u < 4; ++u) c[u] = win[j + u];
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int q = 0; q < NV; ++q)
                    if (on[q]) xv[u][q] = ldg_f4(row_ptr(c[u].x) + q * 128);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int q = 0; q < NV; ++q)
                    if (on[q]) fma4(acc[q][0], acc[q][1], __int_as_float(c[u].y), xv[u][q]);
        }
        for (; j < cnt; ++j) {
            const int2 c = win[j];
#pragma unroll
            for (int q = 0; q < NV; ++q)
                if (on[q]) {
                    const float4 xv = ldg_f4(row_ptr(c.x) + q * 128);
                    fma4(acc[q][0], acc[q][1], __int_as_float(c.y), xv);
                }
        }
    }
    float* dp = dst + (size_t)t * d_ts + (size_t)i * d_ns + f0;
#pragma unroll
    for (int q = 0; q < NV; ++q)
        if (on[q]) st_f4(dp + q * 128, make_float4(acc[q][0].x, acc[q][0].y, acc[q][1].x, acc[q][1].y));
}

// CSR, scalar fallback (F not a multiple of 4 or unaligned views): lane owns columns lane+32q
__global__ void __launch_bounds__(kSpmmWarps * 32)
spmm_csr_scalar(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                const float* __restrict__ val, const int32_t* __restrict__ row_order,
                const float* __restrict__ src, int64_t s_ts, int64_t s_ns,
                const float* __restrict__ src2, int64_t s2_ts, int64_t s2_ns, int n_split,
                float* __restrict__ dst, int64_t d_ts, int64_t d_ns,
                int n_rows, int F, long long total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long gw = (long long)blockIdx.x * kSpmmWarps + warp;
    if (gw >= total) return;
    const int t = (int)(gw / n_rows), r = (int)(gw % n_rows);
    const int i = row_order ? row_order[r] : r;
    const int beg = rowptr[i], end = rowptr[i + 1];
    for (int f = blockIdx.y * 128 + lane; f < min(F, (int)(blockIdx.y + 1) * 128); f += 32) {
        float acc = 0.f;
        for (int e = beg; e < end; ++e) {
            const int c = __ldg(col + e);
            const float* p = (c < n_split) ? src + (size_t)t * s_ts + (size_t)c * s_ns
                                           : src2 + (size_t)t * s2_ts + (size_t)(c - n_split) * s2_ns;
            acc = fmaf(__ldg(val + e), __ldg(p + f), acc);
        }
        dst[(size_t)t * d_ts + (size_t)i * d_ns + f] = acc;
    }
}


// ---------------------------------------------------------------------------------------------
// RBU v2: CTA = one group x TSPAN time steps.  The group's slab (union columns + dense [U, R]
// values, a contiguous range of the operator arrays) is staged ONCE per CTA in shared memory
// with cp.async; warp w owns feature chunk (w % nfc) and walks time steps (w / nfc), +tpb, ...
// Per union column a lane then issues only ONE long-latency load (its float4 of the gathered
// source row); column ids and values come from shared memory as warp-broadcast LDS.  Gathers
// are software-pipelined four columns ahead (explicit A/B register buffers) so that ~8 x 512 B
// are in flight per warp while the previous quad's 128 FFMA2 issue.
// ---------------------------------------------------------------------------------------------
constexpr int kRbuPiece = 256;   // union columns staged per pass (R=16: 16 KB values + 1 KB ids)

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem_src));
}

// acc[r] += a[r] * x for the R rows of one union column; the column's R values are already in
// registers (`cur`), the NEXT column's values are fetched from shared memory into `nxt` first so
// that the LDS latency hides behind this column's 2R FFMA2.
template <int R>
__device__ __forceinline__ void rbu_fma_col(float2 (&acc)[R][2], float4 (&cur)[R / 4],
                                            const float* __restrict__ next_av, const float4& x) {
    float4 nxt[R / 4];
#pragma unroll
    for (int k = 0; k < R / 4; ++k) nxt[k] = *reinterpret_cast<const float4*>(next_av + 4 * k);
#pragma unroll
    for (int k = 0; k < R / 4; ++k) {
        fma4(acc[4 * k + 0][0], acc[4 * k + 0][1], cur[k].x, x);
        fma4(acc[4 * k + 1][0], acc[4 * k + 1][1], cur[k].y, x);
        fma4(acc[4 * k + 2][0], acc[4 * k + 2][1], cur[k].z, x);
        fma4(acc[4 * k + 3][0], acc[4 * k + 3][1], cur[k].w, x);
    }
#pragma unroll
    for (int k = 0; k < R / 4; ++k) cur[k] = nxt[k];
}

template <int R, int MINB>
__global__ void __launch_bounds__(128, MINB)
spmm_rbu_v2(const int32_t* __restrict__ grp_ptr, const int32_t* __restrict__ grp_rows,
            const int32_t* __restrict__ ucol, const float* __restrict__ uval,
            int nfc, int tpb, int tspan,
            const float* __restrict__ src, int64_t s_ts, int64_t s_ns,
            const float* __restrict__ src2, int64_t s2_ts, int64_t s2_ns, int n_split,
            float* __restrict__ dst, int64_t d_ts, int64_t d_ns, int Tc) {
    __shared__ __align__(16) float s_val[(kRbuPiece + 1) * R];   // +1 row: the look-ahead read
    __shared__ __align__(16) int s_col[kRbuPiece];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthr = blockDim.x;
    const int g = blockIdx.x;
    const int fc = warp % nfc, tl = warp / nfc;
    const int t_begin = blockIdx.y * tspan, t_end = min(Tc, t_begin + tspan);
    const int beg = grp_ptr[g], end = grp_ptr[g + 1];
    const int foff = fc * 128 + lane * 4;

    for (int t0 = t_begin; t0 < t_end; t0 += tpb) {
        const int t = t0 + tl;
        const bool live = t < t_end;
        const float* sp = src + (size_t)(live ? t : t_begin) * s_ts + foff;
        const float* sp2 = src2 ? src2 + (size_t)(live ? t : t_begin) * s2_ts + foff : nullptr;
        auto row_ptr = [&](int c) -> const float* {   // halo rows (c >= n_split) live in src2
            return (c < n_split) ? sp + (size_t)c * s_ns : sp2 + (size_t)(c - n_split) * s2_ns;
        };
        float2 acc[R][2];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r][0] = acc[r][1] = make_float2(0.f, 0.f);

        for (int p0 = beg; p0 < end; p0 += kRbuPiece) {
            const int cnt = min(kRbuPiece, end - p0);
            // a single piece (the common case) stays resident for all time steps of the CTA
            if (t0 == t_begin || end - beg > kRbuPiece) {
                __syncthreads();
                for (int i = tid; i < cnt * (R / 4); i += nthr)
                    cp_async16(s_val + i * 4, uval + (size_t)p0 * R + i * 4);
                for (int i = tid; i < cnt; i += nthr) cp_async4(s_col + i, ucol + p0 + i);
                cp_async_commit();
                cp_async_wait<0>();
                __syncthreads();
            }
            if (!live) continue;
            const int nq = cnt >> 2;
            float4 a_cur[R / 4];
#pragma unroll
            for (int k = 0; k < R / 4; ++k) a_cur[k] = *reinterpret_cast<const float4*>(s_val + 4 * k);
            float4 xa[4], xb[4];
            auto gather = [&](float4 (&x)[4], int q) {
                const int4 c = *reinterpret_cast<const int4*>(s_col + 4 * q);
                x[0] = ldg_f4_stream(row_ptr(c.x));
                x[1] = ldg_f4_stream(row_ptr(c.y));
                x[2] = ldg_f4_stream(row_ptr(c.z));
                x[3] = ldg_f4_stream(row_ptr(c.w));
            };
            auto compute = [&](const float4 (&x)[4], int q) {
#pragma unroll
                for (int j = 0; j < 4; ++j) rbu_fma_col<R>(acc, a_cur, s_val + (size_t)(4 * q + j + 1) * R, x[j]);
            };
            int q = 0;
            if (nq > 0) gather(xa, 0);
            for (; q + 2 <= nq; q += 2) {
                gather(xb, q + 1);
                compute(xa, q);
                if (q + 2 < nq) gather(xa, q + 2);
                compute(xb, q + 1);
            }
            if (q < nq) compute(xa, q);
            for (int j = nq * 4; j < cnt; ++j) {
                const float4 x = ldg_f4_stream(row_ptr(s_col[j]));
                rbu_fma_col<R>(acc, a_cur, s_val + (size_t)(j + 1) * R, x);
            }
        }
        if (live) {
            float* dp = dst + (size_t)t * d_ts + foff;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int row = __ldg(grp_rows + (size_t)g * R + r);
                if (row >= 0)
                    st_f4(dp + (size_t)row * d_ns,
                          make_float4(acc[r][0].x, acc[r][0].y, acc[r][1].x, acc[r][1].y));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// RBU v3: same decomposition as v2, but the gathered source rows no longer pass through
// registers: every lane cp.async's its 16-byte slice of 4 source rows per stage straight from L2
// into a warp-private shared-memory ring (LDGSTS, no register cost), kRing stages deep, so that
// ~12 gathers per lane are in flight while the FFMA2 stream of the current quad issues.  A lane
// only ever touches its own 16-byte column of the ring, so the ring needs no warp barrier —
// cp.async.wait_group in program order is enough.
// ---------------------------------------------------------------------------------------------
constexpr int kRing = 4;

template <int R>
__global__ void __launch_bounds__(128, 4)
spmm_rbu_v3(const int32_t* __restrict__ grp_ptr, const int32_t* __restrict__ grp_rows,
            const int32_t* __restrict__ ucol, const float* __restrict__ uval,
            int nfc, int tpb, int tspan,
            const float* __restrict__ src, int64_t s_ts, int64_t s_ns,
            const float* __restrict__ src2, int64_t s2_ts, int64_t s2_ns, int n_split,
            float* __restrict__ dst, int64_t d_ts, int64_t d_ns, int Tc) {
    extern __shared__ __align__(16) float dyn[];
    float* s_val = dyn;                                        // [(kRbuPiece + 1) * R]
    int* s_col = reinterpret_cast<int*>(s_val + (kRbuPiece + 1) * R);   // [kRbuPiece]
    float* ring = reinterpret_cast<float*>(s_col + kRbuPiece);          // [warps][kRing][4][128]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthr = blockDim.x;
    const int g = blockIdx.x;
    const int fc = warp % nfc, tl = warp / nfc;
    const int t_begin = blockIdx.y * tspan, t_end = min(Tc, t_begin + tspan);
    const int beg = grp_ptr[g], end = grp_ptr[g + 1];
    const int foff = fc * 128 + lane * 4;
    float* myring = ring + (size_t)warp * kRing * 512 + lane * 4;

    for (int t0 = t_begin; t0 < t_end; t0 += tpb) {
        const int t = t0 + tl;
        const bool live = t < t_end;
        const float* sp = src + (size_t)(live ? t : t_begin) * s_ts + foff;
        const float* sp2 = src2 ? src2 + (size_t)(live ? t : t_begin) * s2_ts + foff : nullptr;
        auto row_ptr = [&](int c) -> const float* {   // halo rows (c >= n_split) live in src2
            return (c < n_split) ? sp + (size_t)c * s_ns : sp2 + (size_t)(c - n_split) * s2_ns;
        };
        float2 acc[R][2];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r][0] = acc[r][1] = make_float2(0.f, 0.f);

        for (int p0 = beg; p0 < end; p0 += kRbuPiece) {
            const int cnt = min(kRbuPiece, end - p0);
            if (t0 == t_begin || end - beg > kRbuPiece) {
                cp_async_wait<0>();
                __syncthreads();
                for (int i = tid; i < cnt * (R / 4); i += nthr)
                    cp_async16(s_val + i * 4, uval + (size_t)p0 * R + i * 4);
                for (int i = tid; i < cnt; i += nthr) cp_async4(s_col + i, ucol + p0 + i);
                cp_async_commit();
                cp_async_wait<0>();
                __syncthreads();
            }
            if (!live) continue;
            const int nq = (cnt + 3) >> 2;          // the last quad may be partial
            auto gather = [&](int q) {              // stage q % kRing <- source rows of quad q
                if (q < nq) {
                    float* st = myring + (q % kRing) * 512;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (4 * q + j < cnt) cp_async16(st + j * 128, row_ptr(s_col[4 * q + j]));
                }
                cp_async_commit();
            };
#pragma unroll
            for (int q = 0; q < kRing - 1; ++q) gather(q);
            float4 a_cur[R / 4];
#pragma unroll
            for (int k = 0; k < R / 4; ++k) a_cur[k] = *reinterpret_cast<const float4*>(s_val + 4 * k);
            for (int q = 0; q < nq; ++q) {
                gather(q + kRing - 1);
                cp_async_wait<kRing - 1>();
                const float* st = myring + (q % kRing) * 512;
                if (4 * q + 4 <= cnt) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 x = *reinterpret_cast<const float4*>(st + j * 128);
                        rbu_fma_col<R>(acc, a_cur, s_val + (size_t)(4 * q + j + 1) * R, x);
                    }
                } else {
                    for (int j = 0; 4 * q + j < cnt; ++j) {
                        const float4 x = *reinterpret_cast<const float4*>(st + j * 128);
                        rbu_fma_col<R>(acc, a_cur, s_val + (size_t)(4 * q + j + 1) * R, x);
                    }
                }
            }
        }
        if (live) {
            float* dp = dst + (size_t)t * d_ts + foff;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int row = __ldg(grp_rows + (size_t)g * R + r);
                if (row >= 0)
                    st_f4(dp + (size_t)row * d_ns,
                          make_float4(acc[r][0].x, acc[r][0].y, acc[r][1].x, acc[r][1].y));
            }
        }
    }
    cp_async_wait<0>();
}

static int check_views(const char* who, const void* src, int64_t s_ts, int64_t s_ns, const void* dst,
                       int64_t d_ts, int64_t d_ns, int F) {
    SGP_REQUIRE(src && dst, SGP_EINVAL, "%s: null src/dst", who);
    SGP_REQUIRE(F >= 1, SGP_EINVAL, "%s: F=%d", who, F);
    (void)s_ts; (void)s_ns; (void)d_ts; (void)d_ns;
    return SGP_OK;
}

static bool vec_views(const void* src, int64_t s_ts, int64_t s_ns, const void* dst, int64_t d_ts,
                      int64_t d_ns, int F) {
    return F % 4 == 0 && aligned16(src) && aligned16(dst) && s_ts % 4 == 0 && s_ns % 4 == 0 &&
           d_ts % 4 == 0 && d_ns % 4 == 0;
}

}  // namespace sgp

using namespace sgp;

extern "C" int sgp_spmm_halo(const int32_t* rowptr, const int32_t* col, const float* val,
                             const int32_t* row_order, const float* src, int64_t src_t_stride,
                             int64_t src_n_stride, const float* src2, int64_t src2_t_stride,
                             int64_t src2_n_stride, int n_split, float* dst, int64_t dst_t_stride,
                             int64_t dst_n_stride, int n_rows, int F, int Tc, void* stream) {
    SGP_REQUIRE(rowptr, SGP_EINVAL, "sgp_spmm: null rowptr");
    SGP_REQUIRE(n_rows >= 0 && Tc >= 0, SGP_EINVAL, "sgp_spmm: n_rows=%d Tc=%d", n_rows, Tc);
    if (int rc = check_views("sgp_spmm", src, src_t_stride, src_n_stride, dst, dst_t_stride, dst_n_stride, F)) return rc;
    if (n_rows == 0 || Tc == 0) return SGP_OK;
    if (!src2) n_split = INT32_MAX;
    cudaStream_t st = as_stream(stream);
    const long long total = (long long)Tc * n_rows;
    const long long blocks = (total + kSpmmWarps - 1) / kSpmmWarps;
    SGP_REQUIRE(blocks < (1ll << 31), SGP_EUNSUPPORTED, "sgp_spmm: Tc*n_rows too large for one launch");
    const bool vec = vec_views(src, src_t_stride, src_n_stride, dst, dst_t_stride, dst_n_stride, F) &&
                     (!src2 || (aligned16(src2) && src2_t_stride % 4 == 0 && src2_n_stride % 4 == 0));
    if (vec) {
        if (F <= 128) {
            spmm_csr_vec<1><<<dim3((unsigned)blocks, 1), kSpmmWarps * 32, 0, st>>>(rowptr, col, val, row_order, src, src_t_stride, src_n_stride, src2, src2_t_stride, src2_n_stride, n_split, dst, dst_t_stride, dst_n_stride, n_rows, F, total);
        } else {
            const int ny = (F + 255) / 256;
            spmm_csr_vec<2><<<dim3((unsigned)blocks, ny), kSpmmWarps * 32, 0, st>>>(rowptr, col, val, row_order, src, src_t_stride, src_n_stride, src2, src2_t_stride, src2_n_stride, n_split, dst, dst_t_stride, dst_n_stride, n_rows, F, total);
        }
        SGP_LAUNCH_CHECK("spmm_csr_vec");
    } else {
        const int ny = (F + 127) / 128;
        spmm_csr_scalar<<<dim3((unsigned)blocks, ny), kSpmmWarps * 32, 0, st>>>(rowptr, col, val, row_order, src, src_t_stride, src_n_stride, src2, src2_t_stride, src2_n_stride, n_split, dst, dst_t_stride, dst_n_stride, n_rows, F, total);
        SGP_LAUNCH_CHECK("spmm_csr_scalar");
    }
    return SGP_OK;
}

extern "C" int sgp_spmm(const int32_t* rowptr, const int32_t* col, const float* val,
                        const int32_t* row_order, const float* src, int64_t src_t_stride,
                        int64_t src_n_stride, float* dst, int64_t dst_t_stride, int64_t dst_n_stride,
                        int n_rows, int F, int Tc, void* stream) {
    return sgp_spmm_halo(rowptr, col, val, row_order, src, src_t_stride, src_n_stride, nullptr, 0, 0, 0,
                         dst, dst_t_stride, dst_n_stride, n_rows, F, Tc, stream);
}

extern "C" int sgp_khop_spmm(const int32_t* rowptr, const int32_t* col, const float* val,
                             const int32_t* row_order, float* buf, int64_t t_stride, int64_t n_stride,
                             int block_in, int block_out0, int hops, int N, int F, int Tc,
                             void* stream) {
    SGP_REQUIRE(buf, SGP_EINVAL, "sgp_khop_spmm: null buffer");
    SGP_REQUIRE(hops >= 0 && block_in >= 0 && block_out0 >= 0, SGP_EINVAL,
                "sgp_khop_spmm: hops=%d block_in=%d block_out0=%d", hops, block_in, block_out0);
    SGP_REQUIRE(block_in < block_out0 || block_in >= block_out0 + hops, SGP_EINVAL,
                "sgp_khop_spmm: input block %d lies inside the output range [%d, %d)", block_in,
                block_out0, block_out0 + hops);
    for (int h = 0; h < hops; ++h) {
        const int bi = (h == 0) ? block_in : block_out0 + h - 1;
        const int bo = block_out0 + h;
        int rc = sgp_spmm(rowptr, col, val, row_order, buf + (size_t)bi * F, t_stride, n_stride,
                          buf + (size_t)bo * F, t_stride, n_stride, N, F, Tc, stream);
        if (rc) return rc;
    }
    return SGP_OK;
}

extern "C" int sgp_spmm_rbu(const int32_t* grp_ptr, const int32_t* grp_rows, const int32_t* ucol,
                            const float* uval, int R, int n_groups, const float* src,
                            int64_t src_t_stride, int64_t src_n_stride, float* dst,
                            int64_t dst_t_stride, int64_t dst_n_stride, int F, int Tc, void* stream) {
    return sgp_spmm_rbu_halo(grp_ptr, grp_rows, ucol, uval, R, n_groups, src, src_t_stride, src_n_stride,
                             nullptr, 0, 0, 0, dst, dst_t_stride, dst_n_stride, F, Tc, stream);
}

extern "C" int sgp_spmm_rbu_halo(const int32_t* grp_ptr, const int32_t* grp_rows, const int32_t* ucol,
                                 const float* uval, int R, int n_groups, const float* src,
                                 int64_t src_t_stride, int64_t src_n_stride, const float* src2,
                                 int64_t src2_t_stride, int64_t src2_n_stride, int n_split, float* dst,
                                 int64_t dst_t_stride, int64_t dst_n_stride, int F, int Tc, void* stream) {
    SGP_REQUIRE(grp_ptr && grp_rows && ucol && uval, SGP_EINVAL, "sgp_spmm_rbu: null operator");
    if (int rc = check_views("sgp_spmm_rbu", src, src_t_stride, src_n_stride, dst, dst_t_stride, dst_n_stride, F)) return rc;
    SGP_REQUIRE(R == 4 || R == 8 || R == 16, SGP_EINVAL, "sgp_spmm_rbu: R=%d (4, 8 or 16)", R);
    SGP_REQUIRE(F % 128 == 0, SGP_EUNSUPPORTED, "sgp_spmm_rbu: F=%d is not a multiple of 128", F);
    SGP_REQUIRE(vec_views(src, src_t_stride, src_n_stride, dst, dst_t_stride, dst_n_stride, F) && aligned16(uval),
                SGP_EALIGN, "sgp_spmm_rbu: views must be 16-byte aligned with strides %% 4 == 0");
    SGP_REQUIRE(!src2 || (aligned16(src2) && src2_t_stride % 4 == 0 && src2_n_stride % 4 == 0), SGP_EALIGN,
                "sgp_spmm_rbu: halo view must be 16-byte aligned with strides %% 4 == 0");
    if (n_groups == 0 || Tc == 0) return SGP_OK;
    if (!src2) n_split = INT32_MAX;
    cudaStream_t st = as_stream(stream);
    // tuning knobs, read once per process (never per launch)
    static const int knob_version = getenv("SGP_B200_RBU_KERNEL") ? atoi(getenv("SGP_B200_RBU_KERNEL")) : 0;
    static const int knob_tspan = getenv("SGP_B200_RBU_TSPAN") ? atoi(getenv("SGP_B200_RBU_TSPAN")) : 0;
    static const int knob_minb = getenv("SGP_B200_RBU_MINB") ? atoi(getenv("SGP_B200_RBU_MINB")) : 4;
    // default: the cp.async gather ring (v3) for R = 16, register-staged gathers (v2) for R <= 8
    const int version = (knob_version == 2 || knob_version == 3) ? knob_version : (R == 16 ? 3 : 2);
    // feature slices of at most 4 x 128 columns per launch (a CTA has one warp per 128-column chunk)
    for (int f0 = 0; f0 < F; f0 += 512) {
        const int nfc = (F - f0 < 512 ? F - f0 : 512) / 128;
        const float* s1 = src + f0;
        const float* s2 = src2 ? src2 + f0 : nullptr;
        float* d1 = dst + f0;
        const int tpb = 4 / nfc >= 1 ? 4 / nfc : 1;          // 4 warps per CTA (nfc = 3 -> 1 step, 3 warps)
        int tspan = knob_tspan > 0 ? knob_tspan : tpb;       // one time step per warp: t-major order keeps the gathered panel L2-hot
        tspan = ((tspan + tpb - 1) / tpb) * tpb;
        const int ny = (Tc + tspan - 1) / tspan;
        SGP_REQUIRE(ny <= 65535, SGP_EUNSUPPORTED, "sgp_spmm_rbu: Tc=%d too large for one launch", Tc);
        if (version == 3) {
            const int nwarps = nfc * tpb;
            const size_t smem = ((size_t)(kRbuPiece + 1) * R + kRbuPiece + (size_t)nwarps * kRing * 512) * sizeof(float);
#define SGP_RBU3(RR)                                                                             \
    do {                                                                                         \
        SGP_CUDA(cudaFuncSetAttribute(spmm_rbu_v3<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        spmm_rbu_v3<RR><<<dim3((unsigned)n_groups, ny), nwarps * 32, smem, st>>>(                \
            grp_ptr, grp_rows, ucol, uval, nfc, tpb, tspan, s1, src_t_stride, src_n_stride, s2,  \
            src2_t_stride, src2_n_stride, n_split, d1, dst_t_stride, dst_n_stride, Tc);           \
    } while (0)
            if (R == 4) SGP_RBU3(4);
            else if (R == 8) SGP_RBU3(8);
            else SGP_RBU3(16);
#undef SGP_RBU3
            SGP_LAUNCH_CHECK("spmm_rbu_v3");
            continue;
        }
#define SGP_RBU2(RR, MB)                                                                         \
    spmm_rbu_v2<RR, MB><<<dim3((unsigned)n_groups, ny), nfc * tpb * 32, 0, st>>>(                \
        grp_ptr, grp_rows, ucol, uval, nfc, tpb, tspan, s1, src_t_stride, src_n_stride, s2,      \
        src2_t_stride, src2_n_stride, n_split, d1, dst_t_stride, dst_n_stride, Tc)
        if (R == 4) SGP_RBU2(4, 4);
        else if (R == 8) SGP_RBU2(8, 4);
        else if (knob_minb == 3) SGP_RBU2(16, 3);
        else SGP_RBU2(16, 4);
#undef SGP_RBU2
        SGP_LAUNCH_CHECK("spmm_rbu_v2");
    }
    return SGP_OK;
}
