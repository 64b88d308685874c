// K1-TC — the leaky-ESN scan on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-accurate.
//
// Same contract as sgp_reservoir_scan (one layer, Tc steps, state carried; replaces the Python
// loop of lib/nn/reservoir/reservoir.py:158-186 around :77-81) for H in {128, 256} and Fin <= 8.
// Per CTA: a tile of 128 nodes marches over time; per step
//     D[m, n] = sum_k h[m, k] * W_hh[n, k]              m: 128 nodes (TMEM lanes), n, k: H
// is a dense [128 x H] x [H x H] GEMM issued as 3xTF32 tcgen05.mma (a 1000-step
// recurrence needs fp32-level accuracy; plain TF32 is 8.7e-4 off, the 3-pass split 1.4e-6):
//   * A hi  = the state itself, fp32, K-major SWIZZLE_128B in shared memory (the tensor core
//     ignores the low 13 mantissa bits: measured, tools/microbench/umma_tf32_test.cu);
//   * A lo  = h - tf32(h), resident in TMEM (TS-mode MMA), written with tcgen05.st;
//   * B     = W_hh, split hi / lo at pack time into [128 n x 32 k] K-major SWIZZLE_128B images
//     (sgp_reservoir_tc_pack) that stream from L2 through a cp.async ring every step — W_hh in
//     two tf32 parts is 512 KB at H = 256, more than two SMs' shared memory;
//   * D     = fp32 accumulators in TMEM ([128 lanes x H columns]); TMEM is exactly full at H = 256.
// Epilogue (8 warps, thread = node): tcgen05.ld the pre-activations, add bias and the small input
// projection x_t W_ih^T, activation, leaky blend with the old state (read back from the A-hi
// tile), then write the new state three times: fp32 into the A-hi tile, its lo part into TMEM,
// and the fp32 row into its feature block of the encoder output in HBM.
// Warp roles: 0-7 epilogue, 8-11 W producers, 12 MMA issuer; all hand-offs through mbarriers.
// Bound: tensor pipe (96 N=256 MMAs = 12.3k cycles per 128-node step at H = 256) + the serial
// epilogue; the CUDA-core kernel needs ~155k cycles for the same 128 nodes.
#include <stdlib.h>

#include "common.cuh"

namespace sgp {

constexpr int kRtEpiWarps = 8, kRtProdWarps = 4;
constexpr int kRtThreads = (kRtEpiWarps + kRtProdWarps + 1) * 32;
constexpr int kRtWStage = 128 * 32 * 4;      // 16 KB: one [128 n x 32 k] image (hi OR lo)
constexpr int kRtWStages = 5;
constexpr int kRtLag = 3;                    // cp.async groups a producer thread keeps in flight
constexpr int kRtMaxFin = 8;

__device__ __forceinline__ uint32_t rt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rt_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(rt_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void rt_mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(rt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void rt_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(rt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool rt_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
// bounded warp-wide wait (see spmm_tc.cu)
__device__ __forceinline__ bool rt_wait(uint64_t* bar, uint32_t parity, volatile int* abort_s, int* err, int lane) {
    const uint32_t a = rt_smem_u32(bar);
    uint32_t done = 0;
#pragma unroll 1
    for (int it = 0; it < (1 << 24); ++it) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return true;
        if ((it & 63) == 63 && *abort_s) return false;
    }
    if (lane == 0) {
        *abort_s = 1;
        atomicExch(err, 1);
    }
    return false;
}

// tanh with 1e-7 absolute error from two MUFU ops: 1 - 2 / (1 + 2^(2 log2(e) x)); the argument is
// clamped so that 2^y stays finite.  (tanhf costs ~20 instructions; the epilogue evaluates 32k of
// them per 128-node step and is serial with the MMAs.)
__device__ __forceinline__ float fast_tanh(float x) {
    const float y = fminf(fmaxf(x, -15.f), 15.f) * 2.885390081777927f;     // 2 log2(e) x
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    return fmaf(-2.f, r, 1.f);
}

#define SGP_RT_LD32(addr, v)                                                                         \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15," \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                       \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),   \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),          \
                   "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),        \
                   "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),        \
                   "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                              \
                 : "r"(addr))
#define SGP_RT_ST32(addr, v)                                                                         \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,"  \
                 "%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"       \
                 :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),   \
                    "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]),           \
                    "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),        \
                    "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),        \
                    "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory")

// byte offset of element (row m, column k) of the K-major SWIZZLE_128B state tile [128 x H]:
// k-blocks of 32 columns (16 KB each), atoms of 8 rows x 128 B, 16-byte chunk XOR (row % 8)
__device__ __forceinline__ uint32_t a_tile_offset(int m, int k) {
    return (uint32_t)((k >> 5) * 16384 + (m >> 3) * 1024 + (m & 7) * 128 + ((((k & 31) >> 2) ^ (m & 7)) << 4) + (k & 3) * 4);
}

template <int H>
__global__ void __launch_bounds__(kRtThreads, 1)
reservoir_tc_kernel(const float* __restrict__ x, int64_t x_ts, int64_t x_ns, int Fin,
                    const float* __restrict__ wimg /* [H/32][H/128][2][128*32] */,
                    const float* __restrict__ w_ih /* [H, Fin] */, const float* __restrict__ bias,
                    float alpha, float oma, int act,
                    float* __restrict__ h_state, float* __restrict__ out, int64_t o_ts, int64_t o_ns,
                    int Tc, int N, int* err) {
    constexpr int NH = H / 128;                 // output-column halves (MMA N = 128)
    constexpr int NC = H / 32;                  // k-chunks of 32
    constexpr int A_BYTES = 128 * H * 4;        // state tile
    constexpr int kStagesPerStep = NC * NH * 2; // (chunk, half, hi|lo) images per step
    constexpr int TMEM_COLS = (2 * H <= 256) ? 256 : 512;
    constexpr int ALO_OFF = H;                  // TMEM columns [0,H) = D, [H,2H) = A lo
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (rt_smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - rt_smem_u32(smem_raw));
    // [A hi tile | W ring | w_ih (kRtMaxFin x H) | bias (H)]
    const uint32_t a_base = smem_base, w_base = smem_base + A_BYTES;
    float* wih_s = reinterpret_cast<float*>(smem + A_BYTES + kRtWStages * kRtWStage);
    float* bias_s = wih_s + kRtMaxFin * H;
    __shared__ uint64_t wfull[kRtWStages], wempty[kRtWStages], acc_ready, a_ready;
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int abort_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * 128;

    if (tid == 0) {
        abort_s = 0;
        for (int s = 0; s < kRtWStages; ++s) {
            rt_mbar_init(&wfull[s], kRtProdWarps);
            rt_mbar_init(&wempty[s], 1);
        }
        rt_mbar_init(&acc_ready, 1);
        rt_mbar_init(&a_ready, kRtEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < kRtMaxFin * H; i += kRtThreads) {
        const int f = i / H, n = i % H;
        wih_s[i] = (f < Fin) ? w_ih[(size_t)n * Fin + f] : 0.f;      // [f][n]
    }
    for (int i = tid; i < H; i += kRtThreads) bias_s[i] = bias[i];
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(rt_smem_u32(&tmem_base_s)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;

    if (warp >= kRtEpiWarps && warp < kRtEpiWarps + kRtProdWarps) {
        // ================= producers: stream the W images, kStagesPerStep per time step ========
        const int ptid = tid - kRtEpiWarps * 32;
        const long long total = (long long)Tc * kStagesPerStep;
        bool ok = true;
        int s = 0, ph = 0, sig = 0;
        for (long long j = 0; j < total + kRtLag && ok; ++j) {
            if (j < total) {
                if (ph > 0 && !rt_wait(&wempty[s], (ph - 1) & 1, &abort_s, err, lane)) { ok = false; break; }
                const float* src = wimg + (size_t)(j % kStagesPerStep) * (kRtWStage / 4);
                const uint32_t dst = w_base + s * kRtWStage;
#pragma unroll
                for (int q = 0; q < kRtWStage / 16 / (kRtProdWarps * 32); ++q)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n"
                                 :: "r"(dst + (q * kRtProdWarps * 32 + ptid) * 16),
                                    "l"(src + (q * kRtProdWarps * 32 + ptid) * 4));
                if (++s == kRtWStages) { s = 0; ++ph; }
            }
            cp_async_commit();
            if (j >= kRtLag) {
                cp_async_wait<kRtLag>();
                __syncwarp();
                if (lane == 0) rt_mbar_arrive(&wfull[sig]);
                if (++sig == kRtWStages) sig = 0;
            }
        }
        cp_async_wait<0>();
    } else if (warp < kRtEpiWarps) {
        // ================= epilogue warps: thread = node ========================================
        const int q4 = warp & 3, half = warp >> 2;          // TMEM lane quarter, column half
        const int m = q4 * 32 + lane, node = n0 + m;
        const bool live = node < N;
        const uint32_t lane_addr = tmem_d + ((uint32_t)(q4 * 32) << 16);
        constexpr int COLS = H / 2;                          // columns per epilogue warp
        const int cbeg = half * COLS;
        // ---- initial state -> A hi tile (smem) and A lo (TMEM) --------------------------------
#pragma unroll 1
        for (int c0 = cbeg; c0 < cbeg + COLS; c0 += 32) {
            uint32_t lo[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live) v = *reinterpret_cast<const float4*>(h_state + (size_t)node * H + c0 + j);
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};"
                             :: "r"(a_base + a_tile_offset(m, c0 + j)), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                lo[j + 0] = __float_as_uint(v.x - __uint_as_float(__float_as_uint(v.x) & 0xffffe000u));
                lo[j + 1] = __float_as_uint(v.y - __uint_as_float(__float_as_uint(v.y) & 0xffffe000u));
                lo[j + 2] = __float_as_uint(v.z - __uint_as_float(__float_as_uint(v.z) & 0xffffe000u));
                lo[j + 3] = __float_as_uint(v.w - __uint_as_float(__float_as_uint(v.w) & 0xffffe000u));
            }
            SGP_RT_ST32(lane_addr + ALO_OFF + c0, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) rt_mbar_arrive(&a_ready);

        float xr[kRtMaxFin];
#pragma unroll
        for (int f = 0; f < kRtMaxFin; ++f)
            xr[f] = (live && f < Fin) ? __ldg(x + (size_t)node * x_ns + f) : 0.f;
        bool ok = true;
        for (int t = 0; t < Tc && ok; ++t) {
            float xn[kRtMaxFin];                      // x_{t+1}: in flight while this step is finished
#pragma unroll
            for (int f = 0; f < kRtMaxFin; ++f)
                xn[f] = (live && f < Fin && t + 1 < Tc) ? __ldg(x + (size_t)(t + 1) * x_ts + (size_t)node * x_ns + f) : 0.f;
            if (!rt_wait(&acc_ready, t & 1, &abort_s, err, lane)) { ok = false; break; }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float* orow = out + (size_t)t * o_ts + (size_t)node * o_ns;
#pragma unroll 1
            for (int c0 = cbeg; c0 < cbeg + COLS; c0 += 32) {
                uint32_t d[32];
                SGP_RT_LD32(lane_addr + c0, d);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                uint32_t lo[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + j);
                    float z[4] = {__uint_as_float(d[j]) + b4.x, __uint_as_float(d[j + 1]) + b4.y,
                                  __uint_as_float(d[j + 2]) + b4.z, __uint_as_float(d[j + 3]) + b4.w};
#pragma unroll
                    for (int f = 0; f < kRtMaxFin; ++f) {
                        if (f < Fin) {
                            const float4 w4 = *reinterpret_cast<const float4*>(wih_s + f * H + c0 + j);
                            z[0] = fmaf(xr[f], w4.x, z[0]); z[1] = fmaf(xr[f], w4.y, z[1]);
                            z[2] = fmaf(xr[f], w4.z, z[2]); z[3] = fmaf(xr[f], w4.w, z[3]);
                        }
                    }
                    float4 ho;
                    const uint32_t ha = a_base + a_tile_offset(m, c0 + j);
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                 : "=f"(ho.x), "=f"(ho.y), "=f"(ho.z), "=f"(ho.w) : "r"(ha));
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (act == SGP_ACT_TANH) z[e] = fast_tanh(z[e]);
                        else if (act == SGP_ACT_RELU) z[e] = fmaxf(z[e], 0.f);
                    }
                    float4 hn;
                    hn.x = oma * ho.x + alpha * z[0];
                    hn.y = oma * ho.y + alpha * z[1];
                    hn.z = oma * ho.z + alpha * z[2];
                    hn.w = oma * ho.w + alpha * z[3];
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};"
                                 :: "r"(ha), "f"(hn.x), "f"(hn.y), "f"(hn.z), "f"(hn.w) : "memory");
                    if (live) st_f4(orow + c0 + j, hn);
                    lo[j + 0] = __float_as_uint(hn.x - __uint_as_float(__float_as_uint(hn.x) & 0xffffe000u));
                    lo[j + 1] = __float_as_uint(hn.y - __uint_as_float(__float_as_uint(hn.y) & 0xffffe000u));
                    lo[j + 2] = __float_as_uint(hn.z - __uint_as_float(__float_as_uint(hn.z) & 0xffffe000u));
                    lo[j + 3] = __float_as_uint(hn.w - __uint_as_float(__float_as_uint(hn.w) & 0xffffe000u));
                }
                SGP_RT_ST32(lane_addr + ALO_OFF + c0, lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // new state -> tensor proxy
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) rt_mbar_arrive(&a_ready);
#pragma unroll
            for (int f = 0; f < kRtMaxFin; ++f) xr[f] = xn[f];
        }
        // ---- carry the state ---------------------------------------------------------------------
        if (ok && live) {
#pragma unroll 1
            for (int c0 = cbeg; c0 < cbeg + COLS; c0 += 4) {
                float4 v;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a_base + a_tile_offset(m, c0)));
                *reinterpret_cast<float4*>(h_state + (size_t)node * H + c0) = v;
            }
        }
    } else {
        // ================= MMA issuer ==============================================================
        // kind::tf32, fp32 accumulate, A K-major (smem hi / TMEM lo), B K-major, N = 128, M = 128
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) |
                                   ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t hi32 = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO 1024, version, SW128
        constexpr uint32_t lo32 = (16u >> 4) << 16;                            // LBO
        auto desc = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
        auto mma_ss = [](uint32_t d, uint64_t da, uint64_t db, uint32_t idesc_, uint32_t acc) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                         :: "r"(d), "l"(da), "l"(db), "r"(idesc_), "r"(acc) : "memory");
        };
        auto mma_ts = [](uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc_, uint32_t acc) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                         :: "r"(d), "r"(a_tmem), "l"(db), "r"(idesc_), "r"(acc) : "memory");
        };
        const uint32_t a0 = lo32 | (a_base >> 4), w0 = lo32 | (w_base >> 4);
        bool ok = true;
        int s = 0, ph = 0;
        for (int t = 0; t < Tc && ok; ++t) {
            if (!rt_wait(&a_ready, t & 1, &abort_s, err, lane)) { ok = false; break; }   // state of step t-1 in place
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < NC && ok; ++c) {
#pragma unroll 1
                for (int hh = 0; hh < NH; ++hh) {
                    // hi image: Ah x Wh and Al x Wh; lo image: Ah x Wl
                    const uint32_t dcol = tmem_d + hh * 128;
                    if (!rt_wait(&wfull[s], ph & 1, &abort_s, err, lane)) { ok = false; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (rt_elect_one()) {
                        const uint32_t wh = w0 + s * (kRtWStage >> 4);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t da = desc(a0 + c * (16384 >> 4) + ks * 2, hi32);
                            const uint64_t db = desc(wh + ks * 2, hi32);
                            mma_ss(dcol, da, db, idesc, (c | ks) ? 1u : 0u);
                            mma_ts(dcol, tmem_d + ALO_OFF + c * 32 + ks * 8, db, idesc, 1u);
                        }
                        rt_commit(&wempty[s]);
                    }
                    __syncwarp();
                    if (++s == kRtWStages) { s = 0; ++ph; }
                    if (!rt_wait(&wfull[s], ph & 1, &abort_s, err, lane)) { ok = false; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (rt_elect_one()) {
                        const uint32_t wl = w0 + s * (kRtWStage >> 4);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t da = desc(a0 + c * (16384 >> 4) + ks * 2, hi32);
                            const uint64_t db = desc(wl + ks * 2, hi32);
                            mma_ss(dcol, da, db, idesc, 1u);
                        }
                        rt_commit(&wempty[s]);
                    }
                    __syncwarp();
                    if (++s == kRtWStages) { s = 0; ++ph; }
                }
            }
            if (ok && rt_elect_one()) rt_commit(&acc_ready);      // pre-activations of step t complete
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "r"(TMEM_COLS));
}

// W_hh [H, H] -> images [H/32 chunks][H/128 halves][hi | lo][128 n x 32 k, K-major SWIZZLE_128B]
__global__ void reservoir_tc_pack_kernel(const float* __restrict__ w_hh, int H, float* __restrict__ wimg) {
    const int NH = H / 128;
    const int64_t total = (int64_t)H * H;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / H), k = (int)(i % H);
        const float w = w_hh[i];
        const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
        const int c = k >> 5, kk = k & 31, hh = n >> 7, nn = n & 127;
        const size_t img = ((size_t)(c * NH + hh) * 2) * (128 * 32);
        const int off = (nn >> 3) * 256 + (nn & 7) * 32 + (((kk >> 2) ^ (nn & 7)) << 2) + (kk & 3);    // floats
        wimg[img + off] = hi;
        wimg[img + 128 * 32 + off] = w - hi;
    }
}

}  // namespace sgp

using namespace sgp;

extern "C" int sgp_reservoir_tc_pack(const float* w_hh, int H, float* wimg, void* stream) {
    SGP_REQUIRE(w_hh && wimg, SGP_EINVAL, "sgp_reservoir_tc_pack: null pointer");
    SGP_REQUIRE(H == 128 || H == 256, SGP_EUNSUPPORTED, "sgp_reservoir_tc_pack: H=%d (128 or 256)", H);
    reservoir_tc_pack_kernel<<<(H * H + 255) / 256, 256, 0, as_stream(stream)>>>(w_hh, H, wimg);
    SGP_LAUNCH_CHECK("reservoir_tc_pack");
    return SGP_OK;
}

extern "C" int sgp_reservoir_scan_tc(const float* x, int64_t x_t_stride, int64_t x_n_stride, int Fin,
                                     const float* wimg, const float* w_ih, const float* bias, float alpha,
                                     float one_minus_alpha, int act, float* h_state, float* out,
                                     int64_t out_t_stride, int64_t out_n_stride, int Tc, int N, int H,
                                     int* err_flag, void* stream) {
    SGP_REQUIRE(x && wimg && w_ih && bias && h_state && out && err_flag, SGP_EINVAL,
                "sgp_reservoir_scan_tc: null pointer");
    SGP_REQUIRE(H == 128 || H == 256, SGP_EUNSUPPORTED, "sgp_reservoir_scan_tc: H=%d (128 or 256)", H);
    SGP_REQUIRE(Fin >= 1 && Fin <= kRtMaxFin, SGP_EUNSUPPORTED, "sgp_reservoir_scan_tc: Fin=%d (1..%d)", Fin, kRtMaxFin);
    SGP_REQUIRE(act == SGP_ACT_TANH || act == SGP_ACT_RELU || act == SGP_ACT_IDENTITY, SGP_EUNSUPPORTED,
                "sgp_reservoir_scan_tc: activation %d not supported on this path", act);
    SGP_REQUIRE(aligned16(out) && aligned16(h_state) && aligned16(wimg) && out_t_stride % 4 == 0 &&
                    out_n_stride % 4 == 0, SGP_EALIGN, "sgp_reservoir_scan_tc: views must be 16-byte aligned");
    if (N == 0 || Tc == 0) return SGP_OK;
    const size_t smem = (size_t)128 * H * 4 + (size_t)kRtWStages * kRtWStage + (size_t)(kRtMaxFin + 1) * H * 4 + 1024;
    const int grid = (N + 127) / 128;
#define SGP_RT(H_)                                                                                      \
    do {                                                                                                \
        SGP_CUDA(cudaFuncSetAttribute(reservoir_tc_kernel<H_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        reservoir_tc_kernel<H_><<<grid, kRtThreads, smem, as_stream(stream)>>>(                         \
            x, x_t_stride, x_n_stride, Fin, wimg, w_ih, bias, alpha, one_minus_alpha, act, h_state, out, \
            out_t_stride, out_n_stride, Tc, N, err_flag);                                               \
    } while (0)
    if (H == 256) SGP_RT(256); else SGP_RT(128);
#undef SGP_RT
    SGP_LAUNCH_CHECK("reservoir_scan_tc");
    return SGP_OK;
}
