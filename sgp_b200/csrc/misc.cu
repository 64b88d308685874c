// K4 (global-mean block), output checksum sink, halo row gather, and the ABI's error plumbing.
#include "common.cuh"

namespace sgp {

std::atomic<int64_t> g_launches{0};
std::atomic<int> g_tc_cta_limit{kNumSMs};
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// sums[t, f] = sum_n src[t, n, f]; grid (ceil(F/128), Tc, node-splits), atomics across splits.
__global__ void node_sum_kernel(const float* __restrict__ src, int64_t s_ts, int64_t s_ns,
                                float* __restrict__ sums, int N, int F) {
    __shared__ float part[8][128];
    const int fx = threadIdx.x & 127, ny = threadIdx.x >> 7;   // 1024 threads: 128 features x 8 node lanes
    const int f = blockIdx.x * 128 + fx, t = blockIdx.y;
    const int per = (N + gridDim.z - 1) / gridDim.z;
    const int nb = blockIdx.z * per, ne = min(N, nb + per);
    float a = 0.f;
    if (f < F)
        for (int n = nb + ny; n < ne; n += 8) a += __ldg(src + (size_t)t * s_ts + (size_t)n * s_ns + f);
    part[ny][fx] = a;
    __syncthreads();
    if (ny == 0 && f < F) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += part[k][fx];
        atomicAdd(sums + (size_t)t * F + f, s);
    }
}

__global__ void node_mean_bcast_kernel(const float* __restrict__ sums, float n_total,
                                       float* __restrict__ dst, int64_t d_ts, int64_t d_ns,
                                       int N, int F, int Tc) {
    const int64_t total = (int64_t)Tc * N * F;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int f = (int)(i % F);
        const int64_t tn = i / F;
        const int n = (int)(tn % N), t = (int)(tn / N);
        dst[(size_t)t * d_ts + (size_t)n * d_ns + f] = sums[(size_t)t * F + f] / n_total;
    }
}

__global__ void checksum_kernel(const float* __restrict__ buf, int64_t count, double* acc) {
    double s = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x)
        s += (double)buf[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double w[32];
    if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? w[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) atomicAdd(acc, s);
    }
}

// acc += sum over a strided [Tc, N, F] view (the feature block a CUDA-core kernel just wrote): the
// sink for paths whose producing kernel has no fused checksum.
__global__ void checksum_view_kernel(const float* __restrict__ src, int64_t s_ts, int64_t s_ns, int N, int F,
                                     int Tc, double* acc) {
    double s = 0.0;
    const int64_t rows = (int64_t)Tc * N;
    const int lanes_f = F < 32 ? F : 32;                          // threads along the feature axis
    const int fx = threadIdx.x % lanes_f, ry = threadIdx.x / lanes_f, rpb = blockDim.x / lanes_f;
    if (ry < rpb)
        for (int64_t r = (int64_t)blockIdx.x * rpb + ry; r < rows; r += (int64_t)gridDim.x * rpb) {
            const float* p = src + (size_t)(r / N) * s_ts + (size_t)(r % N) * s_ns;
            float part = 0.f;
            for (int f = fx; f < F; f += lanes_f) part += __ldg(p + f);
            s += (double)part;
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double w[32];
    if ((threadIdx.x & 31) == 0) w[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? w[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) atomicAdd(acc, s);
    }
}

// dst[t, k, :] = src[t, index[k], :] — halo packing.  VEC: 16 bytes per thread (F % 4 == 0, aligned
// views): the rows are 0.5-1 KB, so a warp moves whole 512-byte row pieces; grid.y walks time.
template <bool VEC>
__global__ void gather_rows_kernel(const float* __restrict__ src, int64_t s_ts, int64_t s_ns,
                                   const int32_t* __restrict__ index, int n_index,
                                   float* __restrict__ dst, int64_t d_ts, int64_t d_ns, int F, int Tc) {
    const int W = VEC ? F / 4 : F;                       // work units per row
    const uint32_t per_t = (uint32_t)n_index * (uint32_t)W;
    for (int t = blockIdx.y; t < Tc; t += gridDim.y) {
        const float* sp = src + (size_t)t * s_ts;
        float* dp = dst + (size_t)t * d_ts;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < per_t; i += gridDim.x * blockDim.x) {
            const uint32_t k = i / (uint32_t)W, f = i - k * (uint32_t)W;
            const size_t so = (size_t)__ldg(index + k) * s_ns, dofs = (size_t)k * d_ns;
            if (VEC)
                reinterpret_cast<float4*>(dp + dofs)[f] = ldg_f4_stream(sp + so + 4 * f);
            else
                dp[dofs + f] = __ldg(sp + so + f);
        }
    }
}

// Halo PUSH over NVLink: row index[k] of `src` goes, for every time step of the chunk, to the address
// dst_addr[k] + t * d_ts — a slot of ANOTHER rank's halo buffer, mapped into this process (peer
// memory: torch symmetric memory).  The pack and the transfer are one kernel: 16-byte stores
// straight into the peer's HBM, no send buffer, no collective call.
// Launch shape (256 threads, ~34 registers) measured, not guessed: a 64-thread / 32-register variant
// that fits beside the persistent hop CTAs (which leave 6.6 K registers per SM) did overlap with the
// SpMM of the other chunk — and slowed that SpMM by 24 % (its single-thread issue loops share the
// schedulers), 292 ms per pass at 2 GPUs against 272 ms for this shape, which runs in the gap
// between two hop launches at NVLink rate.  (Measured against the tf32 hop.  The fp16x3 hop's CTA leaves
// 9-12 K registers, so ONE CTA of this shape can sit beside it; the 2 / 4 / 8-GPU lines of the final
// kernels were measured in that configuration, with 12 SMs kept free of hop CTAs from 4 GPUs up.)
__global__ void push_rows_kernel(const float* __restrict__ src, int64_t s_ts, int64_t s_ns,
                                 const int32_t* __restrict__ index, const int64_t* __restrict__ dst_addr,
                                 int n_index, int64_t d_ts, int F4, int Tc) {
    const uint32_t per_t = (uint32_t)n_index * (uint32_t)F4;
    for (int t = blockIdx.y; t < Tc; t += gridDim.y) {
        const float* sp = src + (size_t)t * s_ts;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < per_t; i += gridDim.x * blockDim.x) {
            const uint32_t k = i / (uint32_t)F4, f = i - k * (uint32_t)F4;
            float4* dp = reinterpret_cast<float4*>(reinterpret_cast<float*>(__ldg(dst_addr + k)) + (size_t)t * d_ts);
            dp[f] = ldg_f4_stream(sp + (size_t)__ldg(index + k) * s_ns + 4 * f);
        }
    }
}

// dst[m, :] = src[t_idx[m], n_idx[m], :] — the IID (t, n) sampler's gather from the device-resident
// encoder output (lib/datasets/iid_dataset.py:57-99: `tens[(step_index, None, None, node_index)]`).
// One warp per sample walks the row in 512-byte pieces (16 bytes per lane) when VEC, else scalars.
template <bool VEC>
__global__ void gather_tn_kernel(const float* __restrict__ src, int64_t s_ts, int64_t s_ns, int F,
                                 const int64_t* __restrict__ t_idx, const int64_t* __restrict__ n_idx,
                                 int64_t M, float* __restrict__ dst, int64_t d_ms) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t m = warp; m < M; m += n_warps) {
        const float* sp = src + (size_t)__ldg(t_idx + m) * s_ts + (size_t)__ldg(n_idx + m) * s_ns;
        float* dp = dst + (size_t)m * d_ms;
        if (VEC) {
            for (int f = lane; f < F / 4; f += 32) reinterpret_cast<float4*>(dp)[f] = ldg_f4_stream(sp + 4 * f);
        } else {
            for (int f = lane; f < F; f += 32) dp[f] = __ldg(sp + f);
        }
    }
}

static int grid_1d(int64_t total, int threads) {
    int64_t g = (total + threads - 1) / threads;
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace sgp

using namespace sgp;

extern "C" int sgp_tc_set_cta_limit(int n_ctas) {
    SGP_REQUIRE(n_ctas >= 1, SGP_EINVAL, "sgp_tc_set_cta_limit: n_ctas=%d", n_ctas);
    g_tc_cta_limit.store(n_ctas < kNumSMs ? n_ctas : kNumSMs, std::memory_order_relaxed);
    return SGP_OK;
}

extern "C" int sgp_version(void) { return SGP_B200_ABI_VERSION; }
extern "C" const char* sgp_last_error(void) { return g_err; }
extern "C" int64_t sgp_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int sgp_node_sum(const float* src, int64_t src_t_stride, int64_t src_n_stride, float* sums,
                            int N, int F, int Tc, void* stream) {
    SGP_REQUIRE(src && sums, SGP_EINVAL, "sgp_node_sum: null pointer");
    SGP_REQUIRE(N >= 0 && F >= 1 && Tc >= 0, SGP_EINVAL, "sgp_node_sum: N=%d F=%d Tc=%d", N, F, Tc);
    if (Tc == 0) return SGP_OK;
    cudaStream_t st = as_stream(stream);
    SGP_CUDA(cudaMemsetAsync(sums, 0, (size_t)Tc * F * sizeof(float), st));
    if (N == 0) return SGP_OK;
    SGP_REQUIRE(Tc <= 65535, SGP_EUNSUPPORTED, "sgp_node_sum: Tc=%d > 65535", Tc);
    const int fx = (F + 127) / 128;
    int splits = (4 * kNumSMs) / (fx * Tc);
    splits = splits < 1 ? 1 : (splits > (N + 255) / 256 ? (N + 255) / 256 : splits);
    node_sum_kernel<<<dim3(fx, Tc, splits), 1024, 0, st>>>(src, src_t_stride, src_n_stride, sums, N, F);
    SGP_LAUNCH_CHECK("node_sum");
    return SGP_OK;
}

extern "C" int sgp_node_mean_broadcast(const float* sums, int64_t N_total, float* dst,
                                       int64_t dst_t_stride, int64_t dst_n_stride, int N, int F, int Tc,
                                       void* stream) {
    SGP_REQUIRE(sums && dst, SGP_EINVAL, "sgp_node_mean_broadcast: null pointer");
    SGP_REQUIRE(N_total >= 1 && N >= 0 && F >= 1 && Tc >= 0, SGP_EINVAL,
                "sgp_node_mean_broadcast: N_total=%lld N=%d F=%d Tc=%d", (long long)N_total, N, F, Tc);
    const int64_t total = (int64_t)Tc * N * F;
    if (total == 0) return SGP_OK;
    node_mean_bcast_kernel<<<grid_1d(total, 256), 256, 0, as_stream(stream)>>>(
        sums, (float)N_total, dst, dst_t_stride, dst_n_stride, N, F, Tc);
    SGP_LAUNCH_CHECK("node_mean_broadcast");
    return SGP_OK;
}

extern "C" int sgp_checksum(const float* buf, int64_t count, double* acc, void* stream) {
    SGP_REQUIRE(buf && acc && count >= 0, SGP_EINVAL, "sgp_checksum: bad arguments");
    if (count == 0) return SGP_OK;
    checksum_kernel<<<grid_1d(count, 256 * 8), 256, 0, as_stream(stream)>>>(buf, count, acc);
    SGP_LAUNCH_CHECK("checksum");
    return SGP_OK;
}

extern "C" int sgp_checksum_view(const float* src, int64_t src_t_stride, int64_t src_n_stride, int N, int F,
                                 int Tc, double* acc, void* stream) {
    SGP_REQUIRE(src && acc && N >= 0 && F >= 1 && Tc >= 0, SGP_EINVAL, "sgp_checksum_view: bad arguments");
    const int64_t rows = (int64_t)Tc * N;
    if (rows == 0) return SGP_OK;
    const int lanes_f = F < 32 ? F : 32, rpb = 256 / lanes_f;
    checksum_view_kernel<<<grid_1d(rows, rpb * 4), 256, 0, as_stream(stream)>>>(src, src_t_stride, src_n_stride,
                                                                              N, F, Tc, acc);
    SGP_LAUNCH_CHECK("checksum_view");
    return SGP_OK;
}

extern "C" int sgp_push_rows(const float* src, int64_t src_t_stride, int64_t src_n_stride, const int32_t* index,
                             const int64_t* dst_addr, int n_index, int64_t dst_t_stride, int F, int Tc,
                             void* stream) {
    SGP_REQUIRE(src && ((index && dst_addr) || n_index == 0), SGP_EINVAL, "sgp_push_rows: null pointer");
    if (Tc <= 0 || n_index <= 0 || F <= 0) return SGP_OK;
    SGP_REQUIRE(F % 4 == 0 && aligned16(src) && src_t_stride % 4 == 0 && src_n_stride % 4 == 0 && dst_t_stride % 4 == 0,
                SGP_EALIGN, "sgp_push_rows: F %% 4 == 0 and 16-byte aligned views required");
    const int64_t per_t = (int64_t)n_index * (F / 4);
    SGP_REQUIRE(per_t < (1ll << 32), SGP_EUNSUPPORTED, "sgp_push_rows: %d rows x %d features too large", n_index, F);
    const int gy = Tc < 64 ? Tc : 64;
    int gx = (int)((per_t + 255) / 256);
    const int cap = (kNumSMs * 8 + gy - 1) / gy;
    gx = gx < 1 ? 1 : (gx > cap ? cap : gx);
    push_rows_kernel<<<dim3(gx, gy), 256, 0, as_stream(stream)>>>(src, src_t_stride, src_n_stride, index, dst_addr,
                                                                  n_index, dst_t_stride, F / 4, Tc);
    SGP_LAUNCH_CHECK("push_rows");
    return SGP_OK;
}

extern "C" int sgp_gather_tn(const float* src, int64_t src_t_stride, int64_t src_n_stride, int T, int N, int F,
                             const int64_t* t_idx, const int64_t* n_idx, int64_t M, float* dst,
                             int64_t dst_m_stride, void* stream) {
    SGP_REQUIRE(src && dst && ((t_idx && n_idx) || M == 0), SGP_EINVAL, "sgp_gather_tn: null pointer");
    SGP_REQUIRE(T >= 0 && N >= 0 && F >= 1 && M >= 0 && dst_m_stride >= F, SGP_EINVAL,
                "sgp_gather_tn: T=%d N=%d F=%d M=%lld", T, N, F, (long long)M);
    if (M == 0) return SGP_OK;
    const bool vec = F % 4 == 0 && aligned16(src) && aligned16(dst) && src_t_stride % 4 == 0 &&
                     src_n_stride % 4 == 0 && dst_m_stride % 4 == 0;
    const int grid = grid_1d(M, 8);                      // 8 warps (samples) per CTA of 256 threads
    if (vec)
        gather_tn_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(src, src_t_stride, src_n_stride, F, t_idx,
                                                                   n_idx, M, dst, dst_m_stride);
    else
        gather_tn_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(src, src_t_stride, src_n_stride, F, t_idx,
                                                                    n_idx, M, dst, dst_m_stride);
    SGP_LAUNCH_CHECK("gather_tn");
    return SGP_OK;
}

extern "C" int sgp_gather_rows(const float* src, int64_t src_t_stride, int64_t src_n_stride,
                               const int32_t* index, int n_index, float* dst, int64_t dst_t_stride,
                               int64_t dst_n_stride, int F, int Tc, void* stream) {
    SGP_REQUIRE(src && dst && (index || n_index == 0), SGP_EINVAL, "sgp_gather_rows: null pointer");
    if (Tc <= 0 || n_index <= 0 || F <= 0) return SGP_OK;
    const bool vec = F % 4 == 0 && aligned16(src) && aligned16(dst) && src_t_stride % 4 == 0 &&
                     src_n_stride % 4 == 0 && dst_t_stride % 4 == 0 && dst_n_stride % 4 == 0;
    const int64_t per_t = (int64_t)n_index * (vec ? F / 4 : F);
    SGP_REQUIRE(per_t < (1ll << 32), SGP_EUNSUPPORTED, "sgp_gather_rows: %d rows x %d features too large", n_index, F);
    const int gy = Tc < 64 ? Tc : 64;
    int gx = (int)((per_t + 255) / 256);
    const int cap = (kNumSMs * 16 + gy - 1) / gy;
    gx = gx < 1 ? 1 : (gx > cap ? cap : gx);
    if (vec)
        gather_rows_kernel<true><<<dim3(gx, gy), 256, 0, as_stream(stream)>>>(
            src, src_t_stride, src_n_stride, index, n_index, dst, dst_t_stride, dst_n_stride, F, Tc);
    else
        gather_rows_kernel<false><<<dim3(gx, gy), 256, 0, as_stream(stream)>>>(
            src, src_t_stride, src_n_stride, index, n_index, dst, dst_t_stride, dst_n_stride, F, Tc);
    SGP_LAUNCH_CHECK("gather_rows");
    return SGP_OK;
}
