// tcgen05.mma issue-rate microbenchmark (sm_100a): cycles per kind::tf32 MMA (M=128, K=8) for
// A M-major (SWIZZLE_128B_BASE32B) vs K-major (SWIZZLE_128B), N in {64,128,256}, operands in smem.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t lt) {
    uint64_t d = 0; d |= (uint64_t)((addr >> 4) & 0x3FFF); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)lt << 61; return d;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int nacc>
__global__ void __launch_bounds__(128, 1) rate(int N, int amajor, int iters, long long* out, int distinct, int ts) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar; __shared__ uint32_t tb;
    for (int i = threadIdx.x; i < (8 * 16384 + 2 * 32768) / 4; i += 128) ((float*)smem)[i] = 1.0f;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tb)), "r"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)amajor << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 8 * 16384;
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t a = a0 + (distinct ? (it & 7) * 16384 : 0), b = b0 + (distinct ? (it & 1) * 32768 : 0);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t da = amajor ? make_desc(a + ks * 4096, 512, 2048, 1) : make_desc(a + ks * 32, 16, 1024, 2);
                const uint64_t db = make_desc(b + ks * 32, 16, 1024, 2);
                if (ts) {
                    const uint32_t idts = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                                 :: "r"(tb + (ks % nacc) * 64), "r"(tb + 256 + (distinct ? (it & 3) * 64 : 0) + ks * 8), "l"(db), "r"(idts), "r"(1u) : "memory");
                } else
                mma(tb, da, db, idesc, 1u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
        long long t1 = clock64();
        uint32_t done = 0;
        for (long long s = 0; s < (1ll << 26) && !done; ++s)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0; out[2] = done;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(512));
}
int main() {
    long long* d; cudaMalloc(&d, 32); long long h[3];
    cudaFuncSetAttribute(rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);cudaFuncSetAttribute(rate<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);cudaFuncSetAttribute(rate<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int nacc : {1, 2, 4})
    for (int ts = 1; ts < 2; ++ts)
    for (int distinct = 1; distinct < 2; ++distinct)
    for (int amajor = 1; amajor < 2; ++amajor)
        for (int N : {64, 128})
            for (int rep = 0; rep < 2; ++rep) {
                const int iters = 500;
                if (N * nacc > 256) continue;
                if (nacc == 1) rate<1><<<1, 128, 198 * 1024>>>(N, amajor, iters, d, distinct, ts);
                else if (nacc == 2) rate<2><<<1, 128, 198 * 1024>>>(N, amajor, iters, d, distinct, ts);
                else rate<4><<<1, 128, 198 * 1024>>>(N, amajor, iters, d, distinct, ts);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
                if (rep) printf("nacc=%d ", nacc);
                if (rep) printf("ts=%d distinct=%d A %s-major N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (done=%lld, %s)  -> %.0f MAC/clk\n", ts, distinct, amajor ? "M" : "K", N,
                       (double)h[0] / (iters * 4), (double)h[1] / (iters * 4), h[2], cudaGetErrorString(e), 128.0 * N * 8 / ((double)h[1] / (iters * 4)));
            }
    return 0;
}
