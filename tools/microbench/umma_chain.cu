// Does a chain of tcgen05.mma into ONE accumulator run slower than the same MMAs spread over
// several accumulators?  kind::tf32, M=128, K=8, A from TMEM (TS mode), B K-major SWIZZLE_128B.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_chain umma_chain.cu && ./umma_chain
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
template <int N, int NACC, int NCOMMIT, int MODE>
__global__ void __launch_bounds__(128, 1) rate(int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar, bar2[4]; __shared__ uint32_t tb_s;
    for (int i = threadIdx.x; i < 65536 / 4; i += 128) ((float*)smem)[i] = 1.0f;
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); for (int i = 0; i < 3; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" :: "r"(smem_u32(&bar2[i]))); asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar2[3]))); asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"(smem_u32(&bar2[3])) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tb_s)), "r"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = tb_s;
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (threadIdx.x < 32) {
        long long t0 = 0, t1 = 0, t2 = 0;
        if (elect_one()) {
            const uint64_t hi = ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29))) << 32;
            const uint32_t b0 = ((16u >> 4) << 16) | (smem_u32(smem) >> 4);
            t0 = clock64();
#pragma unroll 1
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int j = 0; j < 12; ++j) {          // like one item of the hop: 4 k-steps x 3 products
                    const uint64_t db = hi | (b0 + (j % 4) * 2 + (j / 4 == 2 ? 512 : 0));
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                                 :: "r"(tb + (j % NACC) * N), "r"(tb + 256 + (j % 4) * 8 + (j / 4 == 1 ? 32 : 0) + (j % NACC) * 64), "l"(db), "r"(idesc), "r"(1u) : "memory");
                }
                if (MODE == 2) { uint32_t dn; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(dn) : "r"(smem_u32(&bar2[3])), "r"(0) : "memory"); if (!dn) break; }
#pragma unroll
                for (int q = 0; q < NCOMMIT; ++q)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar2[q])) : "memory");
                if (MODE == 1) { uint32_t dn; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(dn) : "r"(smem_u32(&bar2[3])), "r"(0) : "memory"); if (!dn) break; }
                if (MODE == 3) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
            t1 = clock64();
        }
        __syncwarp();
        uint32_t done = 0;
        for (long long s = 0; s < (1ll << 26) && !done; ++s)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        t2 = clock64();
        if (t0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tb), "r"(512));
}
template <int N, int NACC, int NCOMMIT, int MODE>
void run(long long* d) {
    long long h[2];
    const int iters = 400;
    cudaFuncSetAttribute(rate<N, NACC, NCOMMIT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; ++rep) {
        rate<N, NACC, NCOMMIT, MODE><<<1, 128, 80 * 1024>>>(iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (rep) printf("mode=%d commits/item=%d ", MODE, NCOMMIT);
        if (rep) printf("N=%3d accumulators=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (%s) -> %.0f MAC/clk\n", N, NACC,
                        (double)h[0] / (iters * 12), (double)h[1] / (iters * 12), cudaGetErrorString(e), 128.0 * N * 8 / ((double)h[1] / (iters * 12)));
    }
}
int main() {
    long long* d; cudaMalloc(&d, 32);
    run<64, 1, 1, 0>(d); run<64, 1, 1, 1>(d); run<64, 1, 0, 1>(d); run<64, 1, 1, 2>(d); run<64, 1, 1, 3>(d);
    return 0;
}
