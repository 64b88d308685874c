// Microbenchmark: fp32 FMA issue rates on sm_100a (scalar FFMA vs packed FFMA2, with and without
// the broadcast-scalar operand form), to establish the real fp32 roofline used in DESIGN.md.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_peak fma_peak.cu && ./fma_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a0) {
    float2 acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a0 + i * 1e-6f;
    float2 x0 = make_float2(a0 * 0.5f, a0 * 0.25f), x1 = make_float2(a0 * 0.125f, a0 * 0.75f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (MODE == 0) {           // scalar FFMA x2
                acc[i].x = fmaf(a[i & 7], x0.x, acc[i].x);
                acc[i].y = fmaf(a[i & 7], x0.y, acc[i].y);
            } else if (MODE == 1) {    // FFMA2, broadcast scalar operand
                acc[i] = __ffma2_rn(make_float2(a[i & 7], a[i & 7]), (i & 1) ? x1 : x0, acc[i]);
            } else {                   // FFMA2, all packed operands
                acc[i] = __ffma2_rn(make_float2(a[i & 7], a[(i + 1) & 7]), (i & 1) ? x1 : x0, acc[i]);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int blocks_per_sm, float* out) {
    const int iters = 4000, sms = 148;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms * blocks_per_sm, 256>>>(out, 10, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<sms * blocks_per_sm, 256>>>(out, iters, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)sms * blocks_per_sm * 256 * iters * 64;   // 64 FMA per thread-iter
    printf("%-28s %d CTA/SM (%2d warps/SM): %7.2f TFLOP/s  (%.1f FMA/clk/SM @1.965GHz)\n", name,
           blocks_per_sm, blocks_per_sm * 8, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
}

int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    for (int b : {1, 2, 4}) {
        run<0>("FFMA scalar", b, out);
        run<1>("FFMA2 broadcast-scalar", b, out);
        run<2>("FFMA2 packed", b, out);
    }
    return 0;
}
