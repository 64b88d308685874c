// Is the hop SpMM's gather bound by a per-SM ingest port or by the chip-wide L2 output?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_ingest l2_ingest.cu && ./l2_ingest
// G persistent CTAs (one per SM, 1024 threads) read random 512-byte row pieces (LDG.128 per lane, 8
// rows in flight per warp: the producers' access pattern) from a footprint that stays in L2.  Run
// with G = 148, 74, 37 CTAs: if bytes/clk/SM is flat the limit is the SM's own L2 port; if it
// grows as G shrinks the limit is shared (L2 slices / crossbar).  The SM clock is measured with
// clock64 against the event time, so the result is in bytes per SM clock whatever the power state.
// `mode` 1 adds a streaming store of 1/5 of the gathered bytes (the hop's output).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__global__ void __launch_bounds__(1024, 1)
ingest(const float* __restrict__ x, int stride, const int* __restrict__ ids, long n_ids, float* out, int mode,
       long long* clocks, float* sink) {
    const int lane = threadIdx.x & 31;
    const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * (long)blockDim.x) >> 5;
    const long long c0 = clock64();
    float acc = 0.f;
    for (long base = warp * 32; base < n_ids; base += n_warps * 32) {
        const int my = ids[base + lane];
#pragma unroll
        for (int j0 = 0; j0 < 32; j0 += 8) {
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = __shfl_sync(0xffffffffu, my, j0 + j);
                v[j] = __ldcg(reinterpret_cast<const float4*>(x + (size_t)r * stride) + lane);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += v[j].x + v[j].w;
            if (mode == 1 && (j0 & 8) == 0 && j0 < 16) {     // 1 stored row piece per ~5 gathered
                float4 o = make_float4(acc, acc, acc, acc);
                __stcs(reinterpret_cast<float4*>(out + ((size_t)(base + j0) % (1 << 22)) * 128) + lane, o);
            }
        }
    }
    if (threadIdx.x == 0) clocks[blockIdx.x] = clock64() - c0;
    if (acc == 123.456f) *sink = acc;
}

int main() {
    const int stride = 1280, R = 100000;                  // 100k rows of 5 KB: the C4 chunk buffer; footprint 51 MB
    float* x; cudaMalloc(&x, (size_t)R * stride * 4); cudaMemset(x, 0, (size_t)R * stride * 4);
    float* out; cudaMalloc(&out, (size_t)(1 << 22) * 128 * 4);
    float* sink; cudaMalloc(&sink, 4);
    long long* clk; cudaMalloc(&clk, 148 * 8);
    const long n_ids = 148L * 32 * 32 * 64;               // 9.7M row pieces = 5 GB of gathers
    std::vector<int> h(n_ids);
    srand(1);
    for (long i = 0; i < n_ids; i += 32) {                // clustered like a group's union: a window of 2048 rows
        const int w0 = (int)(((long)rand() * 32768 + rand()) % (R - 2048));
        for (int j = 0; j < 32; ++j) h[i + j] = w0 + rand() % 2048;
    }
    int* ids; cudaMalloc(&ids, n_ids * 4);
    cudaMemcpy(ids, h.data(), n_ids * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode)
        for (int G : {148, 111, 74, 37}) {
            const long n = n_ids * G / 148;               // same work per CTA at every G
            for (int w = 0; w < 2; ++w) ingest<<<G, 1024>>>(x, stride, ids, n, out, mode, clk, sink);
            cudaEventRecord(e0);
            for (int w = 0; w < 3; ++w) ingest<<<G, 1024>>>(x, stride, ids, n, out, mode, clk, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
            long long hc[148]; cudaMemcpy(hc, clk, G * 8, cudaMemcpyDeviceToHost);
            double cyc = 0; for (int i = 0; i < G; ++i) cyc += hc[i]; cyc /= G;
            const double bytes = (double)n * 512;
            printf("mode %d  CTAs %3d: %7.1f GB/s gathered, SM clock %.0f MHz, %.1f B/clk/SM gathered%s\n", mode, G,
                   bytes / ms / 1e6, cyc / ms / 1e3, bytes / G / cyc, mode ? " (+20% stored)" : "");
        }
    return 0;
}
