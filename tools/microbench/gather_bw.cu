// Random 512-byte row-piece gather bandwidth (the access pattern of the hop SpMM's producers).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bw gather_bw.cu && ./gather_bw
// rows: R rows of `stride` floats; each warp gathers 512-byte pieces (one per row id) with LDG.128,
// 8 rows in flight per warp.  `panels` = how many 128-float column panels of each row are visited
// (footprint = R * 512 B * panels); ids are random (scattered) or sorted (clustered).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

__global__ void gather(const float* __restrict__ x, int stride, const int* __restrict__ ids, long n_ids, int panels, float* sink) {
    const int lane = threadIdx.x & 31;
    const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * (long)blockDim.x) >> 5;
    float acc = 0.f;
    for (long base = warp * 32; base < n_ids; base += n_warps * 32) {
        const int my = ids[base + lane];
        for (int p = 0; p < panels; ++p) {
#pragma unroll
            for (int j0 = 0; j0 < 32; j0 += 8) {
                float4 v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int r = __shfl_sync(0xffffffffu, my, j0 + j);
                    v[j] = __ldcg(reinterpret_cast<const float4*>(x + (size_t)r * stride + p * 128) + lane);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) acc += v[j].x + v[j].w;
            }
        }
    }
    if (acc == 123.456f) *sink = acc;
}

int main() {
    const int stride = 1280;
    const int Rmax = 200000;
    float* x; cudaMalloc(&x, (size_t)Rmax * stride * 4); cudaMemset(x, 0, (size_t)Rmax * stride * 4);
    float* sink; cudaMalloc(&sink, 4);
    const long n_ids = 148L * 64 * 32 * 16;      // 4.8M row pieces = 2.5 GB of traffic per panel
    int* ids; cudaMalloc(&ids, n_ids * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int Rs[] = {2000, 20000, 100000, 200000};
    for (int mode = 0; mode < 2; ++mode)
        for (int R : Rs)
            for (int panels : {1, 4}) {
                std::vector<int> h(n_ids);
                srand(1);
                for (long i = 0; i < n_ids; ++i) h[i] = (int)(((long)rand() * 32768 + rand()) % R);
                if (mode == 1)   // clustered: every 32-id chunk comes from a window of 512 consecutive rows
                    for (long i = 0; i < n_ids; i += 32) {
                        const int w0 = h[i] % std::max(R - 512, 1);
                        for (int j = 0; j < 32; ++j) h[i + j] = w0 + h[i + j] % 512;
                    }
                cudaMemcpy(ids, h.data(), n_ids * 4, cudaMemcpyHostToDevice);
                for (int w = 0; w < 2; ++w) gather<<<148 * 8, 256>>>(x, stride, ids, n_ids, panels, sink);
                cudaEventRecord(e0);
                for (int w = 0; w < 3; ++w) gather<<<148 * 8, 256>>>(x, stride, ids, n_ids, panels, sink);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
                printf("%s rows %6d span %5.0f MB footprint %6.1f MB: %7.1f GB/s\n", mode ? "clustered" : "scattered", R,
                       (double)R * stride * 4 / 1e6, (double)R * 512 * panels / 1e6, (double)n_ids * 512 * panels / ms / 1e6);
            }
    return 0;
}
