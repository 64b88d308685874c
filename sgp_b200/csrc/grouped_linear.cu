// Grouped 1x1 convolution = block-diagonal linear map over the encoder's feature blocks
// (SURVEY.md 8(f3)): the first decoder layer of SGPModel (lib/nn/models/sgp_model.py:41-52),
//     nn.Conv1d(in_channels = G * Cin, out_channels = G * Cout, kernel_size = 1, groups = G)
// applied to 'b n f -> b f n'.  Group g reads features [g Cin, (g+1) Cin) — exactly one (hop, layer)
// block of the encoder output — and writes outputs [g Cout, (g+1) Cout):
//     y[r, g Cout + o] = bias[g Cout + o] + sum_i x[r, g Cin + i] * w[g Cout + o, i]
// One CTA = 32 rows x one group; the x tile and the group's weight slice are staged in shared memory
// in k-chunks of 32, each thread owns one row and every 8th output column.  fp32 FFMA (the
// reference's CPU precision; the contraction is 2 * rows * Cin * Cout * G flops — small next to the
// encoder).
#include "common.cuh"

namespace sgp {

constexpr int kGlRows = 32, kGlKC = 32, kGlMaxAcc = 32;     // Cout <= 8 * 32 = 256 per group

__global__ void __launch_bounds__(256)
grouped_linear_kernel(const float* __restrict__ x, int64_t x_rs, const float* __restrict__ w,
                      const float* __restrict__ bias, float* __restrict__ y, int64_t y_rs,
                      int64_t rows, int Cin, int Cout) {
    extern __shared__ float smem[];
    float* xs = smem;                                   // [kGlRows][kGlKC + 1]
    float* ws = smem + kGlRows * (kGlKC + 1);           // [kGlKC][Cout + 1]
    const int g = blockIdx.y;
    const int64_t r0 = (int64_t)blockIdx.x * kGlRows;
    const int tr = threadIdx.x >> 3, tc = threadIdx.x & 7;      // row in tile, column lane
    const float* xg = x + (size_t)g * Cin;
    const float* wg = w + (size_t)g * Cout * Cin;
    float acc[kGlMaxAcc];
#pragma unroll
    for (int a = 0; a < kGlMaxAcc; ++a) acc[a] = 0.f;
    for (int k0 = 0; k0 < Cin; k0 += kGlKC) {
        const int kc = min(kGlKC, Cin - k0);
        for (int i = threadIdx.x; i < kGlRows * kGlKC; i += 256) {
            const int rr = i / kGlKC, kk = i % kGlKC;
            const int64_t r = r0 + rr;
            xs[rr * (kGlKC + 1) + kk] = (r < rows && kk < kc) ? __ldg(xg + (size_t)r * x_rs + k0 + kk) : 0.f;
        }
        for (int i = threadIdx.x; i < Cout * kGlKC; i += 256) {
            const int o = i / kGlKC, kk = i % kGlKC;
            ws[kk * (Cout + 1) + o] = (kk < kc) ? __ldg(wg + (size_t)o * Cin + k0 + kk) : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int kk = 0; kk < kGlKC; ++kk) {
            const float xv = xs[tr * (kGlKC + 1) + kk];
            const float* wr = ws + kk * (Cout + 1);
#pragma unroll
            for (int a = 0; a < kGlMaxAcc; ++a) {
                const int o = tc + 8 * a;
                if (o < Cout) acc[a] = fmaf(xv, wr[o], acc[a]);
            }
        }
        __syncthreads();
    }
    const int64_t r = r0 + tr;
    if (r < rows) {
#pragma unroll
        for (int a = 0; a < kGlMaxAcc; ++a) {
            const int o = tc + 8 * a;
            if (o < Cout) y[(size_t)r * y_rs + (size_t)g * Cout + o] = acc[a] + (bias ? __ldg(bias + (size_t)g * Cout + o) : 0.f);
        }
    }
}

}  // namespace sgp

using namespace sgp;

extern "C" int sgp_grouped_linear(const float* x, int64_t x_row_stride, const float* weight, const float* bias,
                                  float* y, int64_t y_row_stride, int64_t rows, int groups, int Cin, int Cout,
                                  void* stream) {
    SGP_REQUIRE(x && weight && y, SGP_EINVAL, "sgp_grouped_linear: null pointer");
    SGP_REQUIRE(rows >= 0 && groups >= 1 && Cin >= 1 && Cout >= 1, SGP_EINVAL,
                "sgp_grouped_linear: rows=%lld groups=%d Cin=%d Cout=%d", (long long)rows, groups, Cin, Cout);
    SGP_REQUIRE(Cout <= 8 * kGlMaxAcc, SGP_EUNSUPPORTED, "sgp_grouped_linear: Cout=%d per group (max %d)", Cout, 8 * kGlMaxAcc);
    SGP_REQUIRE(groups <= 65535, SGP_EUNSUPPORTED, "sgp_grouped_linear: groups=%d", groups);
    if (rows == 0) return SGP_OK;
    const size_t smem = ((size_t)kGlRows * (kGlKC + 1) + (size_t)kGlKC * (Cout + 1)) * sizeof(float);
    const int64_t gx = (rows + kGlRows - 1) / kGlRows;
    SGP_REQUIRE(gx < (1ll << 31), SGP_EUNSUPPORTED, "sgp_grouped_linear: too many rows for one launch");
    grouped_linear_kernel<<<dim3((unsigned)gx, groups), 256, smem, as_stream(stream)>>>(
        x, x_row_stride, weight, bias, y, y_row_stride, rows, Cin, Cout);
    SGP_LAUNCH_CHECK("grouped_linear");
    return SGP_OK;
}
