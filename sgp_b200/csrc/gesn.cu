// DynGESN state update: the graph echo-state layer of
// lib/nn/reservoir/graph_reservoir.py:85-93
//     h' = (1 - alpha) h + alpha * act( x W_ih^T + b + S (h W_hh^T) )
// Per time step the dense product G = h W_hh^T runs in the K1 scan kernel (identity activation, one
// step), P = S G in the K2 SpMM kernels, and this kernel fuses the rest: input projection, bias, the
// propagated term, activation and the leaky blend, one warp per node row (the self-normalising
// activation needs the row's 2-norm).  HBM-bound: reads P, h, x once, writes h' (state) and the
// output row.
#include "common.cuh"

namespace sgp {

template <int ACT>
__global__ void __launch_bounds__(256)
gesn_update_kernel(const float* __restrict__ x, int64_t x_ns, int Fin, const float* __restrict__ w_ih,
                   const float* __restrict__ bias, const float* __restrict__ prop, int64_t p_ns,
                   float* __restrict__ h, float* __restrict__ out, int64_t o_ns, float alpha, float oma,
                   int N, int H) {
    const int lane = threadIdx.x & 31;
    const int row = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    if (row >= N) return;
    const float* xr = x + (size_t)row * x_ns;
    const float* pr = prop + (size_t)row * p_ns;
    float* hr = h + (size_t)row * H;
    float* orow = out + (size_t)row * o_ns;
    float nrm = 0.f;
    for (int j = lane; j < H; j += 32) {
        float z = pr[j] + (bias ? __ldg(bias + j) : 0.f);
        for (int f = 0; f < Fin; ++f) z = fmaf(__ldg(xr + f), __ldg(w_ih + (size_t)j * Fin + f), z);
        if (ACT == SGP_ACT_TANH) z = tanhf(z);
        else if (ACT == SGP_ACT_RELU) z = fmaxf(z, 0.f);
        if (ACT == SGP_ACT_SELF_NORM) {
            nrm = fmaf(z, z, nrm);
            orow[j] = z;                                   // parked until the norm is known
        } else {
            const float hv = fmaf(alpha, z, oma * hr[j]);
            hr[j] = hv;
            orow[j] = hv;
        }
    }
    if (ACT == SGP_ACT_SELF_NORM) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
        const float inv = 1.f / fmaxf(sqrtf(nrm), 1e-12f);   // F.normalize(p=2, eps=1e-12), r = 1
        for (int j = lane; j < H; j += 32) {
            const float hv = fmaf(alpha, orow[j] * inv, oma * hr[j]);
            hr[j] = hv;
            orow[j] = hv;
        }
    }
}

}  // namespace sgp

using namespace sgp;

extern "C" int sgp_gesn_update(const float* x, int64_t x_n_stride, int Fin, const float* w_ih, const float* bias,
                               const float* prop, int64_t prop_n_stride, float alpha, float one_minus_alpha,
                               int act, float* h_state, float* out, int64_t out_n_stride, int N, int H,
                               void* stream) {
    SGP_REQUIRE(x && w_ih && prop && h_state && out, SGP_EINVAL, "sgp_gesn_update: null pointer");
    SGP_REQUIRE(N >= 0 && H >= 1 && Fin >= 1, SGP_EINVAL, "sgp_gesn_update: N=%d H=%d Fin=%d", N, H, Fin);
    SGP_REQUIRE(act >= SGP_ACT_TANH && act <= SGP_ACT_IDENTITY, SGP_EINVAL, "sgp_gesn_update: activation %d", act);
    if (N == 0) return SGP_OK;
    const int grid = (int)(((int64_t)N * 32 + 255) / 256);
#define SGP_GESN(A_)                                                                                  \
    gesn_update_kernel<A_><<<grid, 256, 0, as_stream(stream)>>>(x, x_n_stride, Fin, w_ih, bias, prop,  \
                                                                  prop_n_stride, h_state, out, out_n_stride, \
                                                                  alpha, one_minus_alpha, N, H)
    if (act == SGP_ACT_TANH) SGP_GESN(SGP_ACT_TANH);
    else if (act == SGP_ACT_RELU) SGP_GESN(SGP_ACT_RELU);
    else if (act == SGP_ACT_SELF_NORM) SGP_GESN(SGP_ACT_SELF_NORM);
    else SGP_GESN(SGP_ACT_IDENTITY);
#undef SGP_GESN
    SGP_LAUNCH_CHECK("gesn_update");
    return SGP_OK;
}
