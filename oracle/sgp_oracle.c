/*
 * CPU oracle, C part.  TEST INFRASTRUCTURE ONLY (see oracle/sgp_oracle.py for the rules on who
 * may load this).  Built by oracle/Makefile or oracle/sgp_oracle.py:build_c() into
 * oracle/_build/libsgp_oracle.so.
 *
 * oracle_spmm_csr_f32 restates the CPU kernel that executes the reference's `adj @ x`
 * (lib/sgp_preprocessing.py:202): torch_sparse `spmm_cpu` with reduce = sum.  torch_sparse is an
 * un-vendored dependency of the reference (rusty1s/pytorch_sparse ~0.6.12, pulled by pyg=2.0 in
 * conda_env.yml:8-10); its published algorithm is: parallel-for over the flattened
 * (batch, row) index space, and for each (b, m) a sequential float accumulation over the row's
 * stored entries e in [rowptr[m], rowptr[m+1]) of value[e] * mat[b, col[e], 0..K), int64 indices.
 * This is the loop below; OpenMP stands in for at::parallel_for.
 */
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_spmm_csr_f32(const int64_t* rowptr, const int64_t* col, const float* val,
                         int64_t B, int64_t M, int64_t F,
                         const float* x, int64_t x_bstride, int64_t x_rstride,
                         float* out, int64_t o_bstride, int64_t o_rstride) {
    const int64_t total = B * M;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < total; ++i) {
        const int64_t b = i / M, m = i % M;
        float* __restrict__ acc = out + b * o_bstride + m * o_rstride;
        const float* __restrict__ xb = x + b * x_bstride;
        memset(acc, 0, (size_t)F * sizeof(float));
        for (int64_t e = rowptr[m]; e < rowptr[m + 1]; ++e) {
            const float v = val[e];
            const float* __restrict__ src = xb + col[e] * x_rstride;
            for (int64_t f = 0; f < F; ++f) acc[f] += v * src[f];
        }
    }
}
