// K3 — edge list -> normalised CSR shift operator, on the device.
//
// Replaces preprocess_adj (lib/sgp_preprocessing.py:67-105: SparseTensor(row,col,value) ->
// set_diag / remove_diag -> row sums -> D^-1 S or D^-1/2 S D^-1/2) plus the two edge-list
// rewrites of sgp_spatial_embedding: to_undirected (:182-185, PyG: append reversed edges and
// coalesce duplicates with add) and edge_index[[1, 0]] for the bidirectional pass (:205-207).
// One-off O(E log E): key = row*N + col, stable LSD radix sort (cub) = torch_sparse's (row, col)
// order with duplicates kept in input order.
#include <cub/cub.cuh>

#include "common.cuh"

namespace sgp {

struct CsrCounters {
    int bad_index;       // some edge endpoint outside [0, N)
    int num_runs;        // ReduceByKey output
    int nnz;
    int pad;
};

__global__ void csr_emit_keys(const int64_t* __restrict__ esrc, const int64_t* __restrict__ edst,
                              const float* __restrict__ w, int64_t E, int32_t N, int flags,
                              uint64_t* __restrict__ keys, float* __restrict__ vals,
                              CsrCounters* ctr) {
    const uint64_t dropped = (uint64_t)N * (uint64_t)N;
    const bool sym = flags & SGP_CSR_SYMMETRIZE, tr = flags & SGP_CSR_TRANSPOSE;
    const bool no_diag = flags & (SGP_CSR_SET_DIAG | SGP_CSR_REMOVE_DIAG);
    const int64_t M = sym ? 2 * E : E;
    const int64_t P = M + ((flags & SGP_CSR_SET_DIAG) ? N : 0);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < P;
         i += (int64_t)gridDim.x * blockDim.x) {
        if (i >= M) {   // inserted unit diagonal
            const uint64_t d = (uint64_t)(i - M);
            keys[i] = d * N + d;
            vals[i] = 1.f;
            continue;
        }
        const int64_t e = i < E ? i : i - E;
        int64_t c = esrc[e], r = edst[e];     // ":80  col, row = edge_index"
        if (tr) { const int64_t s = c; c = r; r = s; }
        if (i >= E) { const int64_t s = c; c = r; r = s; }   // reversed copy (to_undirected)
        if (r < 0 || r >= N || c < 0 || c >= N) {
            ctr->bad_index = 1;
            keys[i] = dropped;
            vals[i] = 0.f;
            continue;
        }
        keys[i] = (no_diag && r == c) ? dropped : (uint64_t)r * N + (uint64_t)c;
        vals[i] = w ? w[e] : 1.f;
    }
}

// number of keys < limit in a sorted array (binary search by one thread; one-off)
__global__ void csr_count_valid(const uint64_t* __restrict__ keys, int64_t P, uint64_t limit,
                                const int* runs_or_null, CsrCounters* ctr) {
    int64_t n = runs_or_null ? (int64_t)*runs_or_null : P;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys[mid] < limit) lo = mid + 1; else hi = mid;
    }
    ctr->nnz = (int)lo;
}

__global__ void csr_finalize_structure(const uint64_t* __restrict__ keys, const CsrCounters* ctr,
                                       int32_t N, int32_t* __restrict__ rowptr,
                                       int32_t* __restrict__ col) {
    const int nnz = ctr->nnz;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e <= nnz;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int row = (e < nnz) ? (int)(keys[e] / (uint64_t)N) : N;
        const int prev = (e == 0) ? -1 : (int)(keys[e - 1] / (uint64_t)N);
        for (int r = prev + 1; r <= row; ++r) rowptr[r] = (int)e;
        if (e < nnz) col[e] = (int)(keys[e] % (uint64_t)N);
    }
}

// deg[i] = sum of row i (stored order); s[i] = deg^-1 or deg^-1/2 with inf -> 0
__global__ void csr_degree_scale(const int32_t* __restrict__ rowptr, const float* __restrict__ w,
                                 int32_t N, int gcn, int unit_w, int no_norm, float* __restrict__ s) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        if (no_norm) { s[i] = 1.f; continue; }
        float d = 0.f;
        for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) d += unit_w ? 1.f : w[e];
        float v = gcn ? (1.0f / sqrtf(d)) : (1.0f / d);
        if (isinf(v)) v = 0.f;
        s[i] = v;
    }
}

__global__ void csr_apply_scale(const uint64_t* __restrict__ keys, const float* __restrict__ w,
                                const float* __restrict__ s, const CsrCounters* ctr, int32_t N,
                                int gcn, int unit_w, float* __restrict__ val) {
    const int nnz = ctr->nnz;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz;
         e += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[e];
        const int r = (int)(k / (uint64_t)N), c = (int)(k % (uint64_t)N);
        float v = s[r] * (unit_w ? 1.f : w[e]);
        if (gcn) v = v * s[c];
        val[e] = v;
    }
}

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

struct CsrLayout {
    int64_t P;
    size_t keys_a, keys_b, vals_a, vals_b, scale, ctr, cub, total, cub_bytes;
};

static int end_bit_for(int32_t N) {
    const unsigned long long top = (unsigned long long)N * (unsigned long long)N;   // dropped key
    int b = 1;
    while (b < 64 && (top >> b) != 0) ++b;
    return b;
}

static CsrLayout csr_layout(int64_t E, int32_t N, int flags) {
    CsrLayout L{};
    const int64_t M = (flags & SGP_CSR_SYMMETRIZE) ? 2 * E : E;
    L.P = M + ((flags & SGP_CSR_SET_DIAG) ? N : 0);
    const size_t P = (size_t)(L.P > 0 ? L.P : 1);
    size_t sort_b = 0, red_b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_b, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                    (const float*)nullptr, (float*)nullptr, (int)P, 0, end_bit_for(N));
    cub::DeviceReduce::ReduceByKey(nullptr, red_b, (const uint64_t*)nullptr, (uint64_t*)nullptr,
                                   (const float*)nullptr, (float*)nullptr, (int*)nullptr, cub::Sum(),
                                   (int)P);
    L.cub_bytes = sort_b > red_b ? sort_b : red_b;
    size_t off = 0;
    L.keys_a = off; off += align_up(P * sizeof(uint64_t));
    L.keys_b = off; off += align_up(P * sizeof(uint64_t));
    L.vals_a = off; off += align_up(P * sizeof(float));
    L.vals_b = off; off += align_up(P * sizeof(float));
    L.scale = off; off += align_up((size_t)(N > 0 ? N : 1) * sizeof(float));
    L.ctr = off; off += align_up(sizeof(CsrCounters));
    L.cub = off; off += align_up(L.cub_bytes);
    L.total = off;
    return L;
}

}  // namespace sgp

using namespace sgp;

extern "C" size_t sgp_csr_build_workspace_bytes(int64_t E, int32_t N, int flags) {
    if (E < 0 || N < 0) return 0;
    return csr_layout(E, N, flags).total;
}

extern "C" int sgp_csr_build(const int64_t* edge_src, const int64_t* edge_dst, const float* weight,
                             int64_t E, int32_t N, int flags, int32_t* rowptr, int32_t* col,
                             float* val, int64_t cap, int64_t* nnz_out, void* workspace,
                             size_t workspace_bytes, void* stream) {
    SGP_REQUIRE(E >= 0 && N >= 0, SGP_EINVAL, "sgp_csr_build: E=%lld N=%d", (long long)E, N);
    SGP_REQUIRE(rowptr && nnz_out && workspace, SGP_EINVAL, "sgp_csr_build: null pointer");
    SGP_REQUIRE(E == 0 || (edge_src && edge_dst), SGP_EINVAL, "sgp_csr_build: null edge list");
    const CsrLayout L = csr_layout(E, N, flags);
    SGP_REQUIRE(L.P < (1ll << 31) - 1, SGP_EUNSUPPORTED, "sgp_csr_build: %lld entries exceed int32",
                (long long)L.P);
    SGP_REQUIRE(workspace_bytes >= L.total, SGP_ECAPACITY,
                "sgp_csr_build: workspace %zu B < required %zu B", workspace_bytes, L.total);
    SGP_REQUIRE(cap >= L.P, SGP_ECAPACITY, "sgp_csr_build: col/val capacity %lld < %lld",
                (long long)cap, (long long)L.P);
    SGP_REQUIRE(L.P == 0 || (col && val), SGP_EINVAL, "sgp_csr_build: null col/val");
    cudaStream_t st = as_stream(stream);
    char* ws = static_cast<char*>(workspace);
    uint64_t* keys_a = reinterpret_cast<uint64_t*>(ws + L.keys_a);
    uint64_t* keys_b = reinterpret_cast<uint64_t*>(ws + L.keys_b);
    float* vals_a = reinterpret_cast<float*>(ws + L.vals_a);
    float* vals_b = reinterpret_cast<float*>(ws + L.vals_b);
    float* scale = reinterpret_cast<float*>(ws + L.scale);
    CsrCounters* ctr = reinterpret_cast<CsrCounters*>(ws + L.ctr);
    void* cub_ws = ws + L.cub;
    size_t cub_bytes = L.cub_bytes;
    const int P = (int)L.P;
    const uint64_t dropped = (uint64_t)N * (uint64_t)N;
    const int threads = 256;
    auto grid_for = [&](int64_t n) { int64_t g = (n + threads - 1) / threads; return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g)); };

    SGP_CUDA(cudaMemsetAsync(ctr, 0, sizeof(CsrCounters), st));
    const uint64_t* keys = keys_b;
    const float* w = vals_b;
    if (P > 0) {
        csr_emit_keys<<<grid_for(P), threads, 0, st>>>(edge_src, edge_dst, weight, E, N, flags, keys_a, vals_a, ctr);
        SGP_LAUNCH_CHECK("csr_emit_keys");
        SGP_CUDA(cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, keys_a, keys_b, vals_a, vals_b, P, 0,
                                                 end_bit_for(N), st));
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (flags & SGP_CSR_SYMMETRIZE) {
            cub_bytes = L.cub_bytes;
            SGP_CUDA(cub::DeviceReduce::ReduceByKey(cub_ws, cub_bytes, keys_b, keys_a, vals_b, vals_a,
                                                    &ctr->num_runs, cub::Sum(), P, st));
            g_launches.fetch_add(1, std::memory_order_relaxed);
            keys = keys_a;
            w = vals_a;
            csr_count_valid<<<1, 1, 0, st>>>(keys, P, dropped, &ctr->num_runs, ctr);
        } else {
            csr_count_valid<<<1, 1, 0, st>>>(keys, P, dropped, nullptr, ctr);
        }
        SGP_LAUNCH_CHECK("csr_count_valid");
    }
    csr_finalize_structure<<<grid_for((int64_t)P + 1), threads, 0, st>>>(keys, ctr, N, rowptr, col);
    SGP_LAUNCH_CHECK("csr_finalize_structure");
    if (P > 0) {
        const int gcn = (flags & SGP_CSR_GCN_NORM) ? 1 : 0;
        // to_undirected without edge weights only de-duplicates: every distinct edge counts once
        const int unit_w = (!weight && (flags & SGP_CSR_SYMMETRIZE)) ? 1 : 0;
        const int no_norm = (flags & SGP_CSR_NO_NORM) ? 1 : 0;      // values stay as given (s = 1)
        csr_degree_scale<<<grid_for(N), threads, 0, st>>>(rowptr, w, N, gcn && !no_norm, unit_w, no_norm, scale);
        SGP_LAUNCH_CHECK("csr_degree_scale");
        csr_apply_scale<<<grid_for(P), threads, 0, st>>>(keys, w, scale, ctr, N, gcn && !no_norm, unit_w, val);
        SGP_LAUNCH_CHECK("csr_apply_scale");
    }
    CsrCounters host{};
    SGP_CUDA(cudaMemcpyAsync(&host, ctr, sizeof(host), cudaMemcpyDeviceToHost, st));
    SGP_CUDA(cudaStreamSynchronize(st));
    SGP_REQUIRE(!host.bad_index, SGP_EINVAL, "sgp_csr_build: edge index outside [0, %d)", N);
    *nnz_out = host.nnz;
    return SGP_OK;
}
