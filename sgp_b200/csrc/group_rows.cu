// Host-side, one-off graph analysis for the RBU operator format (see khop_spmm.cu): decide which
// R rows share a register tile.  The reference has no counterpart (torch_sparse runs plain CSR);
// this is the B200-side answer to "the gathered panel does not stay in L1": rows that share most
// of their neighbours are processed together so every gathered source row is loaded once per
// group.  Sequential greedy, O(nnz):
//   1. breadth-first order of the rows over the stored adjacency (locality of the SCHEDULE: groups
//      that run concurrently touch a compact part of the source panel, which stays L2-resident);
//   2. walking that order, every still-unassigned row seeds a group and pulls in its R-1
//      still-unassigned neighbours with the largest operator weight (on kNN / kernel graphs:
//      the nearest ones), which are the rows most likely to share its neighbourhood;
//   3. rows left without enough free neighbours are chunked R at a time in BFS order.
#include <algorithm>
#include <vector>

#include "common.cuh"

extern "C" int sgp_group_rows(const int32_t* rowptr, const int32_t* col, const float* val, int32_t N,
                              int32_t R, int32_t* grp_rows, int32_t* n_groups_out) {
    using namespace sgp;
    SGP_REQUIRE(rowptr && grp_rows && n_groups_out, SGP_EINVAL, "sgp_group_rows: null pointer");
    SGP_REQUIRE(N >= 0 && R >= 1, SGP_EINVAL, "sgp_group_rows: N=%d R=%d", N, R);
    const int32_t n_groups = (N + R - 1) / R;
    *n_groups_out = n_groups;
    if (N == 0) return SGP_OK;
    SGP_REQUIRE(rowptr[N] == 0 || (col && val), SGP_EINVAL, "sgp_group_rows: null col/val");

    // 1. BFS order (all components)
    std::vector<int32_t> order;
    order.reserve(N);
    std::vector<uint8_t> seen(N, 0);
    for (int32_t s = 0; s < N; ++s) {
        if (seen[s]) continue;
        seen[s] = 1;
        size_t head = order.size();
        order.push_back(s);
        while (head < order.size()) {
            const int32_t i = order[head++];
            for (int32_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
                const int32_t j = col[e];
                if (j >= N) continue;     // rectangular operator (halo columns): not a row
                if (!seen[j]) { seen[j] = 1; order.push_back(j); }
            }
        }
    }

    // 2. greedy grouping
    std::vector<uint8_t> taken(N, 0);
    std::vector<int32_t> late;
    std::vector<std::pair<float, int32_t>> cand;
    int32_t g = 0;
    for (int32_t i : order) {
        if (taken[i]) continue;
        cand.clear();
        for (int32_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
            const int32_t j = col[e];
            if (j != i && j < N && !taken[j]) cand.emplace_back(-val[e], j);
        }
        std::sort(cand.begin(), cand.end());
        if ((int32_t)cand.size() < R - 1) { late.push_back(i); continue; }
        int32_t* out = grp_rows + (size_t)g * R;
        out[0] = i;
        taken[i] = 1;
        int32_t filled = 1;
        for (size_t k = 0; k < cand.size() && filled < R; ++k) {
            const int32_t j = cand[k].second;
            if (taken[j]) continue;   // duplicate edge to the same neighbour
            out[filled++] = j;
            taken[j] = 1;
        }
        if (filled < R) {             // duplicates shrank the candidate list: undo and defer
            for (int32_t k = 0; k < filled; ++k) taken[out[k]] = 0;
            late.push_back(i);
            continue;
        }
        ++g;
    }
    // 3. leftovers, chunked in BFS order
    int32_t slot = 0;
    for (int32_t i : late) {
        if (taken[i]) continue;
        grp_rows[(size_t)g * R + slot] = i;
        taken[i] = 1;
        if (++slot == R) { slot = 0; ++g; }
    }
    if (slot) {
        for (; slot < R; ++slot) grp_rows[(size_t)g * R + slot] = -1;
        ++g;
    }
    SGP_REQUIRE(g == n_groups, SGP_EINVAL, "sgp_group_rows: internal error, %d groups != %d", g, n_groups);
    return SGP_OK;
}
