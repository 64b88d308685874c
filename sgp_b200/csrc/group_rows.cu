// Host-side, one-off graph analysis for the RBU operator format (see khop_spmm.cu): decide which
// R rows share a register tile.  The reference has no counterpart (torch_sparse runs plain CSR);
// this is the B200-side answer to "the gathered panel does not stay in L1": rows that share most
// of their neighbours are processed together so every gathered source row is loaded once per
// group.  Sequential greedy, O(nnz):
//   1. breadth-first order of the rows over the stored adjacency (locality of the SCHEDULE: groups
//      that run concurrently touch a compact part of the source panel, which stays L2-resident);
//   2. walking that order, every still-unassigned row seeds a group and pulls in its R-1
//      still-unassigned neighbours with the largest operator weight (on kNN / kernel graphs:
//      the nearest ones), which are the rows most likely to share its neighbourhood; a seed with
//      too few free neighbours continues breadth-first over the group's own members.
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
    constexpr int32_t kFanout = 12;
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
            // a strided sample of the row's list is enough to carry the front forward (the order
            // only has to be spatially coherent) and keeps this pass at O(kFanout N), not O(nnz):
            // it runs inside the end-to-end timed region
            const int32_t deg = rowptr[i + 1] - rowptr[i];
            const int32_t stride = deg > kFanout ? deg / kFanout : 1;
            for (int32_t e = rowptr[i]; e < rowptr[i + 1]; e += stride) {
                const int32_t j = col[e];
                if (j >= N) continue;     // rectangular operator (halo columns): not a row
                if (!seen[j]) { seen[j] = 1; order.push_back(j); }
            }
        }
    }

    // 2. greedy grouping.  Every group is a compact blob: the seed's free neighbours by weight,
    //    then — when the seed sits at the edge of what is already grouped and has too few — the
    //    free neighbours of the members chosen so far (breadth-first over the blob), and only when
    //    nothing free is reachable any more the next free row in BFS order.  (Deferring such seeds
    //    and chunking them at the end left 3% of the rows in ~45 groups of scattered "holes" whose
    //    column unions were 10x the median: 20% of all gathers and MMAs.)
    std::vector<uint8_t> taken(N, 0);
    std::vector<std::pair<float, int32_t>> cand;
    size_t next_free = 0;             // scan position in `order` for the jump fallback
    int32_t g = 0;
    for (int32_t i : order) {
        if (taken[i]) continue;
        int32_t* out = grp_rows + (size_t)g * R;
        out[0] = i;
        taken[i] = 1;
        int32_t filled = 1;
        cand.clear();
        for (int32_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
            const int32_t j = col[e];
            if (j != i && j < N && !taken[j]) cand.emplace_back(-val[e], j);
        }
        std::sort(cand.begin(), cand.end());
        for (size_t k = 0; k < cand.size() && filled < R; ++k) {
            const int32_t j = cand[k].second;
            if (taken[j]) continue;   // duplicate edge to the same neighbour
            out[filled++] = j;
            taken[j] = 1;
        }
        int32_t expand = 1;           // members[expand..) have not been expanded yet
        while (filled < R) {
            if (expand < filled) {
                const int32_t m = out[expand++];
                for (int32_t e = rowptr[m]; e < rowptr[m + 1] && filled < R; ++e) {
                    const int32_t j = col[e];
                    if (j < N && !taken[j]) { out[filled++] = j; taken[j] = 1; }
                }
            } else {
                while (next_free < order.size() && taken[order[next_free]]) ++next_free;
                if (next_free == order.size()) break;
                const int32_t j = order[next_free];
                out[filled++] = j;
                taken[j] = 1;
            }
        }
        for (; filled < R; ++filled) out[filled] = -1;      // only the last group can be short
        ++g;
    }
    SGP_REQUIRE(g == n_groups, SGP_EINVAL, "sgp_group_rows: internal error, %d groups != %d", g, n_groups);
    return SGP_OK;
}
