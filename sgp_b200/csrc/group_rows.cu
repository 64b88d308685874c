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

// ---------------------------------------------------------------------------------------------
// Row partition for the row-sharded encoder (sgp_b200/sharded.py): `parts` compact, equally sized
// patches of the graph, by recursive bisection.  A patch is split across its long axis without
// coordinates: a, b = two mutually far nodes of the patch (breadth-first twice), every node gets
// the key d(a, .) - d(b, .) (hop distances inside the patch), and the patch is cut at the key's
// weighted median (ties: nearer to a first) — the graph analogue of cutting along the
// perpendicular bisector of its diameter.  Straight, short cuts keep the halo (the source rows a
// rank must receive per hop) near the geometric minimum; contiguous ranges of one global
// breadth-first order (round 1) gave annuli with 2x the boundary.  Host, O(log(parts) * sampled nnz).
// ---------------------------------------------------------------------------------------------
namespace {

struct PatchBfs {
    const int32_t* rowptr;
    const int32_t* col;
    int32_t N;
    std::vector<int32_t> queue;

    // hop distances from `start` inside the patch `id` (nodes with patch[node] == id); nodes the
    // sampled search cannot reach get far + 1.  Returns the last node reached.
    int32_t run(const std::vector<int32_t>& nodes, const std::vector<int32_t>& patch, int32_t id, int32_t start,
                std::vector<int32_t>& dist) {
        constexpr int32_t kFanout = 16;
        for (int32_t v : nodes) dist[v] = -1;
        queue.clear();
        queue.push_back(start);
        dist[start] = 0;
        size_t head = 0;
        int32_t far = 0;
        size_t scan = 0;
        for (;;) {
            while (head < queue.size()) {
                const int32_t i = queue[head++];
                far = dist[i];
                const int32_t deg = rowptr[i + 1] - rowptr[i];
                const int32_t stride = deg > kFanout ? deg / kFanout : 1;
                for (int32_t e = rowptr[i]; e < rowptr[i + 1]; e += stride) {
                    const int32_t j = col[e];
                    if (j < N && patch[j] == id && dist[j] < 0) {
                        dist[j] = dist[i] + 1;
                        queue.push_back(j);
                    }
                }
            }
            // unreached nodes of the patch (another component, or only reachable against the edge
            // direction): continue from the next one, one level further out
            while (scan < nodes.size() && dist[nodes[scan]] >= 0) ++scan;
            if (scan == nodes.size()) break;
            dist[nodes[scan]] = far + 1;
            queue.push_back(nodes[scan]);
        }
        return queue.back();
    }
};

}  // namespace

extern "C" int sgp_partition_rows(const int32_t* rowptr, const int32_t* col, int32_t N, int32_t parts,
                                  int32_t* owner) {
    using namespace sgp;
    SGP_REQUIRE(rowptr && owner, SGP_EINVAL, "sgp_partition_rows: null pointer");
    SGP_REQUIRE(N >= 0 && parts >= 1, SGP_EINVAL, "sgp_partition_rows: N=%d parts=%d", N, parts);
    if (N == 0) return SGP_OK;
    SGP_REQUIRE(rowptr[N] == 0 || col, SGP_EINVAL, "sgp_partition_rows: null col");
    std::vector<int32_t> patch(N, 0), da(N), db(N);
    PatchBfs bfs{rowptr, col, N, {}};
    bfs.queue.reserve(N);
    struct Job { std::vector<int32_t> nodes; int32_t parts, first, id; };
    std::vector<Job> stack;
    {
        Job j;
        j.nodes.resize(N);
        for (int32_t i = 0; i < N; ++i) j.nodes[i] = i;
        j.parts = parts; j.first = 0; j.id = 0;
        stack.push_back(std::move(j));
    }
    int32_t next_id = 1;
    std::vector<std::pair<float, int32_t>> keyed;
    std::vector<float> ka(N), kb(N);
    constexpr int kSmoothRounds = 2;
    while (!stack.empty()) {
        Job job = std::move(stack.back());
        stack.pop_back();
        if (job.parts == 1 || job.nodes.empty()) {
            for (int32_t v : job.nodes) owner[v] = job.first;
            continue;
        }
        const int32_t p0 = job.parts / 2, p1 = job.parts - p0;
        const size_t n0 = job.nodes.size() * (size_t)p0 / (size_t)job.parts;
        const int32_t a = bfs.run(job.nodes, patch, job.id, job.nodes[0], da);
        const int32_t b = bfs.run(job.nodes, patch, job.id, a, da);          // da = d(a, .)
        bfs.run(job.nodes, patch, job.id, b, db);                            // db = d(b, .)
        // Hop counts are integers with +-1 of noise, so the raw key leaves a band a few hops wide
        // in which the two sides interleave.  Two rounds of averaging the key over each node's
        // neighbours inside the patch turn it into a smooth field whose median level set is a clean
        // curve (each round averages ~deg noisy values).
        for (int32_t v : job.nodes) ka[v] = (float)(da[v] - db[v]);
        for (int round = 0; round < kSmoothRounds; ++round) {
            for (int32_t v : job.nodes) {
                float acc = ka[v];
                int32_t cnt = 1;
                for (int32_t e = rowptr[v]; e < rowptr[v + 1]; ++e) {
                    const int32_t j = col[e];
                    if (j < N && patch[j] == job.id) { acc += ka[j]; ++cnt; }
                }
                kb[v] = acc / (float)cnt;
            }
            for (int32_t v : job.nodes) ka[v] = kb[v];
        }
        keyed.clear();
        for (int32_t v : job.nodes) keyed.emplace_back(ka[v], v);
        std::nth_element(keyed.begin(), keyed.begin() + n0, keyed.end());
        Job left, right;
        left.nodes.reserve(n0);
        right.nodes.reserve(keyed.size() - n0);
        left.parts = p0; left.first = job.first; left.id = next_id++;
        right.parts = p1; right.first = job.first + p0; right.id = next_id++;
        for (size_t k = 0; k < keyed.size(); ++k) {
            const int32_t v = keyed[k].second;
            if (k < n0) { left.nodes.push_back(v); patch[v] = left.id; }
            else { right.nodes.push_back(v); patch[v] = right.id; }
        }
        // node order inside a patch only seeds the searches: keep it deterministic
        std::sort(left.nodes.begin(), left.nodes.end());
        std::sort(right.nodes.begin(), right.nodes.end());
        stack.push_back(std::move(right));
        stack.push_back(std::move(left));
    }
    return SGP_OK;
}
