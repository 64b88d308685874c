// K2-TC — the hop SpMM on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-accurate.
//
// Same contract as sgp_spmm_rbu (dst[t, i, :] = sum_e val[e] * src[t, col[e], :], replacing
// `x = adj @ x` of lib/sgp_preprocessing.py:200-203), different formulation: rows are grouped 64
// at a time (compact blobs, sgp_group_rows) and a hop becomes, per group,
//     D[f, r] = sum_u  X[u, f] * B[r, u]        f: 128 features, r: 64 rows, u: union columns
// i.e. a dense [128 x U] x [U x 64] GEMM whose A operand is the GATHERED source rows and whose B
// operand is the group's slab of operator values (zero where a row lacks the column).  The zero
// fill that costs the CUDA-core RBU kernel 1.8x of its FFMA2 budget is free here, while the
// gather traffic drops another 2x (U/R = 4.96 at R = 64 against 11.2 at R = 16 on the 100-NN
// graph).
//
//   * A = X^T chunk (32 gathered rows x 128 features), fed to the MMA from TMEM (TS mode: lane =
//     feature, column = gathered row).  The gather is cp.async (16 bytes per lane, a warp per 512-byte
//     row piece) into a row-major ring stage; the copies signal the stage's mbarrier themselves
//     (cp.async.mbarrier.arrive.noinc), so the whole 8-stage ring stays in flight.
//   * B = slab chunk (64 rows x 32 columns), K-major SWIZZLE_128B, stored pre-swizzled as fp32 (8 KB)
//     at operator build time: one TMA bulk copy per chunk, split into tf32 hi | lo images by a
//     dedicated warp, reused by the chunk's 4 items (2 time steps x 2 feature chunks at F = 256).
//   * precision: 3xTF32.  hi = x with the low 13 mantissa bits cleared, lo = x - hi, and
//     D += Ah*Bh + Al*Bh + Ah*Bl with fp32 accumulation in TMEM.  The tensor core ignores the low
//     13 bits of a tf32 operand, so the gathered fp32 rows are used as Ah as they are and only Al
//     is computed (in registers, on the way from the ring stage to TMEM); B is split at operator
//     build time.  Measured 1e-6..3e-6 relative (tools/microbench/umma_tf32_test.cu).
//   * TMEM: 4 accumulators of [128 lanes x 64 columns] + 4 A tiles of (32 hi + 32 lo) columns =
//     all 512 columns, one persistent CTA per SM.
//   * 23 warps, all hand-offs through mbarriers: 4 producer warps (one per accumulator index), 16
//     "split" warps (4 groups x TMEM lane quarter: ring stage -> hi | lo -> tcgen05.st, and the
//     accumulator drain), 2 MMA-issuing warps (one elected thread each, 12 tcgen05.mma per item),
//     1 slab warp.
//     Lessons that shaped it (profiles/r1_trace_tc.txt): a lone warp issues dependent instructions
//     ~5 cycles apart, so every role's per-item loop must be a few dozen instructions (barrier
//     addresses precomputed, nothing recomputed per item, no trace code compiled in); `lane == 0`
//     makes ptxas wrap every MMA in an R2UR waterfall (elect.sync does not); wait_group-based
//     signalling keeps only `lag` stages in flight; a TMA bulk copy per 512-byte row costs ~75
//     cycles per request; 128 threads arriving/polling on one mbarrier serialise.
// Bound: HBM (algorithmic bytes per hop-panel 8 nnz / Tb + 4(N+1) / Tb + 8 N F: 210 MB at C4, Tb =
// 16); today the tensor pipe (384 cycles per item) plus the exposed accumulator drain.
#include <stdlib.h>

#include "common.cuh"

namespace sgp {

constexpr int kTcR = 64;            // rows per group  (MMA N)
constexpr int kTcKC = 32;           // union columns per chunk
constexpr int kTcAcc = 4;           // accumulators per CTA (time steps x feature chunks)
constexpr int kTcABufs = 4;         // A (hi | lo) tiles resident in TMEM
#ifndef SGP_TC_PRODUCER_WARPS
#define SGP_TC_PRODUCER_WARPS 4
#endif
#ifndef SGP_TC_SPLIT_GROUPS
#define SGP_TC_SPLIT_GROUPS 4
#endif
constexpr int kTcSplitGroups = SGP_TC_SPLIT_GROUPS;   // split groups take items round-robin (group = accumulator index at 4)
constexpr int kTcSplitWarps = 4 * kTcSplitGroups;   // lo-pass + epilogue warps (warp & 3 = TMEM lane quarter)
constexpr int kTcIssuers = 2;       // MMA-issuing warps (one elected thread each)
constexpr int kTcPadWarps = 1;      // the slab warp
constexpr int kTcStages = 8;        // gathered-row ring in shared memory
constexpr int kTcStageBytes = kTcKC * 128 * 4;      // 16 KB: 32 rows x 128 features, row-major
constexpr int kTcBBytes = 2 * kTcR * kTcKC * 4;     // 16 KB: hi + lo image of one chunk
constexpr int kTcBBufs = 3;         // slab-image buffers (3: a producer may only wait on MMAs >= 3 chunks old,
                                    // anything newer can depend on items it has not signalled yet)
constexpr int kTcBRawBytes = kTcR * kTcKC * 4;      // 8 KB: the chunk's fp32 slab image as stored in the operator
constexpr size_t kTcSmem = (size_t)kTcStages * kTcStageBytes + kTcBBufs * (kTcBBytes + kTcBRawBytes) + 1024;
// TMEM columns: [0, 256) four fp32 accumulators of 64 columns, [256, 512) four A tiles of
// (32 hi + 32 lo) columns
constexpr int kTcTmemCols = 512;
constexpr int kTcAOff = kTcAcc * kTcR;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// L2 eviction policies: slab images and the hop's output stream through L2 once (evict_first) so
// that they do not push out the gathered panel rows, which neighbouring groups re-read (the
// cross-ring reuse distance of the breadth-first group order is about one wave of CTAs).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// elect.sync: exactly one lane of a converged warp.  ptxas knows a single lane is active under this
// predicate and moves MMA operands to uniform registers directly (under `lane == 0` it emits a
// per-operand ELECT / R2UR.BROADCAST / branch waterfall: ~13 instructions per tcgen05.mma).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(smem_u32(bar)) : "memory");
}

// Bounded warp-wide wait (all 32 lanes poll the same word: one broadcast shared-memory access per
// try).  A barrier that never completes raises the error flag and the CTA-wide abort flag (so that
// every other role stops at once) instead of hanging the GPU.
__device__ __forceinline__ bool warp_wait(uint64_t* bar, uint32_t parity, volatile int* abort_s, int* err, int lane) {
    const uint32_t a = smem_u32(bar);
    uint32_t done = 0;
#pragma unroll 1
    for (int it = 0; it < (1 << 24); ++it) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return true;
        if ((it & 63) == 63 && *abort_s) return false;
    }
    if (lane == 0) {
        *abort_s = 1;
        atomicExch(err, 1);
    }
    return false;
}

#define SGP_TMEM_ST16(addr, arr)                                                                   \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
                 :: "r"(addr), "r"(arr[0]), "r"(arr[1]), "r"(arr[2]), "r"(arr[3]), "r"(arr[4]), "r"(arr[5]),  \
                    "r"(arr[6]), "r"(arr[7]), "r"(arr[8]), "r"(arr[9]), "r"(arr[10]), "r"(arr[11]),            \
                    "r"(arr[12]), "r"(arr[13]), "r"(arr[14]), "r"(arr[15]) : "memory")
// Warp roles (23 warps; warp 22 splits the slab images): warps 0-15 "split" in four groups of four (group = accumulator index,
// warp & 3 = the TMEM lane quarter the warp may touch), warps 16-19 "producer" (cp.async gathers;
// producer 0 also fetches the slab images), warps 20-21 issue the MMAs (items a = q mod 2).
// An item is (chunk c, accumulator a); i = 4c + a; ring stage s = i % 8; A tile = a.
// mbarriers: full[s]  producer (its cp.asyncs) -> split : stage s holds item i's gathered rows
//            empty[s] split -> producers  : the split group has copied stage s into TMEM
//            ready[a] split -> MMA        : A tile a (hi | lo) is in TMEM
//            afree[a] MMA (tcgen05.commit) -> split : the MMAs reading A tile a have completed
//            braw[c%3]  producer 0 (TMA bytes) -> slab warp : the chunk's fp32 slab image has landed
//            bfull[c%3] slab warp -> MMA : the chunk's hi | lo slab images are written
//            bfree[c%3] MMA (tcgen05.commit, both issuers) -> producer 0 : slab buffer may be overwritten
//            done     MMA (both issuers) -> split : the work item's accumulators are complete
//            accfree[a] split (16 warps) -> MMA : accumulator a has been read out
// BOTH MMA operands' A side lives in TMEM: the split warps read a gathered stage once from shared
// memory (thread = feature = TMEM lane, 32 k values), form hi / lo in registers and tcgen05.st
// them; the 12 MMAs of an item then fetch only the 2 KB slab operand from shared memory each, so
// shared-memory bandwidth (the limit of the all-in-smem version: 120 KB per item) drops to 56 KB.
template <int NFC, bool HALO, int kTcProducerWarps>
__global__ void __launch_bounds__((kTcSplitWarps + kTcProducerWarps + kTcIssuers + kTcPadWarps) * 32, 1)
spmm_rbu_tc_kernel(const int32_t* __restrict__ chunk_ptr, const int32_t* __restrict__ grp_rows,
                   const int32_t* __restrict__ cols, const float* __restrict__ bimg,
                   int n_groups, int n_work, int group_major, int gather_policy,
                   const float* __restrict__ src, int64_t s_ts, uint32_t s_nb /* row stride, BYTES */,
                   const float* __restrict__ src2, int64_t s2_ts, uint32_t s2_nb, int n_split,
                   float* __restrict__ dst, int64_t d_ts, int64_t d_ns, int Tc, int* err, double* __restrict__ chk,
                   long long* trace) {
    static_assert(kTcAcc == 4 && kTcABufs == 4 && kTcStages == 8, "index arithmetic below");
    constexpr int TB = kTcAcc / NFC;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // shared-window address
    __shared__ uint64_t full[kTcStages], empty[kTcStages], ready[kTcABufs], afree[kTcABufs], bfree[kTcBBufs], bfull[kTcBBufs], braw[kTcBBufs];
    __shared__ uint64_t done, accfree[kTcAcc];
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int abort_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // optional per-item timestamps of CTA 7 (tools/trace_tc.py); trace == nullptr in production
    // (compiled in only with -DSGP_TC_TRACE: even predicated-off, the stamps cost a lone warp ~75 cycles per item)
#ifdef SGP_TC_TRACE
    const bool tr = trace && blockIdx.x == 7;
#define SGP_TRACE(role, i) do { if (tr && lane == 0 && (i) < 512) trace[(role) * 512 + (i)] = clock64(); } while (0)
#define SGP_TRACE1(role, i) do { if (tr) trace[(role) * 512 + min((i), 511)] = clock64(); } while (0)
#else
#define SGP_TRACE(role, i) do { } while (0)
#define SGP_TRACE1(role, i) do { } while (0)
#endif

    if (tid == 0) {
        abort_s = 0;
        for (int s = 0; s < kTcStages; ++s) {
            mbar_init(&full[s], 32);             // the 32 lanes' cp.async arrivals
            mbar_init(&empty[s], 4);
        }
        for (int b = 0; b < kTcABufs; ++b) {
            mbar_init(&ready[b], 4);
            mbar_init(&afree[b], 1);
            mbar_init(&accfree[b], kTcSplitWarps);      // every split warp drains a slice of every accumulator
        }
        for (int b = 0; b < kTcBBufs; ++b) {
            mbar_init(&bfree[b], kTcIssuers);
            mbar_init(&bfull[b], 1);           // the slab warp: hi | lo images written
            mbar_init(&braw[b], 1);            // producer 0's arrive.expect_tx; the bulk copy completes the bytes
        }
        mbar_init(&done, kTcIssuers);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&tmem_base_s)), "r"(kTcTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    constexpr uint32_t kBOff = kTcStages * kTcStageBytes;               // slab images (hi | lo) after the ring
    constexpr uint32_t kBRawOff = kBOff + kTcBBufs * kTcBBytes;         // then the raw fp32 slab images
    // Work order of this (persistent) CTA, s-th work item:
    //   group_major == 0 (default): w = blockIdx.x + s * gridDim.x, (time block, group) =
    //     (w / n_groups, w % n_groups): all CTAs sweep the groups of one time block together, so
    //     neighbouring groups' gathers share panel rows in L2 for as long as the sweep lasts;
    //   group_major == 1: the CTA's groups are blockIdx.x, + gridDim.x, ... and each runs ALL time
    //     blocks back to back (slab images stay in L2) — measured slower (133 vs 104 us per
    //     hop-panel at C4): CTAs drift apart by whole work items and then share nothing.
    const int ny = n_work / n_groups;
    auto work_item = [&](int s_, int& g_, int& t_begin_) -> bool {
        if (group_major) {
            const int k_ = s_ / ny;
            g_ = blockIdx.x + k_ * gridDim.x;
            t_begin_ = (s_ - k_ * ny) * TB;
            return g_ < n_groups;
        }
        const int w_ = blockIdx.x + s_ * gridDim.x;
        const int y_ = w_ / n_groups;
        g_ = w_ - y_ * n_groups;
        t_begin_ = y_ * TB;
        return w_ < n_work;
    };

    // The CTA is persistent.  All barrier phases run on counters that continue across work
    // items: `it` items, `cc` chunks, (bi, bph) slab buffers, `wn` work items with at least one
    // chunk.  The producers simply run on into the next work item while the split warps drain the
    // accumulators of the previous one.
    if (warp >= kTcSplitWarps && warp < kTcSplitWarps + kTcProducerWarps) {
        // ================= producers: warp p gathers the items with accumulator index a = p ======
        // One item = 32 source rows x 512 bytes: a cp.async per row, the 32 lanes covering its 128
        // features, landing row-major in ring stage it % 8.  Completion is signalled by the copies
        // themselves (cp.async.mbarrier.arrive.noinc on full[s]), so a producer waits for nothing
        // but a free stage and the ring stays full; with commit/wait_group signalling only ~5
        // items were in flight at 4300 cycles of loaded gather latency.  Whole items per warp (not
        // rows of every item) because a lone warp issues ~1 instruction per 5 cycles: the per-item
        // overhead (waits, barrier addresses) is paid once per chunk and warp, not four times.
        // (One 512-byte TMA bulk copy per row was tried: ~75 cycles per request, 2400 per item.)
        // The chunk's slab images ride on item a = 0's barrier as one 16 KB TMA bulk copy.
        static_assert(kTcProducerWarps == kTcAcc, "one producer warp per accumulator index");
        const int pw = warp - kTcSplitWarps;
        const uint64_t pol_stream = l2_policy_evict_first();
        const uint64_t pol_keep = gather_policy == 1 ? l2_policy_evict_normal() : gather_policy == 2 ? l2_policy_evict_first() : l2_policy_evict_last();
        const uint32_t full0 = smem_u32(&full[0]), dst0 = smem_base + lane * 16;
        int it = pw, bi = 0, bph = 0;
        bool ok = true;
        // lane j holds the source-row id j of the NEXT chunk, fetched while the current one is issued
        int coln = 0;
        auto first_chunk_of = [&](int s2) -> long long {       // first chunk of the next non-empty work item
            int g2, tb2;
            for (; work_item(s2, g2, tb2); ++s2)
                if (chunk_ptr[g2 + 1] > chunk_ptr[g2]) return chunk_ptr[g2];
            return -1;
        };
        {
            const long long f = first_chunk_of(0);
            if (f >= 0) coln = __ldg(cols + (size_t)f * kTcKC + lane);
        }
        for (int ws = 0; ok; ++ws) {
            int g, t_begin;
            if (!work_item(ws, g, t_begin)) break;
            const int c_beg = chunk_ptr[g], n_chunks = chunk_ptr[g + 1] - c_beg;
            const int t = min(t_begin + pw / NFC, Tc - 1);             // out-of-range steps are clamped
            const char* b1 = reinterpret_cast<const char*>(src + (size_t)t * s_ts + (pw % NFC) * 128) + lane * 16;
            const char* b2 = HALO ? reinterpret_cast<const char*>(src2 + (size_t)t * s2_ts + (pw % NFC) * 128) + lane * 16 : nullptr;
#pragma unroll 1
            for (int c = 0; c < n_chunks && ok; ++c, it += kTcAcc) {
                const int col = coln;
                const bool in2 = HALO && col >= n_split;
                // row `lane` of the chunk inside its source (bit 31: the row lives in src2); the byte
                // offset is formed in 64 bits per copy (one IMAD.WIDE), so rows * stride may exceed 4 GB
                const uint32_t mine = in2 ? ((uint32_t)(col - n_split) | 0x80000000u) : (uint32_t)col;
                {
                    const long long nxt = (c + 1 < n_chunks) ? (long long)(c_beg + c + 1) : first_chunk_of(ws + 1);
                    if (nxt >= 0) coln = __ldg(cols + (size_t)nxt * kTcKC + lane);
                }
                const int s = it & (kTcStages - 1);
                if (it >= kTcStages && !warp_wait(&empty[s], ((it >> 3) + 1) & 1, &abort_s, err, lane)) { ok = false; break; }
                if (pw == 0 && bph > 0 && !warp_wait(&bfree[bi], (bph - 1) & 1, &abort_s, err, lane)) { ok = false; break; }
                SGP_TRACE(0, it);
                const uint32_t dst = dst0 + s * kTcStageBytes, fbar = full0 + s * 8;
#pragma unroll
                for (int j = 0; j < kTcKC; ++j) {
                    const uint32_t r = __shfl_sync(0xffffffffu, mine, j);
                    const char* p = (HALO && (r >> 31)) ? b2 + (uint64_t)(r & 0x7fffffffu) * s2_nb : b1 + (uint64_t)r * s_nb;
                    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n"
                                 :: "r"(dst + j * 512), "l"(p), "l"(pol_keep));
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(fbar) : "memory");
                if (pw == 0 && lane == 0) {   // the chunk's fp32 slab image -> raw buffer bi (the slab warp splits it)
                    const uint32_t bbar = smem_u32(&braw[bi]);
                    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n"
                                 :: "r"(bbar), "r"(kTcBRawBytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                                 :: "r"(smem_base + kBRawOff + bi * kTcBRawBytes), "l"(bimg + (size_t)(c_beg + c) * (kTcBRawBytes / 4)),
                                    "r"(kTcBRawBytes), "r"(bbar), "l"(pol_stream) : "memory");
                }
                if (++bi == kTcBBufs) { bi = 0; ++bph; }
                __syncwarp();
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (warp < kTcSplitWarps) {
        // ================= split warps: stage (smem) -> A hi | lo tiles (TMEM); epilogue =======
        // four groups of 4 warps; group G takes the items of accumulator a = G (warp & 3 = TMEM lane quarter)
        const int grp = warp >> 2, m = tid & 127;                  // m = feature = TMEM lane
        const uint32_t lane_addr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
        int it0 = 0, cc = 0, wn = 0;                               // items, chunks, non-empty work items so far
        bool ok = true;
        double csum = 0.0;          // fused sink: sum of every value this thread stores (sgp_spmm_rbu_tc `checksum`)
        // one chunk's item of my accumulator: ring stage -> A tile (hi | lo) in TMEM
        auto convert_item = [&]() -> bool {
            const int a = grp;
            const int it = it0 + a, s = it & (kTcStages - 1);
            if (!warp_wait(&full[s], (it >> 3) & 1, &abort_s, err, lane)) return false;
            if ((warp & 3) == 0) SGP_TRACE(2, it);
            if (cc > 0 && !warp_wait(&afree[a], (cc - 1) & 1, &abort_s, err, lane)) return false;
            if ((warp & 3) == 0) SGP_TRACE(3, it);
            const uint32_t rs = smem_base + s * kTcStageBytes + m * 4;     // row-major stage: [k][feature]
            const uint32_t ta = lane_addr + kTcAOff + a * 64;
#pragma unroll
            for (int k0 = 0; k0 < kTcKC; k0 += 16) {       // two halves: 32 live registers, not 64
                uint32_t hv[16], lv[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    float x;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(rs + (k0 + k) * 512));
                    hv[k] = __float_as_uint(x);    // the tensor core ignores the low 13 mantissa bits
                    lv[k] = __float_as_uint(x - __uint_as_float(__float_as_uint(x) & 0xffffe000u));
                }
                SGP_TMEM_ST16(ta + k0, hv);
                SGP_TMEM_ST16(ta + 32 + k0, lv);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&ready[a]);
                mbar_arrive(&empty[s]);
            }
            if ((warp & 3) == 0) SGP_TRACE(4, it);
            SGP_TRACE(8 + (warp & 3), it);
            ++cc;
            it0 += kTcAcc;
            return true;
        };
        bool primed = false;        // chunk 0 of the current work item was converted before the previous drain
        for (int ws = 0; ok; ++ws) {
            int g, t_begin;
            if (!work_item(ws, g, t_begin)) break;
            const int n_chunks = chunk_ptr[g + 1] - chunk_ptr[g];
            // destination rows of the group (lane j holds rows j and 32 + j), for the epilogue
            const int my_row0 = __ldg(grp_rows + (size_t)g * kTcR + lane);
            const int my_row1 = __ldg(grp_rows + (size_t)g * kTcR + 32 + lane);
#pragma unroll 1
            for (int c = primed ? 1 : 0; c < n_chunks && ok; ++c) ok = convert_item();
            primed = false;
            if (!ok) break;
            // Prime the next work item's first chunk BEFORE draining: its MMAs then start as soon
            // as the first accumulator is free, under the rest of the drain.
            {
                int g2, tb2;
                if (work_item(ws + 1, g2, tb2) && chunk_ptr[g2 + 1] > chunk_ptr[g2]) {
                    ok = convert_item();
                    primed = true;
                    if (!ok) break;
                }
            }
            if (!ok) break;
            // ---- epilogue of this work item: thread = TMEM lane = feature; my group's accumulators
            if (n_chunks > 0) {
                if (!warp_wait(&done, wn & 1, &abort_s, err, lane)) { ok = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            // Kept lean on purpose: the drain is exposed (the next work item's MMAs need the
            // accumulators), and its first version — a shuffle, a 64-bit multiply and a branch per
            // row, ~15 dependent instructions — took 9k cycles per work item (38% of the hop).
            // All 16 warps drain accumulator 0 first, then 1, 2, 3 (group G takes rows [16 G, 16 G + 16)
            // of each), so that accfree[0] fires after a quarter of the drain and the next work
            // item's MMAs start under the rest of it.
            const uint32_t d_nb = (uint32_t)d_ns * 4u;           // row stride in bytes (host checks < 2^32)
            int rows[16];
#pragma unroll
            for (int e2 = 0; e2 < 16; ++e2)
                rows[e2] = __shfl_sync(0xffffffffu, (grp < 2) ? my_row0 : my_row1, (grp & 1) * 16 + e2);
#pragma unroll 1
            for (int a = 0; a < kTcAcc; ++a) {
                const int t = t_begin + a / NFC;
                const bool t_ok = t < Tc;
                const char* dp = reinterpret_cast<const char*>(dst + (size_t)min(t, Tc - 1) * d_ts + (a % NFC) * 128 + (warp & 3) * 32 + lane);
                uint32_t v[16];
                if (n_chunks > 0) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]),
                                   "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                                   "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                                 : "r"(lane_addr + a * kTcR + grp * 16));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    // my part of accumulator a is in registers: it may be overwritten by the next work item
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&accfree[a]);
                } else {
#pragma unroll
                    for (int e2 = 0; e2 < 16; ++e2) v[e2] = 0u;          // group without entries: zero rows
                }
                float part = 0.f;
#pragma unroll
                for (int e2 = 0; e2 < 16; ++e2) {
                    if (t_ok && rows[e2] >= 0) {
                        asm volatile("st.global.cs.b32 [%0], %1;"      // streaming (evict-first): written once, read by the next hop's launch
                                     :: "l"(dp + (uint64_t)(uint32_t)rows[e2] * d_nb), "r"(v[e2]) : "memory");
                        part += __uint_as_float(v[e2]);
                    }
                }
                csum += (double)part;
            }
            if (n_chunks > 0) ++wn;
        }
        if (chk != nullptr && ok) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
            if (lane == 0) atomicAdd(chk, csum);
        }
    } else {
      if (warp == kTcSplitWarps + kTcProducerWarps + kTcIssuers) {
        // ================= slab warp: fp32 slab image -> tf32 hi | lo images ======================
        // The operator stores each chunk's [64 rows x 32 columns] slab once, as fp32 in the K-major
        // SWIZZLE_128B layout (8 KB); splitting it here instead of at build time halves the slab
        // stream (17% of the hop's L2 traffic, 37% of its DRAM traffic).  Element-wise, so the
        // layout is untouched: lane l handles the float4 pieces l, l + 32, ...
        int bi = 0, bph = 0;
        bool ok = true;
        for (int ws = 0; ok; ++ws) {
            int g, t_begin_unused;
            if (!work_item(ws, g, t_begin_unused)) break;
            const int n_chunks = chunk_ptr[g + 1] - chunk_ptr[g];
#pragma unroll 1
            for (int c = 0; c < n_chunks && ok; ++c) {
                // raw image landed; the previous user of image buffer bi is done (producer 0 waited
                // for bfree[bi] before it requested this copy)
                if (!warp_wait(&braw[bi], bph & 1, &abort_s, err, lane)) { ok = false; break; }
                const uint32_t rawp = smem_base + kBRawOff + bi * kTcBRawBytes + lane * 16;
                const uint32_t img = smem_base + kBOff + bi * kTcBBytes + lane * 16;
#pragma unroll 4
                for (int j = 0; j < kTcBRawBytes / 512; ++j) {
                    float4 w;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                 : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w) : "r"(rawp + j * 512));
                    float4 h;
                    h.x = __uint_as_float(__float_as_uint(w.x) & 0xffffe000u);
                    h.y = __uint_as_float(__float_as_uint(w.y) & 0xffffe000u);
                    h.z = __uint_as_float(__float_as_uint(w.z) & 0xffffe000u);
                    h.w = __uint_as_float(__float_as_uint(w.w) & 0xffffe000u);
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};"
                                 :: "r"(img + j * 512), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};"
                                 :: "r"(img + kTcBRawBytes + j * 512), "f"(w.x - h.x), "f"(w.y - h.y), "f"(w.z - h.z), "f"(w.w - h.w) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic writes -> tensor-core reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&bfull[bi]);
                if (++bi == kTcBBufs) { bi = 0; ++bph; }
            }
        }
      } else {
        // ================= MMA issuers: two warps, ONE elected thread each runs the whole loop ===
        // issuer q takes the items with (a & 1) == q: a lone thread needs ~500 cycles of
        // instruction latency per item, the tensor pipe 384
        // kind::tf32, fp32 accumulate, A from TMEM (lane = feature, column = k), B K-major smem,
        // N = 64, M = 128.  The loop is kept minimal on purpose: a lone warp issues dependent
        // instructions ~5 cycles apart, and 12 MMAs of 32 cycles leave 384 cycles per item — the
        // first version (warp-wide waits, per-item elect, addresses recomputed per barrier: ~85
        // instructions per item) ran the tensor pipe at half rate (profiles/r1_trace_tc.txt).
        // Barrier addresses and descriptors are computed once; the four items of a chunk are
        // unrolled so that every MMA operand is a constant offset from a handful of registers.
        if (elect_one()) {
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) |
                                       ((uint32_t)(kTcR >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            constexpr uint32_t b_hi32 = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO, version, SW128
            constexpr uint32_t b_lo32 = (16u >> 4) << 16;                          // LBO
            const uint32_t ready0 = smem_u32(&ready[0]), afree0 = smem_u32(&afree[0]), bfree0 = smem_u32(&bfree[0]);
            const uint32_t accfree0 = smem_u32(&accfree[0]), done_a = smem_u32(&done), bfull0 = smem_u32(&bfull[0]);
            // single-thread bounded wait
            auto wait1 = [&](uint32_t bar, uint32_t parity) -> bool {
                uint32_t ok1 = 0;
#pragma unroll 1
                for (int spin = 0; spin < (1 << 24); ++spin) {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                                 "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok1) : "r"(bar), "r"(parity) : "memory");
                    if (ok1) return true;
                    if ((spin & 63) == 63 && abort_s) return false;
                }
                abort_s = 1;
                atomicExch(err, 1);
                return false;
            };
            auto commit1 = [](uint32_t bar) {
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
            };
            auto mma_ts = [](uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t acc) {
                asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %3, 0;\n\tmov.b64 db, {%2, %5};\n\t"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}\n"
                             :: "r"(d), "r"(a_tmem), "r"(b_lo), "r"(acc), "r"(idesc), "r"(b_hi32) : "memory");
            };
            const uint32_t bb0 = b_lo32 | ((smem_base + kBOff) >> 4);
            const int q = warp - (kTcSplitWarps + kTcProducerWarps);
            int bi = 0, bph = 0, cc = 0, wn = 0, itn = q;
            bool ok = true;
            for (int ws = 0; ok; ++ws) {
                int g, t_begin_unused;
                if (!work_item(ws, g, t_begin_unused)) break;
                const int n_chunks = chunk_ptr[g + 1] - chunk_ptr[g];
#pragma unroll 1
                for (int c = 0; c < n_chunks && ok; ++c, ++cc) {
                    const uint32_t bh = bb0 + bi * (kTcBBytes >> 4), bl = bh + (kTcBBytes >> 5);
                    const uint32_t par = cc & 1;
                    if (!wait1(bfull0 + bi * 8, bph & 1)) { ok = false; break; }      // the chunk's slab images have landed
#pragma unroll
                    for (int a2 = 0; a2 < kTcAcc; a2 += kTcIssuers, itn += kTcIssuers) {
                        const int a = a2 + q;
                        SGP_TRACE1(7, itn);
                        if (!wait1(ready0 + a * 8, par)) { ok = false; break; }
                        // first MMA into accumulator a of this work item: the previous one must be drained
                        if (c == 0 && wn > 0 && !wait1(accfree0 + a * 8, (wn - 1) & 1)) { ok = false; break; }
                        SGP_TRACE1(5, itn);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t ah = tmem_d + kTcAOff + a * 64, al = ah + 32;
                        const uint32_t d = tmem_d + a * kTcR;
#pragma unroll
                        for (int ks = 0; ks < kTcKC / 8; ++ks) {
                            mma_ts(d, ah + ks * 8, bh + ks * 2, (c | ks) ? 1u : 0u);
                            mma_ts(d, al + ks * 8, bh + ks * 2, 1u);
                            mma_ts(d, ah + ks * 8, bl + ks * 2, 1u);
                        }
                        commit1(afree0 + a * 8);                          // A tile a may be rewritten
                        if (a2 == kTcAcc - kTcIssuers) {                  // this issuer's last item of the chunk
                            commit1(bfree0 + bi * 8);                     // slab buffer may be refilled
                            if (c == n_chunks - 1) commit1(done_a);       // accumulators complete
                        }
                        SGP_TRACE1(6, itn);
                    }
                    if (++bi == kTcBBufs) { bi = 0; ++bph; }
                }
                if (n_chunks > 0) ++wn;
            }
        }
        __syncwarp();
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "r"(kTcTmemCols));
}

}  // namespace sgp

using namespace sgp;

namespace {
// Tuning knobs are read from the environment ONCE, at the first launch of the process (not per
// call: the ABI is re-entrant across streams and must not depend on hidden mutable state).
struct TcKnobs {
    int group_major, gather_policy;
    TcKnobs() {
        const char* o = getenv("SGP_B200_TC_ORDER");
        group_major = o && o[0] == 'g';
        const char* g = getenv("SGP_B200_TC_GATHER");
        gather_policy = g ? atoi(g) : 0;            // 0 evict_last, 1 normal, 2 evict_first
    }
};
const TcKnobs& tc_knobs() {
    static const TcKnobs k;
    return k;
}
}  // namespace

extern "C" int sgp_spmm_rbu_tc(const int32_t* chunk_ptr, const int32_t* grp_rows, const int32_t* cols,
                               const float* bimg, int n_groups, const float* src, int64_t src_t_stride,
                               int64_t src_n_stride, const float* src2, int64_t src2_t_stride,
                               int64_t src2_n_stride, int n_split, float* dst, int64_t dst_t_stride,
                               int64_t dst_n_stride, int F, int Tc, int* err_flag, double* checksum,
                               void* stream) {
    SGP_REQUIRE(chunk_ptr && grp_rows && cols && bimg && src && dst && err_flag, SGP_EINVAL,
                "sgp_spmm_rbu_tc: null pointer");
    const int nfc = F / 128;
    SGP_REQUIRE(F % 128 == 0 && (nfc == 1 || nfc == 2 || nfc == 4), SGP_EUNSUPPORTED,
                "sgp_spmm_rbu_tc: F=%d (128, 256 or 512)", F);
    SGP_REQUIRE(aligned16(src) && aligned16(dst) && aligned16(bimg) && src_t_stride % 4 == 0 &&
                    src_n_stride % 4 == 0 && (!src2 || (aligned16(src2) && src2_t_stride % 4 == 0 &&
                                                        src2_n_stride % 4 == 0)),
                SGP_EALIGN, "sgp_spmm_rbu_tc: views must be 16-byte aligned with strides %% 4 == 0");
    if (n_groups == 0 || Tc == 0) return SGP_OK;
    if (!src2) n_split = INT32_MAX;
    const int tb = kTcAcc / nfc;
    const int ny = (Tc + tb - 1) / tb;
    SGP_REQUIRE(ny <= 65535, SGP_EUNSUPPORTED, "sgp_spmm_rbu_tc: Tc=%d too large for one launch", Tc);
    // row strides travel as 32-bit byte counts; row * stride is formed in 64 bits inside the kernel
    SGP_REQUIRE(src_n_stride > 0 && src_n_stride * 4 < (1ll << 32) && (!src2 || (src2_n_stride > 0 && src2_n_stride * 4 < (1ll << 32))),
                SGP_EUNSUPPORTED, "sgp_spmm_rbu_tc: row stride too large");
    SGP_REQUIRE(dst_n_stride > 0 && dst_n_stride * 4 < (1ll << 32), SGP_EUNSUPPORTED, "sgp_spmm_rbu_tc: dst row stride too large");
    const uint32_t s_nb = (uint32_t)(src_n_stride * 4), s2_nb = (uint32_t)(src2_n_stride * 4);
    const int n_work = n_groups * ny;
    const int group_major = tc_knobs().group_major, gather_policy = tc_knobs().gather_policy;
    const int n_par = group_major ? n_groups : n_work;
    const int cta_limit = g_tc_cta_limit.load(std::memory_order_relaxed);
    const int grid = n_par < cta_limit ? n_par : cta_limit;      // persistent: one CTA per SM (or fewer: row-sharded runs)
#ifdef SGP_TC_TRACE
    // trace builds only (tools/trace_tc.py): device buffer for the per-item timestamps of one CTA
    long long* trace_ptr = getenv("SGP_B200_TC_TRACE") ? (long long*)strtoull(getenv("SGP_B200_TC_TRACE"), nullptr, 10) : nullptr;
#else
    long long* trace_ptr = nullptr;
#endif
#define SGP_TC3(NFC_, HALO_, PW_)                                                                      \
    do {                                                                                               \
        SGP_CUDA(cudaFuncSetAttribute(spmm_rbu_tc_kernel<NFC_, HALO_, PW_>,                            \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));     \
        spmm_rbu_tc_kernel<NFC_, HALO_, PW_><<<grid, (kTcSplitWarps + PW_ + kTcIssuers + kTcPadWarps) * 32, kTcSmem, as_stream(stream)>>>( \
            chunk_ptr, grp_rows, cols, bimg, n_groups, n_work, group_major, gather_policy, src, src_t_stride, s_nb, src2, \
            src2_t_stride, s2_nb, n_split, dst, dst_t_stride, dst_n_stride, Tc, err_flag, checksum, trace_ptr);  \
    } while (0)
#define SGP_TC(NFC_, HALO_)                                                                            \
    do {                                                                                               \
        SGP_TC3(NFC_, HALO_, SGP_TC_PRODUCER_WARPS);                                                                    \
    } while (0)
    if (src2) {
        if (nfc == 1) SGP_TC(1, true); else if (nfc == 2) SGP_TC(2, true); else SGP_TC(4, true);
    } else {
        if (nfc == 1) SGP_TC(1, false); else if (nfc == 2) SGP_TC(2, false); else SGP_TC(4, false);
    }
#undef SGP_TC
#undef SGP_TC3
    SGP_LAUNCH_CHECK("spmm_rbu_tc");
    return SGP_OK;
}
