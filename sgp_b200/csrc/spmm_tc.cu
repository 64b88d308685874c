// K2-TC — the hop SpMM on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-accurate.
//
// Same contract as sgp_spmm_rbu (dst[t, i, :] = sum_e val[e] * src[t, col[e], :], replacing
// `x = adj @ x` of lib/sgp_preprocessing.py:200-203), different formulation: rows are grouped 64
// at a time (the same locality-greedy groups as the RBU format) and a hop becomes, per group,
//     D[f, r] = sum_u  X[u, f] * B[r, u]        f: 128 features, r: 64 rows, u: union columns
// i.e. a dense [128 x U] x [U x 64] GEMM whose A operand is the GATHERED source rows and whose B
// operand is the group's slab of operator values (zero where a row lacks the column).  The zero
// fill that costs the CUDA-core RBU kernel 1.8x of its FFMA2 budget is free here, while the
// gather traffic drops another 2x (U/R = 5.4 at R = 64 against 11.3 at R = 16 on the 100-NN
// graph), which is what lets the hop approach the HBM roofline.
//
//   * A = X^T chunk (32 gathered rows x 128 features): M-major ("MN-major") tf32 operand in the
//     SWIZZLE_128B_BASE32B canonical layout — the only MN-major layout tcgen05 takes for 32-bit
//     types.  Every lane cp.async's 16 bytes of a gathered row straight to its swizzled place, so
//     the gather needs no registers and no transposition.
//   * B = slab chunk (64 rows x 32 columns), K-major SWIZZLE_128B, pre-swizzled at operator build
//     time, streamed linearly with cp.async and reused for 8 accumulators (4 time steps x 2 feature
//     chunks at F = 256) so that its HBM traffic is amortised.
//   * precision: 3xTF32.  hi = x with the low 13 mantissa bits cleared, lo = x - hi, and
//     D += Ah*Bh + Al*Bh + Ah*Bl with fp32 accumulation in TMEM.  The tensor core ignores the low
//     13 bits of a tf32 operand, so the gathered fp32 rows are used as Ah as they are and only Al
//     is produced (one shared-memory pass per tile); B is split at operator build time.
//     Measured 1e-6..3e-6 relative (tools/microbench/umma_tf32_test.cu), well inside 1e-4.
//   * accumulators: 8 x [128 lanes x 64 columns] fp32 = all 512 TMEM columns, one CTA per SM.
//   * warp-specialised pipeline per item (chunk, accumulator), all hand-offs through mbarriers:
//     4 producer warps gather with cp.async into an 8-stage ring (up to 7 items = 112 KB in flight
//     per SM), two groups of 4 "split" warps alternate items and write the lo tiles (4 of them),
//     one elected lane of the MMA warp issues 12 tcgen05.mma per item and tcgen05.commit's the
//     barriers that recycle the ring stage and the lo tile.  Lessons that shaped it (profiles/):
//     fence.proxy.async drains a thread's outstanding cp.async (so gathers and fences live in
//     different warps); 128 threads arriving/polling on one mbarrier serialise (warp-level
//     arrive); `lane == 0` makes ptxas wrap every MMA in an R2UR waterfall (elect.sync does not);
//     a role's loop must stay small (three roles share one 32 KB instruction cache).
//   * measured limit: shared-memory bandwidth (operand fetches of 12 N=64 MMAs = 72 KB + 48 KB
//     of gather / lo-pass traffic per item at 128 B/clk).
// Bound: HBM.  Algorithmic bytes per hop-panel 8 nnz + 4(N+1) + 8 N F (285 MB at C4).
#include <stdlib.h>

#include "common.cuh"

namespace sgp {

constexpr int kTcR = 64;            // rows per group  (MMA N)
constexpr int kTcKC = 32;           // union columns per chunk
constexpr int kTcAcc = 8;           // accumulators per CTA (time steps x feature chunks)
constexpr int kTcProducerWarps = 4; // gather (cp.async) warps
constexpr int kTcSplitGroups = 2;   // split groups alternate items (a even / a odd)
constexpr int kTcSplitWarps = 4 * kTcSplitGroups;   // lo-pass + epilogue warps (warp & 3 = TMEM lane quarter)
constexpr int kTcThreads = (kTcSplitWarps + kTcProducerWarps + 1) * 32;   // + 1 MMA-issuing warp
constexpr int kTcLag = 5;           // cp.async groups a producer thread keeps in flight
constexpr int kTcStages = 8;        // gathered-row ring (the raw fp32 rows ARE the tf32 "hi" operand)
constexpr int kTcStageBytes = kTcKC * 128 * 4;      // 16 KB: 32 rows x 128 features
constexpr int kTcBBytes = 2 * kTcR * kTcKC * 4;     // 16 KB: hi + lo image of one chunk
constexpr int kTcLoTiles = 4;       // A-lo tiles: the lo pass runs up to 4 items ahead of the tensor pipe
constexpr size_t kTcSmem = (size_t)kTcStages * kTcStageBytes + kTcLoTiles * kTcStageBytes + 2 * kTcBBytes + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;       // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// elect.sync: exactly one lane of a converged warp.  ptxas knows a single lane is active under this
// predicate and moves MMA operands to uniform registers directly (under `lane == 0` it emits a
// per-operand ELECT / R2UR.BROADCAST / branch waterfall: ~13 instructions per tcgen05.mma).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(smem_u32(bar)) : "memory");
}

// arrive on `bar` once all cp.async issued so far by this thread have landed (count pre-accounted)
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
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

// byte offset of (gathered row k in [0,32), 16-byte piece q in [0,32)) inside an A stage:
// atoms [k/4][q/8] of 4 rows x 128 B; the 32-byte chunk index inside a row is XORed with k%4
__device__ __forceinline__ int a_stage_offset(int k, int q) {
    const int r = k & 3, ch = q & 7;
    return ((k >> 2) * 4 + (q >> 3)) * 512 + r * 128 + ((((ch >> 1) ^ r) << 5) | ((ch & 1) << 4));
}

// Warp roles: warps 0-3 "split" (lo pass, then the TMEM epilogue: warp w owns TMEM lanes 32w..),
// warps 4-7 "producer" (cp.async gathers + slab images), warp 8 issues the MMAs.
// mbarriers: full[s]  producers -> split/MMA  : stage s holds item i's gathered rows (+ B images)
//            ready[b] split     -> MMA        : lo tile b written and fenced for the tensor proxy
//            empty[s] MMA (tcgen05.commit) -> producers : the MMAs reading stage s have completed
//            lofree[b] MMA (tcgen05.commit) -> split    : the MMAs reading lo tile b have completed
//            done     MMA -> epilogue
// An item is (chunk c, accumulator a): with 8 ring stages and 8 accumulators per chunk the stage
// index IS the accumulator index, so the 8 items of a chunk are fully unrolled and every stage
// address, barrier parity, time step and feature chunk of an item is a compile-time constant.
// (A single warp runs a dependent instruction chain at ~5 cycles per instruction: the per-item
// instruction count of each role, not memory or the tensor pipe, was what bounded this kernel.)
template <int NFC, bool HALO>
__global__ void __launch_bounds__(kTcThreads, 1)
spmm_rbu_tc_kernel(const int32_t* __restrict__ chunk_ptr, const int32_t* __restrict__ grp_rows,
                   const int32_t* __restrict__ cols, const float* __restrict__ bimg,
                   const float* __restrict__ src, int64_t s_ts, uint32_t s_nb /* row stride, BYTES */,
                   const float* __restrict__ src2, int64_t s2_ts, uint32_t s2_nb, int n_split,
                   float* __restrict__ dst, int64_t d_ts, int64_t d_ns, int Tc, int* err) {
    static_assert(kTcStages == kTcAcc && kTcLoTiles == 4, "stage index == accumulator index; lo tile = a & 3");
    constexpr int TB = kTcAcc / NFC;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;      // shared-window address
    __shared__ uint64_t full[kTcStages], empty[kTcStages], ready[kTcLoTiles], lofree[kTcLoTiles], done;
    __shared__ uint32_t tmem_base_s;
    __shared__ int rows_s[kTcR];
    __shared__ volatile int abort_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x;
    const int t_begin = blockIdx.y * TB;
    const int c_beg = chunk_ptr[g], n_chunks = chunk_ptr[g + 1] - c_beg;

    if (n_chunks == 0) {                   // group without stored entries: its rows are zero
        for (int a = 0; a < kTcAcc; ++a) {
            const int t = t_begin + a / NFC, fc = a % NFC;
            if (t >= Tc) continue;
            for (int i = tid; i < kTcR * 32; i += kTcThreads) {
                const int row = grp_rows[(size_t)g * kTcR + (i >> 5)];
                if (row >= 0)
                    st_f4(dst + (size_t)t * d_ts + (size_t)row * d_ns + fc * 128 + (i & 31) * 4,
                          make_float4(0.f, 0.f, 0.f, 0.f));
            }
        }
        return;
    }

    if (tid < kTcR) rows_s[tid] = grp_rows[(size_t)g * kTcR + tid];
    if (tid == 0) {
        abort_s = 0;
        for (int s = 0; s < kTcStages; ++s) {
            mbar_init(&full[s], kTcProducerWarps);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < kTcLoTiles; ++b) {
            mbar_init(&ready[b], 4);
            mbar_init(&lofree[b], 1);
        }
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&tmem_base_s)), "r"(kTcAcc * kTcR));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    constexpr uint32_t kLoOff = kTcStages * kTcStageBytes;              // lo tiles after the ring
    constexpr uint32_t kBOff = (kTcStages + kTcLoTiles) * kTcStageBytes;         // slab images after them

    if (warp >= kTcSplitWarps && warp < kTcSplitWarps + kTcProducerWarps) {
        // ================= producers: item (c, a) -> ring stage a ==============================
        const int pw = warp - kTcSplitWarps, ptid = tid - kTcSplitWarps * 32;
        constexpr int kRowsPerWarp = kTcKC / kTcProducerWarps;       // 8 gathered rows per warp per item
        uint32_t dst_off[kRowsPerWarp];                              // swizzled byte offset of my 16 B
#pragma unroll
        for (int j = 0; j < kRowsPerWarp; ++j)
            dst_off[j] = smem_base + a_stage_offset(pw + j * kTcProducerWarps, lane);
        // per-accumulator source bases (time step, feature chunk; out-of-range steps clamped)
        auto base_of = [&](const float* sbase, int64_t ts, int a) -> const char* {
            const int t = min(t_begin + a / NFC, Tc - 1);
            return reinterpret_cast<const char*>(sbase + (size_t)t * ts + (a % NFC) * 128 + lane * 4);
        };
        uint32_t off1[kRowsPerWarp];      // byte offset of each of my rows inside its source
        bool in2[kRowsPerWarp];
        int colr[kRowsPerWarp];
        auto load_cols = [&](int c) {
#pragma unroll
            for (int j = 0; j < kRowsPerWarp; ++j)
                colr[j] = (c < n_chunks) ? __ldg(cols + (size_t)(c_beg + c) * kTcKC + pw + j * kTcProducerWarps) : 0;
        };
        load_cols(0);
        // Completion is signalled per WARP and kTcLag items late: a thread keeps kTcLag cp.async
        // groups in flight, waits for the oldest one, and lane 0 arrives on that item's barrier.
        bool ok = true;
        for (int c = 0; c < n_chunks + 1 && ok; ++c) {
            if (c < n_chunks) {
#pragma unroll
                for (int j = 0; j < kRowsPerWarp; ++j) {
                    in2[j] = HALO && colr[j] >= n_split;
                    off1[j] = in2[j] ? (uint32_t)(colr[j] - n_split) * s2_nb : (uint32_t)colr[j] * s_nb;
                }
                load_cols(c + 1);          // next chunk's row ids: in flight during this chunk
            }
#pragma unroll 1
            for (int a = 0; a < kTcAcc; ++a) {
                if (c < n_chunks) {
                    if (c > 0 && !warp_wait(&empty[a], (c - 1) & 1, &abort_s, err, lane)) { ok = false; break; }
                    const char* b1 = base_of(src, s_ts, a);
                    const char* b2 = HALO ? base_of(src2, s2_ts, a) : nullptr;
#pragma unroll
                    for (int j = 0; j < kRowsPerWarp; ++j) {
                        const char* p = (HALO && in2[j]) ? b2 + off1[j] : b1 + off1[j];
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n"
                                     :: "r"(dst_off[j] + a * kTcStageBytes), "l"(p));
                    }
                    if (a == 0) {   // the chunk's slab images (hi | lo), reused by its 8 items
                        const float* bs = bimg + (size_t)(c_beg + c) * (kTcBBytes / 4);
                        const uint32_t bd = smem_base + kBOff + (c & 1) * kTcBBytes;
#pragma unroll
                        for (int j = 0; j < kTcBBytes / 16 / (kTcProducerWarps * 32); ++j)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n"
                                         :: "r"(bd + (j * kTcProducerWarps * 32 + ptid) * 16),
                                            "l"(bs + (j * kTcProducerWarps * 32 + ptid) * 4));
                    }
                }
                cp_async_commit();
                // signal the item issued kTcLag commits ago: (c, a - kTcLag) or (c - 1, a - kTcLag + 8)
                if (c > 0 || a >= kTcLag) {
                    cp_async_wait<kTcLag>();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[(a + kTcAcc - kTcLag) % kTcAcc]);
                }
                if (c == n_chunks && a == kTcLag - 1) break;      // drained: all items signalled
            }
        }
        cp_async_wait<0>();
    } else if (warp < kTcSplitWarps) {
        // ================= split warps: lo = x - tf32(x), then the epilogue =====================
        // two groups of 4 warps; group G takes the items with (a & 1) == G, so a group has two item
        // periods for its waits + lo pass and the tensor pipe is never the one that waits
        const int grp = warp >> 2, gtid = tid & 127;
        bool ok = true;
        for (int c = 0; c < n_chunks && ok; ++c) {
#pragma unroll 1
            for (int a = grp; a < kTcAcc; a += kTcSplitGroups) {
                const int b = a & (kTcLoTiles - 1);
                if (!warp_wait(&full[a], c & 1, &abort_s, err, lane)) { ok = false; break; }
                if ((c > 0 || a >= kTcLoTiles) && !warp_wait(&lofree[b], ((a >> 2) + 1) & 1, &abort_s, err, lane)) { ok = false; break; }
                const uint32_t rs = smem_base + a * kTcStageBytes + gtid * 16;
                const uint32_t lo = smem_base + kLoOff + b * kTcStageBytes + gtid * 16;
                float4 v[kTcStageBytes / 16 / 128];
#pragma unroll
                for (int j = 0; j < kTcStageBytes / 16 / 128; ++j)
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                 : "=f"(v[j].x), "=f"(v[j].y), "=f"(v[j].z), "=f"(v[j].w)
                                 : "r"(rs + j * 128 * 16));
#pragma unroll
                for (int j = 0; j < kTcStageBytes / 16 / 128; ++j) {
                    float4 l;
                    l.x = v[j].x - __uint_as_float(__float_as_uint(v[j].x) & 0xffffe000u);
                    l.y = v[j].y - __uint_as_float(__float_as_uint(v[j].y) & 0xffffe000u);
                    l.z = v[j].z - __uint_as_float(__float_as_uint(v[j].z) & 0xffffe000u);
                    l.w = v[j].w - __uint_as_float(__float_as_uint(v[j].w) & 0xffffe000u);
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};"
                                 :: "r"(lo + j * 128 * 16), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w)
                                 : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // lo tile (and, by cumulativity,
                __syncwarp();                                                  // the gathered stage) -> tensor proxy
                if (lane == 0) mbar_arrive(&ready[b]);
            }
        }
        if (ok) ok = warp_wait(&done, 0, &abort_s, err, lane);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (ok) {
            // thread = TMEM lane = feature; 8 accumulators x 64 row columns
            const uint32_t taddr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
#pragma unroll 1
            for (int a = grp; a < kTcAcc; a += kTcSplitGroups) {
                const int t = t_begin + a / NFC;
                if (t >= Tc) continue;
                float* dp = dst + (size_t)t * d_ts + (a % NFC) * 128 + (warp & 3) * 32 + lane;
#pragma unroll
                for (int j = 0; j < kTcR; j += 16) {
                    uint32_t v[16];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]),
                                   "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
                                   "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                                 : "r"(taddr + a * kTcR + j));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int row = rows_s[j + e];
                        if (row >= 0) dp[(size_t)row * d_ns] = __uint_as_float(v[e]);
                    }
                }
            }
        }
    } else {
        // ================= MMA issuer: the whole warp runs the loop, one lane issues ==========
        // kind::tf32, fp32 accumulate, A M-major (gathered rows), B K-major, N = 64, M = 128
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (0u << 16) |
                                   ((uint32_t)(kTcR >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        // descriptors differ only in their 14-bit start-address field
        constexpr uint32_t a_hi32 = (2048u >> 4) | (1u << 14) | (1u << 29);   // SBO, version, SW128_BASE32B
        constexpr uint32_t b_hi32 = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO, version, SW128
        constexpr uint32_t a_lo32 = (512u >> 4) << 16, b_lo32 = (16u >> 4) << 16;   // LBO
        auto desc = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
        const uint32_t raw0 = a_lo32 | (smem_base >> 4);
        const uint32_t lo0 = a_lo32 | ((smem_base + kLoOff) >> 4);
        const uint32_t bb0 = b_lo32 | ((smem_base + kBOff) >> 4);
        bool ok = true;
        for (int c = 0; c < n_chunks && ok; ++c) {
            const uint32_t bh = bb0 + (c & 1) * (kTcBBytes >> 4), bl = bh + (kTcBBytes >> 5);
#pragma unroll 1
            for (int a = 0; a < kTcAcc; ++a) {
                const int b = a & (kTcLoTiles - 1);
                if (!warp_wait(&ready[b], (a >> 2) & 1, &abort_s, err, lane)) { ok = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint32_t ah = raw0 + a * (kTcStageBytes >> 4);
                    const uint32_t al = lo0 + b * (kTcStageBytes >> 4);
                    const uint32_t d = tmem_d + a * kTcR;
#pragma unroll
                    for (int ks = 0; ks < kTcKC / 8; ++ks) {
                        const uint64_t dah = desc(ah + ks * 256, a_hi32), dal = desc(al + ks * 256, a_hi32);
                        const uint64_t dbh = desc(bh + ks * 2, b_hi32), dbl = desc(bl + ks * 2, b_hi32);
                        umma_tf32(d, dah, dbh, idesc, (c | ks) ? 1u : 0u);
                        umma_tf32(d, dal, dbh, idesc, 1u);
                        umma_tf32(d, dah, dbl, idesc, 1u);
                    }
                    umma_commit(&empty[a]);      // ring stage a may be refilled
                    umma_commit(&lofree[b]);     // lo tile b may be rewritten
                }
                __syncwarp();
            }
        }
        if (ok && elect_one()) umma_commit(&done);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "r"(kTcAcc * kTcR));
}

}  // namespace sgp

using namespace sgp;

extern "C" int sgp_spmm_rbu_tc(const int32_t* chunk_ptr, const int32_t* grp_rows, const int32_t* cols,
                               const float* bimg, int n_groups, const float* src, int64_t src_t_stride,
                               int64_t src_n_stride, const float* src2, int64_t src2_t_stride,
                               int64_t src2_n_stride, int n_split, float* dst, int64_t dst_t_stride,
                               int64_t dst_n_stride, int F, int Tc, int* err_flag, void* stream) {
    SGP_REQUIRE(chunk_ptr && grp_rows && cols && bimg && src && dst && err_flag, SGP_EINVAL,
                "sgp_spmm_rbu_tc: null pointer");
    const int nfc = F / 128;
    SGP_REQUIRE(F % 128 == 0 && (nfc == 1 || nfc == 2 || nfc == 4 || nfc == 8), SGP_EUNSUPPORTED,
                "sgp_spmm_rbu_tc: F=%d (128, 256, 512 or 1024)", F);
    SGP_REQUIRE(aligned16(src) && aligned16(dst) && aligned16(bimg) && src_t_stride % 4 == 0 &&
                    src_n_stride % 4 == 0 && (!src2 || (aligned16(src2) && src2_t_stride % 4 == 0 &&
                                                        src2_n_stride % 4 == 0)),
                SGP_EALIGN, "sgp_spmm_rbu_tc: views must be 16-byte aligned with strides %% 4 == 0");
    if (n_groups == 0 || Tc == 0) return SGP_OK;
    if (!src2) n_split = INT32_MAX;
    const int tb = kTcAcc / nfc;
    const int ny = (Tc + tb - 1) / tb;
    SGP_REQUIRE(ny <= 65535, SGP_EUNSUPPORTED, "sgp_spmm_rbu_tc: Tc=%d too large for one launch", Tc);
    // gathered-row byte offsets are 32-bit inside the kernel
    SGP_REQUIRE(src_n_stride > 0 && src_n_stride * 4 < (1ll << 31) / 64 && (!src2 || (src2_n_stride > 0 && src2_n_stride * 4 < (1ll << 31) / 64)),
                SGP_EUNSUPPORTED, "sgp_spmm_rbu_tc: row stride too large");
    const uint32_t s_nb = (uint32_t)(src_n_stride * 4), s2_nb = (uint32_t)(src2_n_stride * 4);
    dim3 grid((unsigned)n_groups, ny);
#define SGP_TC(NFC_, HALO_)                                                                            \
    do {                                                                                               \
        SGP_CUDA(cudaFuncSetAttribute(spmm_rbu_tc_kernel<NFC_, HALO_>,                                 \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));     \
        spmm_rbu_tc_kernel<NFC_, HALO_><<<grid, kTcThreads, kTcSmem, as_stream(stream)>>>(             \
            chunk_ptr, grp_rows, cols, bimg, src, src_t_stride, s_nb, src2, src2_t_stride, s2_nb,      \
            n_split, dst, dst_t_stride, dst_n_stride, Tc, err_flag);                                \
    } while (0)
    if (src2) {
        if (nfc == 1) SGP_TC(1, true); else if (nfc == 2) SGP_TC(2, true);
        else if (nfc == 4) SGP_TC(4, true); else SGP_TC(8, true);
    } else {
        if (nfc == 1) SGP_TC(1, false); else if (nfc == 2) SGP_TC(2, false);
        else if (nfc == 4) SGP_TC(4, false); else SGP_TC(8, false);
    }
#undef SGP_TC
    SGP_LAUNCH_CHECK("spmm_rbu_tc");
    return SGP_OK;
}
