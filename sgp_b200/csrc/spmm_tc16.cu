// K2-TC16 — the hop SpMM on the tensor cores with fp16x3 operands and 96-row groups.
//
// Same contract and the same pipeline as spmm_tc.cu (dst[t, i, :] = sum_e val[e] * src[t, col[e], :],
// replacing `x = adj @ x` of lib/sgp_preprocessing.py:200-203; persistent CTA per SM, cp.async gather
// ring, A operand in TMEM, warp-specialised mbarrier pipeline), with two changes that go together:
//   * operands are fp16 hi | lo pairs (kind::f16: tf32's 11-bit significand at twice the MMA rate,
//     three products as before; round-to-nearest splits, powers of two scale x and the operator values
//     into fp16's normal range, fp32 accumulation, the epilogue undoes the scale exactly), so an A tile
//     takes 32 TMEM columns instead of 64;
//   * the 128 TMEM columns this frees hold wider accumulators: groups of 96 rows (4 x 96 + 4 x 32 = 512
//     columns).  A 96-row blob of a 100-NN graph has (sqrt(96) + 10)^2 / 96 = 4.08 union columns per row
//     against 4.96 at 64 rows: 18 % fewer gathered bytes through L2 — the resource the tf32 hop is
//     closest to (71 % of the 9.5 TB/s gather ceiling, profiles/r2_ncu_full_spmm_tc.txt) — while the
//     MMA time per output row still falls by 38 % (N = 96 at the f16 rate).
// The operator stores each chunk's [96 rows x 32 columns] slab already split, as ONE K-major
// SWIZZLE_128B image whose 128-byte rows hold [32 hi | 32 lo] fp16 values (12 KB, the same 4 bytes per
// slab entry as the fp32 image of the tf32 kernel): it lands by TMA directly where the MMAs read it —
// no slab warp.  B k-steps 0, 1 address the hi half of a row, 2, 3 the lo half.
// Requires a bound on |x| (the caller passes x_scale = the power of two with x_scale * max|x| <= 2^14).
#include <stdlib.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace sgp {

constexpr int kT16R = 96;            // rows per group  (MMA N)
constexpr int kT16KC = 32;           // union columns per chunk
constexpr int kT16Acc = 4;           // accumulators per CTA (time steps x feature chunks)
constexpr int kT16SplitWarps = 16;   // 4 groups x TMEM lane quarter
constexpr int kT16ProducerWarps = 4;
#ifndef SGP_T16_ISSUERS
#define SGP_T16_ISSUERS 1          // measured on one box: 1 / 2 / 4 issuers = 73.0 / 74.1 / 75.3 us per hop-panel (profiles/r2_issuers.txt)
#endif
constexpr int kT16Issuers = SGP_T16_ISSUERS;   // MMA issuer warps (one elected thread each); accumulator a belongs to issuer a % issuers
#ifndef SGP_T16_STAGES
#define SGP_T16_STAGES 8
#endif
#ifndef SGP_T16_BBUFS
#define SGP_T16_BBUFS 3
#endif
constexpr int kT16Stages = SGP_T16_STAGES;   // gathered-row ring
constexpr int kT16StageBytes = kT16KC * 128 * 4;     // 16 KB: 32 rows x 128 features fp32, row-major
constexpr int kT16BBytes = kT16R * 128;              // 12 KB: [96 rows][32 hi | 32 lo] fp16
constexpr int kT16BBufs = SGP_T16_BBUFS;
constexpr size_t kT16Smem = (size_t)kT16Stages * kT16StageBytes + kT16BBufs * kT16BBytes + 1024;
constexpr int kT16TmemCols = 512;    // [0, 384) four accumulators of 96 columns, [384, 512) four A tiles of 16 hi + 16 lo columns
constexpr int kT16AOff = kT16Acc * kT16R;
constexpr int kT16Threads = (kT16SplitWarps + kT16ProducerWarps + kT16Issuers) * 32;

__device__ __forceinline__ uint32_t t16_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// L2 eviction policies: slab images and the hop's output stream through L2 once (evict_first) so
// that they do not push out the gathered panel rows, which neighbouring groups re-read (the
// cross-ring reuse distance of the breadth-first group order is about one wave of CTAs).
__device__ __forceinline__ uint64_t t16_l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t t16_l2_policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t t16_l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// elect.sync: exactly one lane of a converged warp.  ptxas knows a single lane is active under this
// predicate and moves MMA operands to uniform registers directly (under `lane == 0` it emits a
// per-operand ELECT / R2UR.BROADCAST / branch waterfall: ~13 instructions per tcgen05.mma).
__device__ __forceinline__ bool t16_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void t16_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(t16_smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void t16_mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(t16_smem_u32(bar)) : "memory");
}

// Bounded warp-wide wait (all 32 lanes poll the same word: one broadcast shared-memory access per
// try).  A barrier that never completes raises the error flag and the CTA-wide abort flag (so that
// every other role stops at once) instead of hanging the GPU.
__device__ __forceinline__ bool t16_warp_wait(uint64_t* bar, uint32_t parity, volatile int* abort_s, int* err, int lane) {
    const uint32_t a = t16_smem_u32(bar);
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

#define SGP_T16_ST16(addr, arr)                                                                   \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
                 :: "r"(addr), "r"(arr[0]), "r"(arr[1]), "r"(arr[2]), "r"(arr[3]), "r"(arr[4]), "r"(arr[5]),  \
                    "r"(arr[6]), "r"(arr[7]), "r"(arr[8]), "r"(arr[9]), "r"(arr[10]), "r"(arr[11]),            \
                    "r"(arr[12]), "r"(arr[13]), "r"(arr[14]), "r"(arr[15]) : "memory")

__device__ __forceinline__ uint32_t t16_bits(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }

// Warp roles (21 warps): 0-15 split (group = accumulator index, warp & 3 = TMEM lane quarter), 16-19
// producers (cp.async gathers; producer 0 also fetches the slab images), 20 the MMA issuer.
// mbarriers as in spmm_tc.cu, minus the raw-slab hand-off (bfull is completed by the TMA bytes).
template <int NFC, bool HALO>
__global__ void __launch_bounds__(kT16Threads, 1)
spmm_rbu_tc16_kernel(const int32_t* __restrict__ chunk_ptr, const int32_t* __restrict__ grp_rows,
                     const int32_t* __restrict__ cols, const __half* __restrict__ bimg,
                     int n_groups, int n_work,
                     const float* __restrict__ src, int64_t s_ts, uint32_t s_nb /* row stride, BYTES */,
                     const float* __restrict__ src2, int64_t s2_ts, uint32_t s2_nb, int n_split,
                     float* __restrict__ dst, int64_t d_ts, int64_t d_ns, int Tc,
                     float x_scale, float inv_scale, int* err, double* __restrict__ chk) {
    // every accumulator's chain must own its stages (stage = item % stages, accumulator = item % 4): a ring whose depth
    // is not a multiple of 4 lets one chain wait for a stage another chain holds — a deadlock (11 stages: barrier time-out)
    static_assert(kT16Acc == 4 && kT16Stages % kT16Acc == 0 && kT16Smem <= 227 * 1024, "ring depth / shared memory");
    constexpr int TB = kT16Acc / NFC;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (t16_smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t full[kT16Stages], empty[kT16Stages], ready[kT16Acc], afree[kT16Acc], bfree[kT16BBufs], bfull[kT16BBufs];
    __shared__ uint64_t done, accfree[kT16Acc];
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int abort_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        abort_s = 0;
        for (int s = 0; s < kT16Stages; ++s) {
            t16_mbar_init(&full[s], 32);             // the 32 lanes' cp.async arrivals
            t16_mbar_init(&empty[s], 4);
        }
        for (int b = 0; b < kT16Acc; ++b) {
            t16_mbar_init(&ready[b], 4);
            t16_mbar_init(&afree[b], 1);
            t16_mbar_init(&accfree[b], kT16SplitWarps);
        }
        for (int b = 0; b < kT16BBufs; ++b) {
            t16_mbar_init(&bfree[b], kT16Issuers);
            t16_mbar_init(&bfull[b], 1);             // producer 0's arrive.expect_tx; the bulk copy completes the bytes
        }
        t16_mbar_init(&done, kT16Issuers);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(t16_smem_u32(&tmem_base_s)), "r"(kT16TmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    constexpr uint32_t kBOff = kT16Stages * kT16StageBytes;             // slab images after the ring
    // s-th work item of this persistent CTA: w = blockIdx.x + s * gridDim.x = (time block, group), time-major
    auto work_item = [&](int s_, int& g_, int& t_begin_) -> bool {
        const int w_ = blockIdx.x + s_ * gridDim.x;
        const int y_ = w_ / n_groups;
        g_ = w_ - y_ * n_groups;
        t_begin_ = y_ * TB;
        return w_ < n_work;
    };

    if (warp >= kT16SplitWarps && warp < kT16SplitWarps + kT16ProducerWarps) {
        // ================= producers: warp p gathers the items with accumulator index a = p ======
        const int pw = warp - kT16SplitWarps;
        const uint64_t pol_stream = t16_l2_policy_evict_first();
        const uint64_t pol_keep = t16_l2_policy_evict_last();
        const uint32_t full0 = t16_smem_u32(&full[0]), dst0 = smem_base + lane * 16;
        int it = pw, bi = 0, bph = 0;
        bool ok = true;
        int coln = 0;
        auto first_chunk_of = [&](int s2) -> long long {
            int g2, tb2;
            for (; work_item(s2, g2, tb2); ++s2)
                if (chunk_ptr[g2 + 1] > chunk_ptr[g2]) return chunk_ptr[g2];
            return -1;
        };
        {
            const long long f = first_chunk_of(0);
            if (f >= 0) coln = __ldg(cols + (size_t)f * kT16KC + lane);
        }
        for (int ws = 0; ok; ++ws) {
            int g, t_begin;
            if (!work_item(ws, g, t_begin)) break;
            const int c_beg = chunk_ptr[g], n_chunks = chunk_ptr[g + 1] - c_beg;
            const int t = min(t_begin + pw / NFC, Tc - 1);
            const char* b1 = reinterpret_cast<const char*>(src + (size_t)t * s_ts + (pw % NFC) * 128) + lane * 16;
            const char* b2 = HALO ? reinterpret_cast<const char*>(src2 + (size_t)t * s2_ts + (pw % NFC) * 128) + lane * 16 : nullptr;
#pragma unroll 1
            for (int c = 0; c < n_chunks && ok; ++c, it += kT16Acc) {
                const int col = coln;
                const bool in2 = HALO && col >= n_split;
                const uint32_t mine = in2 ? ((uint32_t)(col - n_split) | 0x80000000u) : (uint32_t)col;
                {
                    const long long nxt = (c + 1 < n_chunks) ? (long long)(c_beg + c + 1) : first_chunk_of(ws + 1);
                    if (nxt >= 0) coln = __ldg(cols + (size_t)nxt * kT16KC + lane);
                }
                const int s = it % kT16Stages;
                if (it >= kT16Stages && !t16_warp_wait(&empty[s], ((it / kT16Stages) + 1) & 1, &abort_s, err, lane)) { ok = false; break; }
                if (pw == 0 && bph > 0 && !t16_warp_wait(&bfree[bi], (bph - 1) & 1, &abort_s, err, lane)) { ok = false; break; }
                const uint32_t dstp = dst0 + s * kT16StageBytes, fbar = full0 + s * 8;
#pragma unroll
                for (int j = 0; j < kT16KC; ++j) {
                    const uint32_t r = __shfl_sync(0xffffffffu, mine, j);
                    const char* p = (HALO && (r >> 31)) ? b2 + (uint64_t)(r & 0x7fffffffu) * s2_nb : b1 + (uint64_t)r * s_nb;
                    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n"
                                 :: "r"(dstp + j * 512), "l"(p), "l"(pol_keep));
                }
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(fbar) : "memory");
                if (pw == 0 && lane == 0) {   // the chunk's hi | lo slab image, straight into MMA position
                    const uint32_t bbar = t16_smem_u32(&bfull[bi]);
                    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n"
                                 :: "r"(bbar), "r"(kT16BBytes) : "memory");
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                                 :: "r"(smem_base + kBOff + bi * kT16BBytes), "l"(bimg + (size_t)(c_beg + c) * (kT16BBytes / 2)),
                                    "r"(kT16BBytes), "r"(bbar), "l"(pol_stream) : "memory");
                }
                if (++bi == kT16BBufs) { bi = 0; ++bph; }
                __syncwarp();
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (warp < kT16SplitWarps) {
        // ================= split warps: stage (smem) -> fp16 hi | lo A tiles (TMEM); epilogue =====
        const int grp = warp >> 2, m = tid & 127;                  // group = accumulator; m = feature = TMEM lane
        const uint32_t lane_addr = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
        int it0 = 0, cc = 0, wn = 0;
        bool ok = true;
        double csum = 0.0;
        auto convert_item = [&]() -> bool {
            const int a = grp;
            const int it = it0 + a, s = it % kT16Stages;
            if (!t16_warp_wait(&full[s], (it / kT16Stages) & 1, &abort_s, err, lane)) return false;
            const uint32_t rs = smem_base + s * kT16StageBytes + m * 4;     // row-major stage: [k][feature]
            const uint32_t ta = lane_addr + kT16AOff + a * 32;
            uint32_t hv[16], lv[16];
#pragma unroll
            for (int k = 0; k < kT16KC; k += 2) {
                float x0, x1;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x0) : "r"(rs + k * 512));
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x1) : "r"(rs + (k + 1) * 512));
                const float s0 = x0 * x_scale, s1 = x1 * x_scale;
                const __half2 h2 = __floats2half2_rn(s0, s1);       // element k in the low half
                const float2 hf = __half22float2(h2);
                hv[k >> 1] = t16_bits(h2);
                lv[k >> 1] = t16_bits(__floats2half2_rn(s0 - hf.x, s1 - hf.y));
            }
            // the stage is in registers now: hand it back to the producers, and only then wait for the MMAs
            // that still read this accumulator's A tile — the conversion of chunk c + 1 runs under the MMAs of chunk c
            __syncwarp();
            if (lane == 0) t16_mbar_arrive(&empty[s]);
            if (cc > 0) {
                if (!t16_warp_wait(&afree[a], (cc - 1) & 1, &abort_s, err, lane)) return false;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            SGP_T16_ST16(ta, hv);
            SGP_T16_ST16(ta + 16, lv);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) t16_mbar_arrive(&ready[a]);
            ++cc;
            it0 += kT16Acc;
            return true;
        };
        bool primed = false;
        for (int ws = 0; ok; ++ws) {
            int g, t_begin;
            if (!work_item(ws, g, t_begin)) break;
            const int n_chunks = chunk_ptr[g + 1] - chunk_ptr[g];
#pragma unroll 1
            for (int c = primed ? 1 : 0; c < n_chunks && ok; ++c) ok = convert_item();
            primed = false;
            if (!ok) break;
            {   // prime the next work item's first chunk before draining
                int g2, tb2;
                if (work_item(ws + 1, g2, tb2) && chunk_ptr[g2 + 1] > chunk_ptr[g2]) {
                    ok = convert_item();
                    primed = true;
                    if (!ok) break;
                }
            }
            // destination rows of my slice: group G of the split warps drains rows [24 G, 24 G + 24) of every accumulator
            int rows[24];
#pragma unroll
            for (int e2 = 0; e2 < 24; ++e2) rows[e2] = __ldg(grp_rows + (size_t)g * kT16R + grp * 24 + e2);
            if (n_chunks > 0) {
                if (!t16_warp_wait(&done, wn & 1, &abort_s, err, lane)) { ok = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            const uint32_t d_nb = (uint32_t)d_ns * 4u;
#pragma unroll 1
            for (int a = 0; a < kT16Acc; ++a) {
                const int t = t_begin + a / NFC;
                const bool t_ok = t < Tc;
                const char* dp = reinterpret_cast<const char*>(dst + (size_t)min(t, Tc - 1) * d_ts + (a % NFC) * 128 + (warp & 3) * 32 + lane);
                uint32_t v[24];
                if (n_chunks > 0) {
                    const uint32_t ad = lane_addr + a * kT16R + grp * 24;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(ad));
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(ad + 8));
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]) : "r"(ad + 16));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) t16_mbar_arrive(&accfree[a]);
                } else {
#pragma unroll
                    for (int e2 = 0; e2 < 24; ++e2) v[e2] = 0u;
                }
                float part = 0.f;
#pragma unroll
                for (int e2 = 0; e2 < 24; ++e2) {
                    if (t_ok && rows[e2] >= 0) {
                        const float val = __uint_as_float(v[e2]) * inv_scale;
                        asm volatile("st.global.cs.f32 [%0], %1;"
                                     :: "l"(dp + (uint64_t)(uint32_t)rows[e2] * d_nb), "f"(val) : "memory");
                        part += val;
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
        // ================= MMA issuer(s): kT16Issuers warps, ONE elected thread each ================
        // kind::f16, fp32 accumulate, A from TMEM (lane = feature, fp16 pairs along k), B K-major smem,
        // N = 96, M = 128, K = 16: 6 MMAs of 48 cycles per item
        if (t16_elect_one()) {
            constexpr uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) |
                                       ((uint32_t)(kT16R >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            constexpr uint32_t b_hi32 = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO, version, SW128
            constexpr uint32_t b_lo32 = (16u >> 4) << 16;                          // LBO
            const uint32_t ready0 = t16_smem_u32(&ready[0]), afree0 = t16_smem_u32(&afree[0]), bfree0 = t16_smem_u32(&bfree[0]);
            const uint32_t accfree0 = t16_smem_u32(&accfree[0]), done_a = t16_smem_u32(&done), bfull0 = t16_smem_u32(&bfull[0]);
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
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}\n"
                             :: "r"(d), "r"(a_tmem), "r"(b_lo), "r"(acc), "r"(idesc), "r"(b_hi32) : "memory");
            };
            const uint32_t bb0 = b_lo32 | ((smem_base + kBOff) >> 4);
            const int q = warp - (kT16SplitWarps + kT16ProducerWarps);
            int bi = 0, bph = 0, cc = 0, wn = 0;
            bool ok = true;
            for (int ws = 0; ok; ++ws) {
                int g, t_begin_unused;
                if (!work_item(ws, g, t_begin_unused)) break;
                const int n_chunks = chunk_ptr[g + 1] - chunk_ptr[g];
#pragma unroll 1
                for (int c = 0; c < n_chunks && ok; ++c, ++cc) {
                    const uint32_t bh = bb0 + bi * (kT16BBytes >> 4);        // row = [hi: units 0-3 | lo: units 4-7]
                    const uint32_t par = cc & 1;
                    if (!wait1(bfull0 + bi * 8, bph & 1)) { ok = false; break; }
#pragma unroll
                    for (int a2 = 0; a2 < kT16Acc; a2 += kT16Issuers) {
                        const int a = a2 + q;
                        if (!wait1(ready0 + a * 8, par)) { ok = false; break; }
                        if (c == 0 && wn > 0 && !wait1(accfree0 + a * 8, (wn - 1) & 1)) { ok = false; break; }
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t ah = tmem_d + kT16AOff + a * 32, al = ah + 16;
                        const uint32_t d = tmem_d + a * kT16R;
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {                      // K = 16 per MMA: 2 k-steps per chunk
                            mma_ts(d, ah + ks * 8, bh + ks * 2, (c | ks) ? 1u : 0u);
                            mma_ts(d, al + ks * 8, bh + ks * 2, 1u);
                            mma_ts(d, ah + ks * 8, bh + (2 + ks) * 2, 1u);
                        }
                        commit1(afree0 + a * 8);
                        if (a2 == kT16Acc - kT16Issuers) {
                            commit1(bfree0 + bi * 8);
                            if (c == n_chunks - 1) commit1(done_a);
                        }
                    }
                    if (++bi == kT16BBufs) { bi = 0; ++bph; }
                }
                if (n_chunks > 0) ++wn;
            }
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "r"(kT16TmemCols));
}

}  // namespace sgp

using namespace sgp;

extern "C" int sgp_spmm_rbu_tc16(const int32_t* chunk_ptr, const int32_t* grp_rows, const int32_t* cols,
                                 const void* bimg, int n_groups, const float* src, int64_t src_t_stride,
                                 int64_t src_n_stride, const float* src2, int64_t src2_t_stride,
                                 int64_t src2_n_stride, int n_split, float* dst, int64_t dst_t_stride,
                                 int64_t dst_n_stride, int F, int Tc, float x_scale, float w_scale,
                                 int* err_flag, double* checksum, void* stream) {
    SGP_REQUIRE(chunk_ptr && grp_rows && cols && bimg && src && dst && err_flag, SGP_EINVAL,
                "sgp_spmm_rbu_tc16: null pointer");
    const int nfc = F / 128;
    SGP_REQUIRE(F % 128 == 0 && (nfc == 1 || nfc == 2 || nfc == 4), SGP_EUNSUPPORTED,
                "sgp_spmm_rbu_tc16: F=%d (128, 256 or 512)", F);
    SGP_REQUIRE(x_scale > 0.f && w_scale > 0.f, SGP_EINVAL, "sgp_spmm_rbu_tc16: scales %g / %g", (double)x_scale, (double)w_scale);
    SGP_REQUIRE(aligned16(src) && aligned16(dst) && aligned16(bimg) && src_t_stride % 4 == 0 &&
                    src_n_stride % 4 == 0 && (!src2 || (aligned16(src2) && src2_t_stride % 4 == 0 &&
                                                        src2_n_stride % 4 == 0)),
                SGP_EALIGN, "sgp_spmm_rbu_tc16: views must be 16-byte aligned with strides %% 4 == 0");
    if (n_groups == 0 || Tc == 0) return SGP_OK;
    if (!src2) n_split = INT32_MAX;
    const int tb = kT16Acc / nfc;
    const int ny = (Tc + tb - 1) / tb;
    SGP_REQUIRE(src_n_stride > 0 && src_n_stride * 4 < (1ll << 32) && (!src2 || (src2_n_stride > 0 && src2_n_stride * 4 < (1ll << 32))),
                SGP_EUNSUPPORTED, "sgp_spmm_rbu_tc16: row stride too large");
    SGP_REQUIRE(dst_n_stride > 0 && dst_n_stride * 4 < (1ll << 32), SGP_EUNSUPPORTED, "sgp_spmm_rbu_tc16: dst row stride too large");
    const uint32_t s_nb = (uint32_t)(src_n_stride * 4), s2_nb = (uint32_t)(src2_n_stride * 4);
    const long long n_work_ll = (long long)n_groups * ny;
    SGP_REQUIRE(n_work_ll < (1ll << 31), SGP_EUNSUPPORTED, "sgp_spmm_rbu_tc16: too many work items");
    const int n_work = (int)n_work_ll;
    const int cta_limit = g_tc_cta_limit.load(std::memory_order_relaxed);
    const int grid = n_work < cta_limit ? n_work : cta_limit;      // persistent: one CTA per SM (fewer on sharded runs)
    const float inv_scale = 1.f / (x_scale * w_scale);
#define SGP_T16(NFC_, HALO_)                                                                           \
    do {                                                                                               \
        SGP_CUDA(cudaFuncSetAttribute(spmm_rbu_tc16_kernel<NFC_, HALO_>,                               \
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kT16Smem));    \
        spmm_rbu_tc16_kernel<NFC_, HALO_><<<grid, kT16Threads, kT16Smem, as_stream(stream)>>>(         \
            chunk_ptr, grp_rows, cols, reinterpret_cast<const __half*>(bimg), n_groups, n_work, src,   \
            src_t_stride, s_nb, src2, src2_t_stride, s2_nb, n_split, dst, dst_t_stride, dst_n_stride,  \
            Tc, x_scale, inv_scale, err_flag, checksum);                                               \
    } while (0)
    if (src2) {
        if (nfc == 1) SGP_T16(1, true); else if (nfc == 2) SGP_T16(2, true); else SGP_T16(4, true);
    } else {
        if (nfc == 1) SGP_T16(1, false); else if (nfc == 2) SGP_T16(2, false); else SGP_T16(4, false);
    }
#undef SGP_T16
    SGP_LAUNCH_CHECK("spmm_rbu_tc16");
    return SGP_OK;
}
