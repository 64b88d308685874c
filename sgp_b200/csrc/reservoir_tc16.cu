// K1-TC16 — the leaky-ESN scan on the tensor cores with fp16x3 operands (tcgen05 kind::f16).
//
// Same contract as sgp_reservoir_scan_tc (one layer, Tc steps, state carried; replaces the Python
// loop of lib/nn/reservoir/reservoir.py:158-186 around :77-81) for tanh reservoirs, H in {128, 256},
// Fin <= 8, with states bounded by 1 (tanh + leaky blend from a zero or bounded initial state).
// Why a second tensor-core scan: the tf32 kernel's step is 21 k cycles against 12.3 k of MMAs, the
// rest being waits on the W ring (512 KB of tf32 hi | lo images per step and SM; neither a deeper
// ring — shared memory is full — nor sharing the stream between CTA pairs helped, see
// profiles/r2_scan_pair_multicast.txt).  fp16 has tf32's 11-bit significand, so the same three-
// product split
//     h W^T ~= Hh Wh^T + Hl Wh^T + Hh Wl^T,   Xh = fp16(s X), Xl = fp16(s X - Xh)
// keeps 22 bits (round-to-nearest splits: 2.2e-7 after 1000 steps against 8.5e-7 for the truncating
// tf32 split, CPU emulation and tests), while kind::f16 MMAs run at twice the tf32 rate (96 MMAs of
// N = 128 = 6.1 k cycles per step at H = 256) and the W stream halves (256 KB per step).  Powers of
// two scale both operands into fp16's normal range (states by 2^14, weights by the power of two that
// brings max|W| to [2^13, 2^14)); the accumulators are fp32 and the epilogue undoes the scale exactly.
//   * A hi  = fp16 state tile [128 nodes x H], K-major SWIZZLE_128B in shared memory (64 KB at H = 256);
//   * A lo  = fp16 pairs packed in TMEM (H / 2 columns), TS-mode MMA;
//   * B     = W_hh hi / lo fp16 images [128 n x 64 k] (16 KB), streamed through a 5-stage TMA ring in
//             exactly the order the MMAs consume them;
//   * D     = fp32 accumulators in TMEM [128 lanes x H columns];
//   * the fp32 state is NOT kept: the leaky blend reads the old state back as hi + lo (22 bits: the
//     blend weighs it by 1 - alpha, an error of 0.1 * 2^-22 per step), and every epilogue warp sends
//     its [32 nodes x 32 columns] block of the new state to the encoder output through a private
//     4 KB staging tile and one TMA tensor store.
// Pipelining inside the tile as in reservoir_tc.cu (column halves x 64-column state chunks).
// Warp roles: 0-15 epilogue, 16 W producer, 17 MMA issuer.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace sgp {

constexpr int kR16EpiWarps = 16;
constexpr int kR16Threads = (kR16EpiWarps + 2) * 32;
constexpr int kR16WStage = 128 * 64 * 2;     // 16 KB: one [128 n x 64 k] fp16 image (hi OR lo)
#ifndef SGP_R16_STAGES
#define SGP_R16_STAGES 5
#endif
constexpr int kR16WStages = SGP_R16_STAGES;
constexpr int kR16MaxFin = 8;
constexpr int kR16OutTile = 32 * 32 * 4;     // 4 KB per epilogue warp
constexpr float kR16StateScale = 16384.f;    // 2^14

__device__ __forceinline__ uint32_t r16_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void r16_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(r16_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void r16_mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(r16_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void r16_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(r16_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool r16_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ bool r16_wait(uint64_t* bar, uint32_t parity, volatile int* abort_s, int* err, int lane) {
    const uint32_t a = r16_smem_u32(bar);
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
constexpr float kR16TanhC = 2.885390081777927f;           // 2 log2 e
__device__ __forceinline__ float r16_rcp_ex2p1(float y) {  // 1 / (2^y + 1); tanh(z) = 1 - 2 r at y = 2 z log2 e (1e-7 absolute)
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    return r;
}

#define SGP_R16_LD32(addr, v)                                                                        \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15," \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                       \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),   \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),          \
                   "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),        \
                   "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),        \
                   "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                              \
                 : "r"(addr))
#define SGP_R16_LD16(addr, v)                                                                        \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),   \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) \
                 : "r"(addr))
#define SGP_R16_ST16(addr, v)                                                                        \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
                 :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),   \
                    "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]),           \
                    "r"(v[14]), "r"(v[15]) : "memory")

// byte offset of element (row m, column k) of the fp16 K-major SWIZZLE_128B state tile [128 x H]:
// k-blocks of 64 columns (16 KB each), atoms of 8 rows x 128 B, 16-byte unit XOR (row % 8)
__device__ __forceinline__ uint32_t a16_offset(int m, int k) {
    return (uint32_t)((k >> 6) * 16384 + (m >> 3) * 1024 + (m & 7) * 128 + ((((k & 63) >> 3) ^ (m & 7)) << 4) + (k & 7) * 2);
}
// fp16 pairs, in shared memory and in a TMEM cell alike: element 2j in the low half, 2j + 1 in the high
// half (validated on hardware: the swapped order is 5e-4 off).  All fp32 <-> fp16 conversions are the
// PACKED ones (cvt.rn.f16x2.f32 = F2FP, full rate): the scalar F2F conversions of the first version
// ran on the XU pipe next to the two MUFU ops of every tanh and made the epilogue XU-bound (49 % of
// the pipe's peak over the whole launch, profiles/r2_ncu_full_reservoir_tc16.txt).
__device__ __forceinline__ uint32_t r16_bits(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ __half2 r16_half2(uint32_t w) { return *reinterpret_cast<__half2*>(&w); }

template <int H, int FINP>
__global__ void __launch_bounds__(kR16Threads, 1)
reservoir_tc16_kernel(const float* __restrict__ x, int64_t x_ts, int64_t x_ns, int Fin,
                      const __half* __restrict__ wimg /* [H/128][H/64][hi|lo][128*64], consumption order */,
                      const float* __restrict__ w_ih /* [H, Fin] */, const float* __restrict__ bias,
                      float alpha, float oma, float inv_scale /* 1 / (state scale * weight scale) */,
                      float* __restrict__ h_state, const __grid_constant__ CUtensorMap out_map,
                      int Tc, int N, int* err, double* __restrict__ chk) {
    constexpr int NH = H / 128;                 // output-column halves (MMA N = 128)
    constexpr int NC = H / 64;                  // state chunks of 64 columns (one swizzle atom wide)
    constexpr int A_BYTES = 128 * H * 2;        // fp16 state tile
    constexpr int kStagesPerStep = NC * NH * 2; // (half, chunk, hi|lo) images per step
    constexpr int TMEM_COLS = (H + H / 2 <= 256) ? 256 : 512;
    constexpr int ALO_OFF = H;                  // TMEM columns [0,H) = D, [H, H + H/2) = A lo (fp16 pairs)
    constexpr int NFREE = (NH - 1) * 2;         // chunks owned by the halves before the last
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (r16_smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - r16_smem_u32(smem_raw));
    // [A hi tile | W ring | output staging (16 x 4 KB) | w_ih (FINP x H) | bias (H)]
    const uint32_t a_base = smem_base, w_base = smem_base + A_BYTES;
    const uint32_t o_base = w_base + kR16WStages * kR16WStage;
    float* wih_s = reinterpret_cast<float*>(smem + A_BYTES + kR16WStages * kR16WStage + kR16EpiWarps * kR16OutTile);
    float* bias_s = wih_s + kR16MaxFin * H;
    __shared__ uint64_t wfull[kR16WStages], wempty[kR16WStages];
    __shared__ uint64_t acc_ready[2], d_free[2], a_ready[4], a_free[4];
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int abort_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * 128;

    if (tid == 0) {
        abort_s = 0;
        for (int s = 0; s < kR16WStages; ++s) {
            r16_mbar_init(&wfull[s], 1);
            r16_mbar_init(&wempty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            r16_mbar_init(&acc_ready[i], 1);
            r16_mbar_init(&d_free[i], kR16EpiWarps);
        }
        for (int i = 0; i < 4; ++i) {
            r16_mbar_init(&a_ready[i], 8);        // 4 lane quarters x 2 column groups write a 64-column chunk
            r16_mbar_init(&a_free[i], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < FINP * H; i += kR16Threads) {
        const int f = i / H, n = i % H;
        wih_s[i] = (f < Fin) ? w_ih[(size_t)n * Fin + f] * kR16TanhC : 0.f;      // pre-activation in units of
    }                                                                              // 1 / (2 log2 e): see the blend
    for (int i = tid; i < H; i += kR16Threads) bias_s[i] = bias[i] * kR16TanhC;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(r16_smem_u32(&tmem_base_s)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;

    if (warp == kR16EpiWarps) {
        // ================= producer warp: the W images, in consumption order =====================
        const long long total = (long long)Tc * kStagesPerStep;
        int s = 0, ph = 0, js = 0;
        for (long long j = 0; j < total; ++j) {
            if (ph > 0 && !r16_wait(&wempty[s], (ph - 1) & 1, &abort_s, err, lane)) break;
            if (r16_elect_one()) {
                const __half* src = wimg + (size_t)js * (kR16WStage / 2);
                const uint32_t bar = r16_smem_u32(&wfull[s]);
                asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n"
                             :: "r"(bar), "r"(kR16WStage) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(w_base + s * kR16WStage), "l"(src), "r"(kR16WStage), "r"(bar) : "memory");
            }
            __syncwarp();
            if (++s == kR16WStages) { s = 0; ++ph; }
            if (++js == kStagesPerStep) js = 0;
        }
    } else if (warp < kR16EpiWarps) {
        // ================= epilogue warps: thread = node, warp = (lane quarter, column group) =====
        const int q4 = warp & 3, cg = warp >> 2;
        const int m = q4 * 32 + lane, node = n0 + m;
        const bool live = node < N;
        const uint32_t lane_addr = tmem_d + ((uint32_t)(q4 * 32) << 16);
        const uint32_t my_out = o_base + warp * kR16OutTile;          // [32 rows x 128 B], 16-byte unit XOR (row & 7)
        // new state (32 columns of chunk c starting at ko) -> fp16 hi (smem) + lo pairs (TMEM); publish
        auto publish = [&](int c, int ko, const float (&hn)[32]) {
            uint32_t lo[16];
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint32_t hp[4];
#pragma unroll
                for (int e = 0; e < 8; e += 2) {
                    const float s0 = hn[j + e] * kR16StateScale, s1 = hn[j + e + 1] * kR16StateScale;
                    const __half2 h2 = __floats2half2_rn(s0, s1);
                    const float2 hf = __half22float2(h2);
                    hp[e >> 1] = r16_bits(h2);
                    lo[(j + e) >> 1] = r16_bits(__floats2half2_rn(s0 - hf.x, s1 - hf.y));
                }
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};"
                             :: "r"(a_base + a16_offset(m, c * 64 + ko + j)), "r"(hp[0]), "r"(hp[1]), "r"(hp[2]), "r"(hp[3]) : "memory");
            }
            SGP_R16_ST16(lane_addr + ALO_OFF + c * 32 + (ko >> 1), lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // new state (and the staged output) -> async proxy
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) r16_mbar_arrive(&a_ready[c]);
        };
        // ---- initial state ----------------------------------------------------------------------
#pragma unroll 1
        for (int hh = 0; hh < NH; ++hh) {
            const int c = 2 * hh + (cg >> 1), ko = (cg & 1) * 32;
            float hn[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live) v = *reinterpret_cast<const float4*>(h_state + (size_t)node * H + c * 64 + ko + j);
                hn[j] = v.x; hn[j + 1] = v.y; hn[j + 2] = v.z; hn[j + 3] = v.w;
            }
            publish(c, ko, hn);
        }

        float xr[FINP];
#pragma unroll
        for (int f = 0; f < FINP; ++f)
            xr[f] = (live && f < Fin) ? __ldg(x + (size_t)node * x_ns + f) : 0.f;
        bool ok = true;
        bool store_pending = false;
        double csum = 0.0;
        const float oma_s = oma * (1.f / kR16StateScale);      // exact: a power of two
        // tanh(z) = 1 - 2 / (2^(2 z log2 e) + 1): the factor 2 log2 e is folded into inv_scale, the bias and W_ih
        // (z arrives as the ex2 argument), and alpha * tanh = alpha - 2 alpha r is one FMA
        const float inv_c = inv_scale * kR16TanhC, m2a = -2.f * alpha;
        for (int t = 0; t < Tc && ok; ++t) {
            float xn[FINP];
#pragma unroll
            for (int f = 0; f < FINP; ++f)
                xn[f] = (live && f < Fin && t + 1 < Tc) ? __ldg(x + (size_t)(t + 1) * x_ts + (size_t)node * x_ns + f) : 0.f;
#pragma unroll 1
            for (int hh = 0; hh < NH; ++hh) {
                const int c = 2 * hh + (cg >> 1), ko = (cg & 1) * 32, c0 = c * 64 + ko;     // my 32 state columns
                if (!r16_wait(&acc_ready[hh], t & 1, &abort_s, err, lane)) { ok = false; break; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t d[32], lo_old[16];
                SGP_R16_LD32(lane_addr + c0, d);
                SGP_R16_LD16(lane_addr + ALO_OFF + c * 32 + (ko >> 1), lo_old);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) r16_mbar_arrive(&d_free[hh]);          // my part of D half hh is in registers
                float hn[32];
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    uint32_t hp[4];                                   // old state hi: still in the A tile
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(hp[0]), "=r"(hp[1]), "=r"(hp[2]), "=r"(hp[3]) : "r"(a_base + a16_offset(m, c0 + j)));
#pragma unroll
                    for (int e = 0; e < 8; e += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + j + e);
                        float z[4] = {fmaf(__uint_as_float(d[j + e]), inv_c, b4.x), fmaf(__uint_as_float(d[j + e + 1]), inv_c, b4.y),
                                      fmaf(__uint_as_float(d[j + e + 2]), inv_c, b4.z), fmaf(__uint_as_float(d[j + e + 3]), inv_c, b4.w)};
#pragma unroll
                        for (int f = 0; f < FINP; ++f) {
                            const float4 w4 = *reinterpret_cast<const float4*>(wih_s + f * H + c0 + j + e);
                            z[0] = fmaf(xr[f], w4.x, z[0]); z[1] = fmaf(xr[f], w4.y, z[1]);
                            z[2] = fmaf(xr[f], w4.z, z[2]); z[3] = fmaf(xr[f], w4.w, z[3]);
                        }
#pragma unroll
                        for (int p2 = 0; p2 < 4; p2 += 2) {
                            // old state = (hi + lo) / 2^14, folded into the blend: oma_s = (1 - alpha) / 2^14
                            const float2 hf = __half22float2(r16_half2(hp[(e + p2) >> 1]));
                            const float2 lf = __half22float2(r16_half2(lo_old[(j + e + p2) >> 1]));
                            hn[j + e + p2] = fmaf(oma_s, hf.x, fmaf(oma_s, lf.x, fmaf(m2a, r16_rcp_ex2p1(z[p2]), alpha)));
                            hn[j + e + p2 + 1] = fmaf(oma_s, hf.y, fmaf(oma_s, lf.y, fmaf(m2a, r16_rcp_ex2p1(z[p2 + 1]), alpha)));
                        }
                    }
                }
                if (live) {
                    float part = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) part += hn[j];
                    csum += (double)part;
                }
                // output block [32 nodes x 32 columns] -> my staging tile -> one TMA tensor store
                if (store_pending) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};"
                                 :: "r"(my_out + lane * 128 + (((j >> 2) ^ (lane & 7)) << 4)), "f"(hn[j]), "f"(hn[j + 1]),
                                    "f"(hn[j + 2]), "f"(hn[j + 3]) : "memory");
                // the MMAs of this step that read the old chunk: done for the last half when its
                // acc_ready fired; signalled per chunk (a_free) for the halves before it
                if (hh < NH - 1) {
                    if (!r16_wait(&a_free[c], t & 1, &abort_s, err, lane)) { ok = false; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                publish(c, ko, hn);                                   // (its proxy fence + __syncwarp cover the staging tile)
                if (lane == 0) {
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                                 :: "l"(&out_map), "r"(c0), "r"(n0 + q4 * 32), "r"(t), "r"(my_out) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                store_pending = true;
                if (t == Tc - 1 && live) {                            // carry the state: exactly what was written out
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(h_state + (size_t)node * H + c0 + j) = make_float4(hn[j], hn[j + 1], hn[j + 2], hn[j + 3]);
                }
            }
#pragma unroll
            for (int f = 0; f < FINP; ++f) xr[f] = xn[f];
        }
        if (store_pending && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        if (chk != nullptr && ok) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
            if (lane == 0) atomicAdd(chk, csum);
        }
    } else {
        // ================= MMA issuer ==============================================================
        // kind::f16 (fp16 x fp16 -> fp32), A K-major (smem hi / TMEM lo), B K-major, N = 128, M = 128, K = 16
        constexpr uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) |
                                   ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t hi32 = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO 1024, version, SW128
        constexpr uint32_t lo32 = (16u >> 4) << 16;                            // LBO
        auto desc = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
        auto mma_ss = [](uint32_t d, uint64_t da, uint64_t db, uint32_t idesc_, uint32_t acc) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                         :: "r"(d), "l"(da), "l"(db), "r"(idesc_), "r"(acc) : "memory");
        };
        auto mma_ts = [](uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc_, uint32_t acc) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                         :: "r"(d), "r"(a_tmem), "l"(db), "r"(idesc_), "r"(acc) : "memory");
        };
        const uint32_t a0 = lo32 | (a_base >> 4), w0 = lo32 | (w_base >> 4);
        bool ok = true;
        int s = 0, ph = 0;
        for (int t = 0; t < Tc && ok; ++t) {
#pragma unroll 1
            for (int hh = 0; hh < NH && ok; ++hh) {
                const uint32_t dcol = tmem_d + hh * 128;
                if (t > 0 && !r16_wait(&d_free[hh], (t - 1) & 1, &abort_s, err, lane)) { ok = false; break; }
#pragma unroll 1
                for (int c = 0; c < NC; ++c) {
                    if (hh == 0 && !r16_wait(&a_ready[c], t & 1, &abort_s, err, lane)) { ok = false; break; }
                    // hi image: Ah x Wh and Al x Wh; lo image: Ah x Wl
                    if (!r16_wait(&wfull[s], ph & 1, &abort_s, err, lane)) { ok = false; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (r16_elect_one()) {
                        const uint32_t wh = w0 + s * (kR16WStage >> 4);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t da = desc(a0 + c * (16384 >> 4) + ks * 2, hi32);
                            const uint64_t db = desc(wh + ks * 2, hi32);
                            mma_ss(dcol, da, db, idesc, (c | ks) ? 1u : 0u);
                            mma_ts(dcol, tmem_d + ALO_OFF + c * 32 + ks * 8, db, idesc, 1u);
                        }
                        r16_commit(&wempty[s]);
                    }
                    __syncwarp();
                    if (++s == kR16WStages) { s = 0; ++ph; }
                    if (!r16_wait(&wfull[s], ph & 1, &abort_s, err, lane)) { ok = false; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (r16_elect_one()) {
                        const uint32_t wl = w0 + s * (kR16WStage >> 4);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t da = desc(a0 + c * (16384 >> 4) + ks * 2, hi32);
                            const uint64_t db = desc(wl + ks * 2, hi32);
                            mma_ss(dcol, da, db, idesc, 1u);
                        }
                        r16_commit(&wempty[s]);
                        if (hh == NH - 1 && c < NFREE) r16_commit(&a_free[c]);       // old state chunk c is dead
                        if (c == NC - 1) r16_commit(&acc_ready[hh]);                 // pre-activations of half hh complete
                    }
                    __syncwarp();
                    if (++s == kR16WStages) { s = 0; ++ph; }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "r"(TMEM_COLS));
}

// W_hh [H, H] * scale -> fp16 hi / lo images [H/128 halves][H/64 chunks][hi | lo][128 n x 64 k, K-major
// SWIZZLE_128B], i.e. in the order the scan consumes them
__global__ void reservoir_tc16_pack_kernel(const float* __restrict__ w_hh, int H, float scale, __half* __restrict__ wimg) {
    const int NC = H / 64;
    const int64_t total = (int64_t)H * H;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / H), k = (int)(i % H);
        const float w = w_hh[i] * scale;
        const __half hi = __float2half_rn(w);
        const __half lo = __float2half_rn(w - __half2float(hi));
        const int c = k >> 6, kk = k & 63, hh = n >> 7, nn = n & 127;
        const size_t img = ((size_t)(hh * NC + c) * 2) * (128 * 64);
        const int off = (nn >> 3) * 512 + (nn & 7) * 64 + ((((kk >> 3) ^ (nn & 7))) << 3) + (kk & 7);     // halves
        wimg[img + off] = hi;
        wimg[img + 128 * 64 + off] = lo;
    }
}

}  // namespace sgp

using namespace sgp;

typedef CUresult (*PFN_encodeTiled16)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" int sgp_reservoir_tc16_pack(const float* w_hh, int H, float w_scale, void* wimg, void* stream) {
    SGP_REQUIRE(w_hh && wimg, SGP_EINVAL, "sgp_reservoir_tc16_pack: null pointer");
    SGP_REQUIRE(H == 128 || H == 256, SGP_EUNSUPPORTED, "sgp_reservoir_tc16_pack: H=%d (128 or 256)", H);
    SGP_REQUIRE(w_scale > 0.f, SGP_EINVAL, "sgp_reservoir_tc16_pack: w_scale=%g", (double)w_scale);
    reservoir_tc16_pack_kernel<<<(H * H + 255) / 256, 256, 0, as_stream(stream)>>>(w_hh, H, w_scale,
                                                                                    reinterpret_cast<__half*>(wimg));
    SGP_LAUNCH_CHECK("reservoir_tc16_pack");
    return SGP_OK;
}

extern "C" int sgp_reservoir_scan_tc16(const float* x, int64_t x_t_stride, int64_t x_n_stride, int Fin,
                                       const void* wimg, float w_scale, const float* w_ih, const float* bias,
                                       float alpha, float one_minus_alpha, float* h_state, float* out,
                                       int64_t out_t_stride, int64_t out_n_stride, int Tc, int N, int H,
                                       int* err_flag, double* checksum, void* stream) {
    SGP_REQUIRE(x && wimg && w_ih && bias && h_state && out && err_flag, SGP_EINVAL,
                "sgp_reservoir_scan_tc16: null pointer");
    SGP_REQUIRE(H == 128 || H == 256, SGP_EUNSUPPORTED, "sgp_reservoir_scan_tc16: H=%d (128 or 256)", H);
    SGP_REQUIRE(Fin >= 1 && Fin <= kR16MaxFin, SGP_EUNSUPPORTED, "sgp_reservoir_scan_tc16: Fin=%d (1..%d)", Fin, kR16MaxFin);
    SGP_REQUIRE(w_scale > 0.f, SGP_EINVAL, "sgp_reservoir_scan_tc16: w_scale=%g", (double)w_scale);
    SGP_REQUIRE(aligned16(out) && aligned16(h_state) && aligned16(wimg) && out_t_stride % 4 == 0 &&
                    out_n_stride % 4 == 0, SGP_EALIGN, "sgp_reservoir_scan_tc16: views must be 16-byte aligned");
    if (N == 0 || Tc == 0) return SGP_OK;
    const size_t smem = (size_t)128 * H * 2 + (size_t)kR16WStages * kR16WStage + (size_t)kR16EpiWarps * kR16OutTile +
                        (size_t)(kR16MaxFin + 1) * H * 4 + 1024;
    const int grid = (N + 127) / 128;
    static PFN_encodeTiled16 encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        SGP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        SGP_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, SGP_ECUDA, "sgp_reservoir_scan_tc16: cuTensorMapEncodeTiled not available");
        encode = reinterpret_cast<PFN_encodeTiled16>(fn);
    }
    // output view [Tc][N][H] as a 3-D tensor map, box = one epilogue warp's block [1][32 nodes][32 columns]
    CUtensorMap out_map;
    {
        const cuuint64_t dims[3] = {(cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)Tc};
        const cuuint64_t strides[2] = {(cuuint64_t)out_n_stride * 4, (cuuint64_t)(Tc > 1 ? out_t_stride : out_n_stride * (int64_t)N) * 4};
        const cuuint32_t box[3] = {32, 32, 1}, estr[3] = {1, 1, 1};
        const CUresult r = encode(&out_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SGP_REQUIRE(r == CUDA_SUCCESS, SGP_EINVAL, "sgp_reservoir_scan_tc16: cuTensorMapEncodeTiled failed (%d): out strides %lld / %lld",
                    (int)r, (long long)out_t_stride, (long long)out_n_stride);
    }
    const float inv_scale = 1.f / (kR16StateScale * w_scale);
#define SGP_R16(H_, F_)                                                                                 \
    do {                                                                                                \
        SGP_CUDA(cudaFuncSetAttribute(reservoir_tc16_kernel<H_, F_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        reservoir_tc16_kernel<H_, F_><<<grid, kR16Threads, smem, as_stream(stream)>>>(                  \
            x, x_t_stride, x_n_stride, Fin, reinterpret_cast<const __half*>(wimg), w_ih, bias, alpha,   \
            one_minus_alpha, inv_scale, h_state, out_map, Tc, N, err_flag, checksum);                   \
    } while (0)
#define SGP_R16_F(H_)                                                                                   \
    do {                                                                                                \
        if (Fin == 1) SGP_R16(H_, 1); else if (Fin == 2) SGP_R16(H_, 2);                                \
        else if (Fin <= 4) SGP_R16(H_, 4); else SGP_R16(H_, 8);                                         \
    } while (0)
    if (H == 256) SGP_R16_F(256); else SGP_R16_F(128);
#undef SGP_R16_F
#undef SGP_R16
    SGP_LAUNCH_CHECK("reservoir_scan_tc16");
    return SGP_OK;
}
