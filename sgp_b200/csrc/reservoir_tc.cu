// K1-TC — the leaky-ESN scan on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-accurate.
//
// Same contract as sgp_reservoir_scan (one layer, Tc steps, state carried; replaces the Python
// loop of lib/nn/reservoir/reservoir.py:158-186 around :77-81) for H in {128, 256} and Fin <= 8.
// Per CTA: a tile of 128 nodes marches over time; per step
//     D[m, n] = sum_k h[m, k] * W_hh[n, k]              m: 128 nodes (TMEM lanes), n, k: H
// is a dense [128 x H] x [H x H] GEMM issued as 3xTF32 tcgen05.mma (a 1000-step
// recurrence needs fp32-level accuracy; plain TF32 is 8.7e-4 off, the 3-pass split 1.4e-6):
//   * A hi  = the state itself, fp32, K-major SWIZZLE_128B in shared memory (the tensor core
//     ignores the low 13 mantissa bits: measured, tools/microbench/umma_tf32_test.cu);
//   * A lo  = h - tf32(h), resident in TMEM (TS-mode MMA), written with tcgen05.st;
//   * B     = W_hh, split hi / lo at pack time into [128 n x 32 k] K-major SWIZZLE_128B images
//     (sgp_reservoir_tc_pack) that stream from L2 through a 5-stage ring of 16 KB bulk copies every step — W_hh in
//     two tf32 parts is 512 KB at H = 256, more than two SMs' shared memory;
//   * D     = fp32 accumulators in TMEM ([128 lanes x H columns]); TMEM is exactly full at H = 256.
// The state tile (128 KB at H = 256) and TMEM admit one tile per SM and no second buffer, so the
// step is pipelined INSIDE the tile, by output-column half hh (128 columns) and k-chunk c (32
// state columns):
//   MMA order  [hh = 0: c = 0..NC-1] [hh = 1: c = 0..NC-1]; after (last half, c) the old state
//              chunk c is dead (a_free[c]), after the last chunk of half hh its pre-activations
//              are complete (acc_ready[hh]);
//   epilogue   16 warps = 4 TMEM lane quarters x 4 column groups; per half a warp owns the 32
//              columns of chunk c = 4 hh + cg: tcgen05.ld them, add bias and x_t W_ih^T, activation,
//              leaky blend with the old state (read from the A-hi tile), store the row piece to
//              the encoder output, release the D half (d_free[hh]), then — once a_free[c] says
//              the MMAs are done with the old chunk — write the new state chunk (fp32 to the A-hi
//              tile, lo part to TMEM) and publish it (a_ready[c]);
//   next step  MMA (t+1, hh = 0, c) starts as soon as a_ready[c] and d_free[0] allow, i.e. while
//              the epilogue is still working on the second half of step t.
// So half 0's epilogue runs under half 1's MMAs and half 1's epilogue under the next step's first
// MMAs; the tensor pipe (192 N=128 MMAs = 12.3k cycles per step at H = 256) is the bound.
// Warp roles: 0-15 epilogue, 16 W producer (TMA bulk copies), 17 MMA issuer; all hand-offs through mbarriers.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace sgp {

constexpr int kRtEpiWarps = 16, kRtProdWarps = 1;     // 18 warps: 96 registers per thread
constexpr int kRtThreads = (kRtEpiWarps + kRtProdWarps + 1) * 32;
constexpr int kRtWStage = 128 * 32 * 4;      // 16 KB: one [128 n x 32 k] image (hi OR lo)
constexpr int kRtWStages = 5;
constexpr int kRtMaxFin = 8;

__device__ __forceinline__ uint32_t rt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rt_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(rt_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void rt_mbar_arrive(uint64_t* bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(rt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void rt_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(rt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool rt_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
// bounded warp-wide wait (see spmm_tc.cu)
__device__ __forceinline__ bool rt_wait(uint64_t* bar, uint32_t parity, volatile int* abort_s, int* err, int lane) {
    const uint32_t a = rt_smem_u32(bar);
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

// tanh with 1e-7 absolute error from two MUFU ops: 1 - 2 / (1 + 2^(2 log2(e) x)).  No clamp is
// needed: 2^y overflows to +inf, whose approximate reciprocal is +0 (-> 1), and underflows to 0
// (-> -1).  (tanhf costs ~20 instructions; the epilogue evaluates 32k per 128-node step.)
__device__ __forceinline__ float fast_tanh(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));     // 2 log2(e) x
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    return fmaf(-2.f, r, 1.f);
}

#define SGP_RT_LD32(addr, v)                                                                         \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15," \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"                       \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),   \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),          \
                   "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),        \
                   "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),        \
                   "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                                                              \
                 : "r"(addr))
#define SGP_RT_ST32(addr, v)                                                                         \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,"  \
                 "%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"       \
                 :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),   \
                    "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]),           \
                    "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),        \
                    "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),        \
                    "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory")

#define SGP_RT_LD16(addr, v)                                                                         \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),   \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) \
                 : "r"(addr))
#define SGP_RT_ST16(addr, v)                                                                         \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" \
                 :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),   \
                    "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]),           \
                    "r"(v[14]), "r"(v[15]) : "memory")

// byte offset of element (row m, column k) of the K-major SWIZZLE_128B state tile [128 x H]:
// k-blocks of 32 columns (16 KB each), atoms of 8 rows x 128 B, 16-byte chunk XOR (row % 8)
__device__ __forceinline__ uint32_t a_tile_offset(int m, int k) {
    return (uint32_t)((k >> 5) * 16384 + (m >> 3) * 1024 + (m & 7) * 128 + ((((k & 31) >> 2) ^ (m & 7)) << 4) + (k & 3) * 4);
}

template <int ACT>
__device__ __forceinline__ float rt_activate(float z) {
    if (ACT == SGP_ACT_TANH) return fast_tanh(z);
    if (ACT == SGP_ACT_RELU) return fmaxf(z, 0.f);
    return z;
}
__device__ __forceinline__ uint32_t rt_lo_part(float v) {
    return __float_as_uint(v - __uint_as_float(__float_as_uint(v) & 0xffffe000u));
}

template <int H, int FINP, int ACT>
__global__ void __launch_bounds__(kRtThreads, 1)
reservoir_tc_kernel(const float* __restrict__ x, int64_t x_ts, int64_t x_ns, int Fin,
                    const float* __restrict__ wimg /* [H/32][H/128][2][128*32] */,
                    const float* __restrict__ w_ih /* [H, Fin] */, const float* __restrict__ bias,
                    float alpha, float oma,
                    float* __restrict__ h_state, const __grid_constant__ CUtensorMap out_map,
                    int Tc, int N, int* err, double* __restrict__ chk, long long* trace) {
    constexpr int NH = H / 128;                 // output-column halves (MMA N = 128)
    constexpr int NC = H / 32;                  // k-chunks of 32
    constexpr int A_BYTES = 128 * H * 4;        // state tile
    constexpr int kStagesPerStep = NC * NH * 2; // (half, chunk, hi|lo) images per step
    constexpr int TMEM_COLS = (2 * H <= 256) ? 256 : 512;
    constexpr int ALO_OFF = H;                  // TMEM columns [0,H) = D, [H,2H) = A lo
    constexpr int NFREE = (NH - 1) * 4;         // chunks whose release is signalled before the step ends
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (rt_smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - rt_smem_u32(smem_raw));
    // [A hi tile | W ring | w_ih (FINP x H) | bias (H)]
    const uint32_t a_base = smem_base, w_base = smem_base + A_BYTES;
    float* wih_s = reinterpret_cast<float*>(smem + A_BYTES + kRtWStages * kRtWStage);
    float* bias_s = wih_s + kRtMaxFin * H;
    __shared__ uint64_t wfull[kRtWStages], wempty[kRtWStages];
    __shared__ uint64_t acc_ready[2], d_free[2], a_ready[8], a_free[4];
    __shared__ uint32_t tmem_base_s;
    __shared__ volatile int abort_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * 128;
    // optional per-step timestamps of CTA 7 (tools/trace_rt.py); trace == nullptr in production
    // (compiled in only with -DSGP_RT_TRACE_ON: the MMA warp's loop is latency-bound, every stamp costs)
#ifdef SGP_RT_TRACE_ON
    const bool tr = trace && blockIdx.x == 7;
#define SGP_RT_TRACE(role, t_, v) do { if (tr && lane == 0 && (t_) < 64) trace[(role) * 64 + (t_)] = (v); } while (0)
#else
    constexpr bool tr = false;
#define SGP_RT_TRACE(role, t_, v) do { } while (0)
#endif

    if (tid == 0) {
        abort_s = 0;
        for (int s = 0; s < kRtWStages; ++s) {
            rt_mbar_init(&wfull[s], 1);           // the producer's arrive.expect_tx; the copy completes the bytes
            rt_mbar_init(&wempty[s], 1);
        }
        for (int i = 0; i < 2; ++i) {
            rt_mbar_init(&acc_ready[i], 1);
            rt_mbar_init(&d_free[i], kRtEpiWarps);
        }
        for (int i = 0; i < 8; ++i) rt_mbar_init(&a_ready[i], 4);
        for (int i = 0; i < 4; ++i) rt_mbar_init(&a_free[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < FINP * H; i += kRtThreads) {
        const int f = i / H, n = i % H;
        wih_s[i] = (f < Fin) ? w_ih[(size_t)n * Fin + f] : 0.f;      // [f][n], zero rows above Fin
    }
    for (int i = tid; i < H; i += kRtThreads) bias_s[i] = bias[i];
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(rt_smem_u32(&tmem_base_s)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;

    if (warp == kRtEpiWarps) {
        // ================= producer warp: stream the W images, kStagesPerStep per time step ====
        // One 16 KB bulk copy (TMA, no tensor map: the images are stored pre-swizzled) per ring
        // stage, completion counted in bytes on wfull[s]; the whole ring stays in flight.  Image
        // order of a step = the MMA order [half][chunk][hi | lo].
        const long long total = (long long)Tc * kStagesPerStep;
        int s = 0, ph = 0, js = 0;
        for (long long j = 0; j < total; ++j) {
            if (ph > 0 && !rt_wait(&wempty[s], (ph - 1) & 1, &abort_s, err, lane)) break;
            if (rt_elect_one()) {
                const int hh = js / (NC * 2), c = (js >> 1) % NC, part = js & 1;
                const float* src = wimg + (size_t)((c * NH + hh) * 2 + part) * (kRtWStage / 4);
                const uint32_t bar = rt_smem_u32(&wfull[s]);
                asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}\n"
                             :: "r"(bar), "r"(kRtWStage) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(w_base + s * kRtWStage), "l"(src), "r"(kRtWStage), "r"(bar) : "memory");
            }
            __syncwarp();
            if (++s == kRtWStages) { s = 0; ++ph; }
            if (++js == kStagesPerStep) js = 0;
        }
    } else if (warp < kRtEpiWarps) {
        // ================= epilogue warps: thread = node, warp = (lane quarter, column group) ==
        const int q4 = warp & 3, cg = warp >> 2;
        const int m = q4 * 32 + lane, node = n0 + m;
        const bool live = node < N;
        const uint32_t lane_addr = tmem_d + ((uint32_t)(q4 * 32) << 16);
        // new state chunk -> A hi tile (smem) + A lo (TMEM), then publish it to the MMA warp
        auto publish = [&](int c, const float (&hn)[32]) {
            uint32_t lo[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};"
                             :: "r"(a_base + a_tile_offset(m, c * 32 + j)), "f"(hn[j]), "f"(hn[j + 1]),
                                "f"(hn[j + 2]), "f"(hn[j + 3]) : "memory");
                lo[j] = rt_lo_part(hn[j]);
                lo[j + 1] = rt_lo_part(hn[j + 1]);
                lo[j + 2] = rt_lo_part(hn[j + 2]);
                lo[j + 3] = rt_lo_part(hn[j + 3]);
            }
            SGP_RT_ST32(lane_addr + ALO_OFF + c * 32, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // new state -> tensor proxy
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) rt_mbar_arrive(&a_ready[c]);
        };
        // ---- initial state ----------------------------------------------------------------------
#pragma unroll 1
        for (int hh = 0; hh < NH; ++hh) {
            const int c = hh * 4 + cg;
            float hn[32];
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live) v = *reinterpret_cast<const float4*>(h_state + (size_t)node * H + c * 32 + j);
                hn[j] = v.x; hn[j + 1] = v.y; hn[j + 2] = v.z; hn[j + 3] = v.w;
            }
            publish(c, hn);
        }

        float xr[FINP];
#pragma unroll
        for (int f = 0; f < FINP; ++f)
            xr[f] = (live && f < Fin) ? __ldg(x + (size_t)node * x_ns + f) : 0.f;
        bool ok = true;
        double csum = 0.0;          // fused sink: sum of every state value this thread produces (`checksum`)
        for (int t = 0; t < Tc && ok; ++t) {
            float xn[FINP];                           // x_{t+1}: in flight while this step is finished
#pragma unroll
            for (int f = 0; f < FINP; ++f)
                xn[f] = (live && f < Fin && t + 1 < Tc) ? __ldg(x + (size_t)(t + 1) * x_ts + (size_t)node * x_ns + f) : 0.f;
#pragma unroll 1
            for (int hh = 0; hh < NH; ++hh) {
                const int c = hh * 4 + cg, c0 = c * 32;
                if (!rt_wait(&acc_ready[hh], t & 1, &abort_s, err, lane)) { ok = false; break; }
                if (warp == 0) SGP_RT_TRACE(6 + 3 * hh, t, clock64());
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t d[32];
                SGP_RT_LD32(lane_addr + c0, d);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) rt_mbar_arrive(&d_free[hh]);          // my part of D half hh is in registers
                float hn[32];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + j);
                    float z[4] = {__uint_as_float(d[j]) + b4.x, __uint_as_float(d[j + 1]) + b4.y,
                                  __uint_as_float(d[j + 2]) + b4.z, __uint_as_float(d[j + 3]) + b4.w};
#pragma unroll
                    for (int f = 0; f < FINP; ++f) {
                        const float4 w4 = *reinterpret_cast<const float4*>(wih_s + f * H + c0 + j);
                        z[0] = fmaf(xr[f], w4.x, z[0]); z[1] = fmaf(xr[f], w4.y, z[1]);
                        z[2] = fmaf(xr[f], w4.z, z[2]); z[3] = fmaf(xr[f], w4.w, z[3]);
                    }
                    float4 ho;                        // old state: still in the A-hi tile (chunk c not yet rewritten)
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                 : "=f"(ho.x), "=f"(ho.y), "=f"(ho.z), "=f"(ho.w) : "r"(a_base + a_tile_offset(m, c0 + j)));
                    float4 hv;
                    hv.x = fmaf(alpha, rt_activate<ACT>(z[0]), oma * ho.x);
                    hv.y = fmaf(alpha, rt_activate<ACT>(z[1]), oma * ho.y);
                    hv.z = fmaf(alpha, rt_activate<ACT>(z[2]), oma * ho.z);
                    hv.w = fmaf(alpha, rt_activate<ACT>(z[3]), oma * ho.w);
                    hn[j] = hv.x; hn[j + 1] = hv.y; hn[j + 2] = hv.z; hn[j + 3] = hv.w;
                }
                if (live) {
                    float part = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) part += hn[j];
                    csum += (double)part;
                }
                // the MMAs of this step that read the old chunk c: done for the last half when its
                // acc_ready fired; signalled per chunk (a_free) for the halves before it
                if (warp == 0) SGP_RT_TRACE(7 + 3 * hh, t, clock64());
                if (hh < NH - 1) {
                    if (!rt_wait(&a_free[c], t & 1, &abort_s, err, lane)) { ok = false; break; }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                publish(c, hn);
                if (warp == 0) SGP_RT_TRACE(8 + 3 * hh, t, clock64());
            }
#pragma unroll
            for (int f = 0; f < FINP; ++f) xr[f] = xn[f];
        }
        if (chk != nullptr && ok) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o);
            if (lane == 0) atomicAdd(chk, csum);
        }
        // ---- carry the state (each warp re-reads the chunks it wrote) ----------------------------
        if (ok && live) {
#pragma unroll 1
            for (int hh = 0; hh < NH; ++hh) {
                const int c0 = (hh * 4 + cg) * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 v;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a_base + a_tile_offset(m, c0 + j)));
                    *reinterpret_cast<float4*>(h_state + (size_t)node * H + c0 + j) = v;
                }
            }
        }
    } else {
        // ================= MMA issuer ==============================================================
        // kind::tf32, fp32 accumulate, A K-major (smem hi / TMEM lo), B K-major, N = 128, M = 128
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) |
                                   ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t hi32 = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO 1024, version, SW128
        constexpr uint32_t lo32 = (16u >> 4) << 16;                            // LBO
        auto desc = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
        auto mma_ss = [](uint32_t d, uint64_t da, uint64_t db, uint32_t idesc_, uint32_t acc) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                         :: "r"(d), "l"(da), "l"(db), "r"(idesc_), "r"(acc) : "memory");
        };
        auto mma_ts = [](uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc_, uint32_t acc) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                         :: "r"(d), "r"(a_tmem), "l"(db), "r"(idesc_), "r"(acc) : "memory");
        };
        const uint32_t a0 = lo32 | (a_base >> 4), w0 = lo32 | (w_base >> 4);
        auto store_chunk = [&](int c, int t_out) {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                         :: "l"(&out_map), "r"(c * 32), "r"(n0), "r"(t_out), "r"(a_base + c * 16384) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        };
        bool ok = true;
        int s = 0, ph = 0;
        for (int t = 0; t < Tc && ok; ++t) {
            long long w_wait = 0, s_wait = 0, t0 = 0;
            SGP_RT_TRACE(0, t, clock64());
#pragma unroll 1
            for (int hh = 0; hh < NH && ok; ++hh) {
                const uint32_t dcol = tmem_d + hh * 128;
                // D half hh of the previous step has been read out
                if (tr) t0 = clock64();
                if (t > 0 && !rt_wait(&d_free[hh], (t - 1) & 1, &abort_s, err, lane)) { ok = false; break; }
                if (tr) s_wait += clock64() - t0;
#pragma unroll 1
                for (int c = 0; c < NC; ++c) {
                    // state chunk c of step t-1 in place (the later halves read what half 0 waited for)
                    if (tr) t0 = clock64();
                    if (hh == 0 && !rt_wait(&a_ready[c], t & 1, &abort_s, err, lane)) { ok = false; break; }
                    // ... and complete (all four lane quarters): one TMA tensor store sends the
                    // [128 nodes x 32 columns] chunk of step t-1 from the A-hi tile to the output
                    if (hh == 0 && t > 0 && rt_elect_one()) store_chunk(c, t - 1);
                    __syncwarp();
                    if (tr) { const long long t1 = clock64(); s_wait += t1 - t0; t0 = t1; }
                    // hi image: Ah x Wh and Al x Wh; lo image: Ah x Wl
                    if (!rt_wait(&wfull[s], ph & 1, &abort_s, err, lane)) { ok = false; break; }
                    if (tr) w_wait += clock64() - t0;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (rt_elect_one()) {
                        const uint32_t wh = w0 + s * (kRtWStage >> 4);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t da = desc(a0 + c * (16384 >> 4) + ks * 2, hi32);
                            const uint64_t db = desc(wh + ks * 2, hi32);
                            mma_ss(dcol, da, db, idesc, (c | ks) ? 1u : 0u);
                            mma_ts(dcol, tmem_d + ALO_OFF + c * 32 + ks * 8, db, idesc, 1u);
                        }
                        rt_commit(&wempty[s]);
                    }
                    __syncwarp();
                    if (++s == kRtWStages) { s = 0; ++ph; }
                    if (tr) t0 = clock64();
                    if (!rt_wait(&wfull[s], ph & 1, &abort_s, err, lane)) { ok = false; break; }
                    if (tr) w_wait += clock64() - t0;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (rt_elect_one()) {
                        const uint32_t wl = w0 + s * (kRtWStage >> 4);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t da = desc(a0 + c * (16384 >> 4) + ks * 2, hi32);
                            const uint64_t db = desc(wl + ks * 2, hi32);
                            mma_ss(dcol, da, db, idesc, 1u);
                        }
                        rt_commit(&wempty[s]);
                        // the epilogue may overwrite state chunks once these fire: the output
                        // stores must have read them (issued >10k cycles ago)
                        if (hh == NH - 1 && (c < NFREE || c == NC - 1))
                            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        if (hh == NH - 1 && c < NFREE) rt_commit(&a_free[c]);       // old state chunk c is dead
                        if (c == NC - 1) rt_commit(&acc_ready[hh]);                 // pre-activations of half hh complete
                    }
                    __syncwarp();
                    if (++s == kRtWStages) { s = 0; ++ph; }
                }
                SGP_RT_TRACE(1 + hh, t, clock64());
            }
            SGP_RT_TRACE(3, t, w_wait);
            SGP_RT_TRACE(4, t, s_wait);
        }
        // the last step's state -> output
#pragma unroll 1
        for (int c = 0; c < NC && ok; ++c) {
            if (!rt_wait(&a_ready[c], Tc & 1, &abort_s, err, lane)) { ok = false; break; }
            if (rt_elect_one()) store_chunk(c, Tc - 1);
            __syncwarp();
        }
        if (rt_elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
    }
#undef SGP_RT_TRACE
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "r"(TMEM_COLS));
}

// W_hh [H, H] -> images [H/32 chunks][H/128 halves][hi | lo][128 n x 32 k, K-major SWIZZLE_128B]
__global__ void reservoir_tc_pack_kernel(const float* __restrict__ w_hh, int H, float* __restrict__ wimg) {
    const int NH = H / 128;
    const int64_t total = (int64_t)H * H;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / H), k = (int)(i % H);
        const float w = w_hh[i];
        const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
        const int c = k >> 5, kk = k & 31, hh = n >> 7, nn = n & 127;
        const size_t img = ((size_t)(c * NH + hh) * 2) * (128 * 32);
        const int off = (nn >> 3) * 256 + (nn & 7) * 32 + (((kk >> 2) ^ (nn & 7)) << 2) + (kk & 3);    // floats
        wimg[img + off] = hi;
        wimg[img + 128 * 32 + off] = w - hi;
    }
}

}  // namespace sgp

using namespace sgp;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" int sgp_reservoir_tc_pack(const float* w_hh, int H, float* wimg, void* stream) {
    SGP_REQUIRE(w_hh && wimg, SGP_EINVAL, "sgp_reservoir_tc_pack: null pointer");
    SGP_REQUIRE(H == 128 || H == 256, SGP_EUNSUPPORTED, "sgp_reservoir_tc_pack: H=%d (128 or 256)", H);
    reservoir_tc_pack_kernel<<<(H * H + 255) / 256, 256, 0, as_stream(stream)>>>(w_hh, H, wimg);
    SGP_LAUNCH_CHECK("reservoir_tc_pack");
    return SGP_OK;
}

extern "C" int sgp_reservoir_scan_tc(const float* x, int64_t x_t_stride, int64_t x_n_stride, int Fin,
                                     const float* wimg, const float* w_ih, const float* bias, float alpha,
                                     float one_minus_alpha, int act, float* h_state, float* out,
                                     int64_t out_t_stride, int64_t out_n_stride, int Tc, int N, int H,
                                     int* err_flag, double* checksum, void* stream) {
    SGP_REQUIRE(x && wimg && w_ih && bias && h_state && out && err_flag, SGP_EINVAL,
                "sgp_reservoir_scan_tc: null pointer");
    SGP_REQUIRE(H == 128 || H == 256, SGP_EUNSUPPORTED, "sgp_reservoir_scan_tc: H=%d (128 or 256)", H);
    SGP_REQUIRE(Fin >= 1 && Fin <= kRtMaxFin, SGP_EUNSUPPORTED, "sgp_reservoir_scan_tc: Fin=%d (1..%d)", Fin, kRtMaxFin);
    SGP_REQUIRE(act == SGP_ACT_TANH || act == SGP_ACT_RELU || act == SGP_ACT_IDENTITY, SGP_EUNSUPPORTED,
                "sgp_reservoir_scan_tc: activation %d not supported on this path", act);
    SGP_REQUIRE(aligned16(out) && aligned16(h_state) && aligned16(wimg) && out_t_stride % 4 == 0 &&
                    out_n_stride % 4 == 0, SGP_EALIGN, "sgp_reservoir_scan_tc: views must be 16-byte aligned");
    if (N == 0 || Tc == 0) return SGP_OK;
    const size_t smem = (size_t)128 * H * 4 + (size_t)kRtWStages * kRtWStage + (size_t)(kRtMaxFin + 1) * H * 4 + 1024;
    const int grid = (N + 127) / 128;
    // output view [Tc][N][H] (strides in elements) as a 3-D tensor map, box = one state chunk
    // [1][128 nodes][32 columns] in the SWIZZLE_128B layout the A-hi tile already has
    static PFN_encodeTiled encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        SGP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        SGP_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, SGP_ECUDA, "sgp_reservoir_scan_tc: cuTensorMapEncodeTiled not available");
        encode = reinterpret_cast<PFN_encodeTiled>(fn);
    }
    CUtensorMap out_map;
    {
        const cuuint64_t dims[3] = {(cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)Tc};
        // a dimension of size 1 takes any stride; keep it a valid multiple of 16 bytes
        const cuuint64_t strides[2] = {(cuuint64_t)out_n_stride * 4, (cuuint64_t)(Tc > 1 ? out_t_stride : out_n_stride * (int64_t)N) * 4};
        const cuuint32_t box[3] = {32, 128, 1}, estr[3] = {1, 1, 1};
        const CUresult r = encode(&out_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SGP_REQUIRE(r == CUDA_SUCCESS, SGP_EINVAL, "sgp_reservoir_scan_tc: cuTensorMapEncodeTiled failed (%d): out strides %lld / %lld",
                    (int)r, (long long)out_t_stride, (long long)out_n_stride);
    }
#ifdef SGP_RT_TRACE_ON
    // trace builds only (tools/trace_rt.py): device buffer for the per-step timestamps of one CTA
    long long* trace_ptr = getenv("SGP_B200_RT_TRACE") ? (long long*)strtoull(getenv("SGP_B200_RT_TRACE"), nullptr, 10) : nullptr;
#else
    long long* trace_ptr = nullptr;
#endif
#define SGP_RT(H_, F_, A_)                                                                              \
    do {                                                                                                \
        SGP_CUDA(cudaFuncSetAttribute(reservoir_tc_kernel<H_, F_, A_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        reservoir_tc_kernel<H_, F_, A_><<<grid, kRtThreads, smem, as_stream(stream)>>>(                 \
            x, x_t_stride, x_n_stride, Fin, wimg, w_ih, bias, alpha, one_minus_alpha, h_state, out_map, \
            Tc, N, err_flag, checksum, trace_ptr);                                                      \
    } while (0)
#define SGP_RT_F(H_, A_)                                                                                \
    do {                                                                                                \
        if (Fin == 1) SGP_RT(H_, 1, A_); else if (Fin == 2) SGP_RT(H_, 2, A_);                          \
        else if (Fin <= 4) SGP_RT(H_, 4, A_); else SGP_RT(H_, 8, A_);                                   \
    } while (0)
#define SGP_RT_A(H_)                                                                                    \
    do {                                                                                                \
        if (act == SGP_ACT_TANH) SGP_RT_F(H_, SGP_ACT_TANH);                                            \
        else if (act == SGP_ACT_RELU) SGP_RT_F(H_, SGP_ACT_RELU);                                       \
        else SGP_RT_F(H_, SGP_ACT_IDENTITY);                                                            \
    } while (0)
    if (H == 256) SGP_RT_A(256); else SGP_RT_A(128);
#undef SGP_RT_A
#undef SGP_RT_F
#undef SGP_RT
    SGP_LAUNCH_CHECK("reservoir_scan_tc");
    return SGP_OK;
}
