// K1 — fused leaky-ESN scan (one reservoir layer, a chunk of Tc time steps, all nodes).
//
// Replaces the Python loop of lib/nn/reservoir/reservoir.py:158-186 (Reservoir.forward) around
// :77-81 (ReservoirLayer.forward): two F.linear GEMMs + add + tanh + leaky blend + torch.stack per
// layer-step become ONE persistent kernel per layer and chunk.  Per node-step the state is read
// from shared memory, updated, written once to HBM (straight into its feature block of the
// concatenated encoder output) and never re-read by this kernel.
//
// Tiled kernel (H in {128, 256}), 256 threads = 8 warps, 1 CTA / SM:
//   * warp w owns TM nodes; its A rows  [ x_t (FinP) | h (H) ]  live in shared memory and are
//     private to the warp (only __syncwarp between steps);
//   * lane owns H/32 output columns as NQ = H/128 float4 groups (col = q*128 + 4*lane + j), so a
//     W^T row is read as conflict-free LDS.128 and the output row is written as coalesced STG.128;
//   * W^T ([FinP+H, H], k-major, made by sgp_reservoir_pack) streams from L2 through a 3-stage
//     cp.async ring of KC-row chunks shared by the CTA (one __syncthreads per chunk);
//   * the contraction runs on the fp32 pipe as packed FFMA2 (fma.rn.f32x2, scalar A operand
//     broadcast) — full fp32, no TF32: a 1000-step recurrence must stay within 1e-4 of the
//     reference (SURVEY.md §7 "hard parts").
// Bound: fp32 FMA.  flops / node-step = 2H(Fin+H) + ~6H;  HBM bytes / node-step = 4(Fin + H).
#include "common.cuh"

namespace sgp {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = 8;
constexpr int kKC = 32;      // W^T rows per pipeline chunk
constexpr int kStages = 3;
constexpr int kSmallFin = 8; // Fin <= 8: input projection done in registers, outside the W pipeline

// rows of W_ih^T in the packed weight matrix (zero padded): see sgp_reservoir_pack_rows()
__host__ __device__ inline int fin_padded(int Fin) {
    return Fin <= kSmallFin ? ((Fin + 3) & ~3) : ((Fin + kKC - 1) / kKC) * kKC;
}

__device__ __forceinline__ float activate(float v, int act) {
    if (act == SGP_ACT_TANH) return tanhf(v);
    if (act == SGP_ACT_RELU) return fmaxf(v, 0.f);
    return v;  // identity; self_norm is handled by the caller (needs the row norm)
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float f4_get(const float4& v, int i) {
    return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

// SMALL = true : Fin <= 8.  A row = h (H floats); x_t lives in a tiny per-warp slot, W_ih^T in
//                shared memory; the pipeline streams only W_hh^T (H rows, H/kKC chunks).
// SMALL = false: A row = [x_t (FinP) | h (H)], the pipeline streams all FinP + H rows
//                (deeper layers, Fin = H).
template <int H, int TM, bool SMALL>
__global__ void __launch_bounds__(kScanThreads, 1)
reservoir_scan_tiled(const float* __restrict__ x, int64_t x_ts, int64_t x_ns, int Fin, int FinP,
                     const float* __restrict__ wpack, const float* __restrict__ bias,
                     float alpha, float oma, int act,
                     float* __restrict__ h_state,
                     float* __restrict__ out, int64_t o_ts, int64_t o_ns,
                     int Tc, int N) {
    constexpr int NQ = H / 128;
    constexpr int BM = TM * kScanWarps;
    extern __shared__ __align__(16) float smem[];
    const int KA = SMALL ? H : FinP + H;     // contraction length that goes through the pipeline
    const int HO = SMALL ? 0 : FinP;         // offset of h inside an A row
    const int LDA = KA + 4;                  // +4: the look-ahead fragment read past the last quad
    float* arow = smem;                                        // [BM][LDA]
    float* wbuf = arow + (size_t)BM * LDA;                     // [kStages][kKC][H] (+ one pad row)
    float* wih = wbuf + (size_t)(kStages * kKC + 1) * H;       // SMALL: [kSmallFin][H]
    float* xs = wih + (SMALL ? kSmallFin * H : 0);             // SMALL: [BM][kSmallFin]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0 = blockIdx.x * BM + warp * TM;   // first node of this warp
    float* myrow = arow + (size_t)warp * TM * LDA;
    float* myx = xs + warp * TM * kSmallFin;

    // ---- initial state, zero padding ---------------------------------------------------
#pragma unroll
    for (int m = 0; m < TM; ++m) {
        const int n = n0 + m;
        for (int f = lane; f < HO; f += 32) myrow[m * LDA + f] = 0.f;
        if (lane < 4) myrow[m * LDA + KA + lane] = 0.f;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < N) v = *reinterpret_cast<const float4*>(h_state + (size_t)n * H + q * 128 + lane * 4);
            *reinterpret_cast<float4*>(myrow + m * LDA + HO + q * 128 + lane * 4) = v;
        }
    }
    if (SMALL) {
        for (int i = tid; i < kSmallFin * H; i += kScanThreads) wih[i] = (i < FinP * H) ? wpack[i] : 0.f;
    }
    for (int i = tid; i < H; i += kScanThreads) wbuf[(size_t)kStages * kKC * H + i] = 0.f;
    float4 bq[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) bq[q] = ldg_f4(bias + q * 128 + lane * 4);

    const float* wsrc = wpack + (SMALL ? (size_t)FinP * H : 0);
    const int NCH = KA / kKC;
    const long long total = (long long)Tc * NCH;
    auto issue = [&](long long g) {
        if (g < total) {
            const int c = (int)(g % NCH);
            const float* src = wsrc + (size_t)c * kKC * H;
            float* dst = wbuf + (size_t)(g % kStages) * kKC * H;
#pragma unroll
            for (int i = 0; i < kKC * (H / 4) / kScanThreads; ++i)
                cp_async16(dst + (i * kScanThreads + tid) * 4, src + (i * kScanThreads + tid) * 4);
        }
        cp_async_commit();
    };
    issue(0);
    issue(1);

    // x_t for the small-Fin path: lane l holds element (m = l / 8 + 4*j, f = l % 8) of the slot
    auto load_x_small = [&](int t, float (&xr)[(TM + 3) / 4]) {
#pragma unroll
        for (int j = 0; j < (TM + 3) / 4; ++j) {
            const int m = (lane >> 3) + 4 * j, f = lane & 7, n = n0 + m;
            xr[j] = (t < Tc && m < TM && f < Fin && n < N)
                        ? __ldg(x + (size_t)t * x_ts + (size_t)n * x_ns + f) : 0.f;
        }
    };
    auto store_x_small = [&](const float (&xr)[(TM + 3) / 4]) {
#pragma unroll
        for (int j = 0; j < (TM + 3) / 4; ++j) {
            const int m = (lane >> 3) + 4 * j;
            if (m < TM) myx[m * kSmallFin + (lane & 7)] = xr[j];
        }
    };
    float xr[(TM + 3) / 4];
    if (SMALL) {
        load_x_small(0, xr);
        store_x_small(xr);
    }
    __syncthreads();

    long long g = 0;
    for (int t = 0; t < Tc; ++t) {
        float2 acc[TM][NQ][2];
#pragma unroll
        for (int m = 0; m < TM; ++m)
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                acc[m][q][0] = make_float2(bq[q].x, bq[q].y);
                acc[m][q][1] = make_float2(bq[q].z, bq[q].w);
            }
        if (SMALL) {
            // input projection x_t W_ih^T straight into the accumulators
            for (int f = 0; f < FinP; ++f) {
                float4 w[NQ];
#pragma unroll
                for (int q = 0; q < NQ; ++q) w[q] = *reinterpret_cast<const float4*>(wih + f * H + q * 128 + lane * 4);
#pragma unroll
                for (int m = 0; m < TM; ++m) {
                    const float xv = myx[m * kSmallFin + f];
#pragma unroll
                    for (int q = 0; q < NQ; ++q) fma4(acc[m][q][0], acc[m][q][1], xv, w[q]);
                }
            }
            __syncwarp();
            load_x_small(t + 1, xr);      // in flight during the whole contraction below
        } else {
            const float* xt = x + (size_t)t * x_ts;
#pragma unroll
            for (int m = 0; m < TM; ++m) {
                const int n = n0 + m;
                for (int f = lane; f < Fin; f += 32)
                    myrow[m * LDA + f] = (n < N) ? __ldg(xt + (size_t)n * x_ns + f) : 0.f;
            }
            __syncwarp();
        }

        for (int c = 0; c < NCH; ++c, ++g) {
            cp_async_wait<1>();
            __syncthreads();
            issue(g + 2);
            const float* wst = wbuf + (size_t)(g % kStages) * kKC * H + lane * 4;
            const float* ap = myrow + c * kKC;
            float4 a_cur[TM], w_cur[NQ];
#pragma unroll
            for (int m = 0; m < TM; ++m) a_cur[m] = *reinterpret_cast<const float4*>(ap + m * LDA);
#pragma unroll
            for (int q = 0; q < NQ; ++q) w_cur[q] = *reinterpret_cast<const float4*>(wst + q * 128);
#pragma unroll 2
            for (int kq = 0; kq < kKC; kq += 4) {
                float4 a_nxt[TM];
#pragma unroll
                for (int m = 0; m < TM; ++m) a_nxt[m] = *reinterpret_cast<const float4*>(ap + m * LDA + kq + 4);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    float4 w_nxt[NQ];
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
                        w_nxt[q] = *reinterpret_cast<const float4*>(wst + (kq + kk + 1) * H + q * 128);
#pragma unroll
                    for (int m = 0; m < TM; ++m) {
                        const float av = f4_get(a_cur[m], kk);
#pragma unroll
                        for (int q = 0; q < NQ; ++q) fma4(acc[m][q][0], acc[m][q][1], av, w_cur[q]);
                    }
#pragma unroll
                    for (int q = 0; q < NQ; ++q) w_cur[q] = w_nxt[q];
                }
#pragma unroll
                for (int m = 0; m < TM; ++m) a_cur[m] = a_nxt[m];
            }
        }
        __syncwarp();   // every lane is done reading this warp's A rows

        // ---- epilogue: activation, leaky blend, write state (smem) and output (HBM) ----
        float* ot = out + (size_t)t * o_ts;
#pragma unroll
        for (int m = 0; m < TM; ++m) {
            const int n = n0 + m;
            float scale = 1.f;
            if (act == SGP_ACT_SELF_NORM) {
                float ss = 0.f;
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    ss += acc[m][q][0].x * acc[m][q][0].x + acc[m][q][0].y * acc[m][q][0].y +
                          acc[m][q][1].x * acc[m][q][1].x + acc[m][q][1].y * acc[m][q][1].y;
                }
                ss = warp_sum(ss);
                scale = 1.f / fmaxf(sqrtf(ss), 1e-12f);
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                float* hp = myrow + m * LDA + HO + q * 128 + lane * 4;
                const float4 ho = *reinterpret_cast<const float4*>(hp);
                float4 z = make_float4(acc[m][q][0].x, acc[m][q][0].y, acc[m][q][1].x, acc[m][q][1].y);
                if (act == SGP_ACT_SELF_NORM) {
                    z.x *= scale; z.y *= scale; z.z *= scale; z.w *= scale;
                } else {
                    z.x = activate(z.x, act); z.y = activate(z.y, act);
                    z.z = activate(z.z, act); z.w = activate(z.w, act);
                }
                float4 hn;
                hn.x = oma * ho.x + alpha * z.x;
                hn.y = oma * ho.y + alpha * z.y;
                hn.z = oma * ho.z + alpha * z.z;
                hn.w = oma * ho.w + alpha * z.w;
                *reinterpret_cast<float4*>(hp) = hn;
                if (n < N) st_f4(ot + (size_t)n * o_ns + q * 128 + lane * 4, hn);
            }
        }
        if (SMALL) store_x_small(xr);
        __syncwarp();
    }
    cp_async_wait<0>();

    // ---- carry the state to the next chunk ------------------------------------------------
#pragma unroll
    for (int m = 0; m < TM; ++m) {
        const int n = n0 + m;
        if (n < N) {
#pragma unroll
            for (int q = 0; q < NQ; ++q)
                *reinterpret_cast<float4*>(h_state + (size_t)n * H + q * 128 + lane * 4) =
                    *reinterpret_cast<const float4*>(myrow + m * LDA + HO + q * 128 + lane * 4);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Small reservoirs, ALL layers in one launch (H in {16, 32, 64}: the reference's shipped configs
// sgp_la.yaml H = 64 x L = 2, sgp_pv.yaml H = 16 x L = 8, the default H = 32).  Per time step the
// layers run back to back inside the kernel — layer l reads layer l-1's NEW state of the same step
// (reservoir.py:170-176) from shared memory, never from HBM — and every layer's weights
// ([W_ih^T ; W_hh^T], k-major) stay resident in shared memory for the whole chunk.
//   * warp = TN nodes per lane group; a lane group (LPN = min(32, H) lanes) covers one node row,
//     lane jl owns columns jl, jl + LPN (CPL = H / LPN columns); H = 16 packs two groups per warp;
//   * per 4 k-rows: 4 CPL weight LDS (conflict-free), then per node one LDS.128 of its input /
//     state row (broadcast) feeding 4 CPL FFMA;
//   * the node rows [x_t | h_0 | .. | h_{L-1}] are private to the warp: only __syncwarp per layer.
// Bound: fp32 FMA issue / step latency; HBM bytes per node-step 4 (Fin + L H).
// ---------------------------------------------------------------------------------------------
constexpr int kSmallMaxLayers = 8;
struct SmallLayers {
    const float* w_ih[kSmallMaxLayers];    // [H, Fin_l]
    const float* w_hh[kSmallMaxLayers];    // [H, H]
    const float* bias[kSmallMaxLayers];    // [H]
    float alpha[kSmallMaxLayers];
};

template <int H, int TN>
__global__ void __launch_bounds__(256)
reservoir_scan_small(const float* __restrict__ x, int64_t x_ts, int64_t x_ns, int Fin, const SmallLayers lay,
                     int L, int act, float* __restrict__ h_state, float* __restrict__ out, int64_t o_ts,
                     int64_t o_ns, int Tc, int N) {
    constexpr int LPN = H < 32 ? H : 32;           // lanes per node row
    constexpr int CPL = H / LPN;                    // columns per lane
    constexpr int GPW = 32 / LPN;                   // node groups per warp
    constexpr int NPW = TN * GPW;                   // nodes per warp
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int FinP = (Fin + 3) & ~3;
    const int ROW = FinP + L * H;                   // [x | h_0 | ... | h_{L-1}]
    // weights: layer l at woff(l): (Kin_l + H) rows of H floats, Kin_0 = FinP, Kin_l = H
    float* wsm = smem;
    const int w_total = (FinP + H) * H + (L - 1) * 2 * H * H;
    float* bsm = wsm + w_total;                     // [L][H]
    float* rows = bsm + L * H + (size_t)warp * NPW * ROW;
    for (int l = 0; l < L; ++l) {
        const int Kin = l == 0 ? Fin : H, KinP = l == 0 ? FinP : H;
        float* w = wsm + (l == 0 ? 0 : (FinP + H) * H + (l - 1) * 2 * H * H);
        for (int i = threadIdx.x; i < (KinP + H) * H; i += blockDim.x) {
            const int k = i / H, j = i % H;
            float v = 0.f;
            if (k < KinP) { if (k < Kin) v = __ldg(lay.w_ih[l] + (size_t)j * Kin + k); }
            else v = __ldg(lay.w_hh[l] + (size_t)j * H + (k - KinP));
            w[i] = v;
        }
        for (int j = threadIdx.x; j < H; j += blockDim.x) bsm[l * H + j] = __ldg(lay.bias[l] + j);
    }
    const int grp = lane / LPN, jl = lane % LPN;
    const int n0 = (blockIdx.x * (blockDim.x >> 5) + warp) * NPW;
    // initial state, zero the x padding
    for (int i = lane; i < NPW * ROW; i += 32) {
        const int m = i / ROW, c = i % ROW;
        const int n = n0 + m;
        float v = 0.f;
        if (c >= FinP && n < N) { const int l = (c - FinP) / H; v = h_state[((size_t)l * N + n) * H + (c - FinP) % H]; }
        rows[i] = v;
    }
    __syncthreads();
    for (int t = 0; t < Tc; ++t) {
        for (int i = lane; i < NPW * Fin; i += 32) {
            const int m = i / Fin, f = i % Fin, n = n0 + m;
            rows[m * ROW + f] = n < N ? __ldg(x + (size_t)t * x_ts + (size_t)n * x_ns + f) : 0.f;
        }
        __syncwarp();
        for (int l = 0; l < L; ++l) {
            const int KinP = l == 0 ? FinP : H;
            const float* w = wsm + (l == 0 ? 0 : (FinP + H) * H + (l - 1) * 2 * H * H);
            const int in_off = l == 0 ? 0 : FinP + (l - 1) * H, st_off = FinP + l * H;
            float acc[TN][CPL];
#pragma unroll
            for (int m = 0; m < TN; ++m)
#pragma unroll
                for (int c = 0; c < CPL; ++c) acc[m][c] = bsm[l * H + jl + c * LPN];
            // two row segments: the layer's input (x or the previous layer's new state), then its own state
#pragma unroll 1
            for (int seg = 0; seg < 2; ++seg) {
                const int K = seg == 0 ? KinP : H, a_off = seg == 0 ? in_off : st_off;
                const float* ws = w + (seg == 0 ? 0 : KinP * H);
#pragma unroll 2
                for (int k = 0; k < K; k += 4) {
                    float wv[4][CPL];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                        for (int c = 0; c < CPL; ++c) wv[kk][c] = ws[(k + kk) * H + jl + c * LPN];
#pragma unroll
                    for (int m = 0; m < TN; ++m) {
                        const float4 a = *reinterpret_cast<const float4*>(rows + (grp * TN + m) * ROW + a_off + k);
#pragma unroll
                        for (int c = 0; c < CPL; ++c) {
                            acc[m][c] = fmaf(a.x, wv[0][c], acc[m][c]);
                            acc[m][c] = fmaf(a.y, wv[1][c], acc[m][c]);
                            acc[m][c] = fmaf(a.z, wv[2][c], acc[m][c]);
                            acc[m][c] = fmaf(a.w, wv[3][c], acc[m][c]);
                        }
                    }
                }
            }
            const float alpha = lay.alpha[l], oma = 1.f - alpha;
            float scale[TN];
#pragma unroll
            for (int m = 0; m < TN; ++m) {
                scale[m] = 1.f;
                if (act == SGP_ACT_SELF_NORM) {
                    float ss = 0.f;
#pragma unroll
                    for (int c = 0; c < CPL; ++c) ss = fmaf(acc[m][c], acc[m][c], ss);
#pragma unroll
                    for (int o = LPN / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                    scale[m] = 1.f / fmaxf(sqrtf(ss), 1e-12f);
                }
            }
            __syncwarp();                            // every lane is done reading the old state rows
#pragma unroll
            for (int m = 0; m < TN; ++m) {
                const int n = n0 + grp * TN + m;
                float* r = rows + (grp * TN + m) * ROW + st_off;
#pragma unroll
                for (int c = 0; c < CPL; ++c) {
                    const int j = jl + c * LPN;
                    const float z = act == SGP_ACT_SELF_NORM ? acc[m][c] * scale[m] : activate(acc[m][c], act);
                    const float v = fmaf(alpha, z, oma * r[j]);
                    r[j] = v;
                    if (n < N) out[(size_t)t * o_ts + (size_t)n * o_ns + l * H + j] = v;
                }
            }
            __syncwarp();
        }
    }
    for (int i = lane; i < NPW * L * H; i += 32) {
        const int m = i / (L * H), c = i % (L * H), n = n0 + m;
        if (n < N) h_state[((size_t)(c / H) * N + n) * H + c % H] = rows[m * ROW + FinP + c];
    }
}

// Generic kernel: any H, any Fin.  One warp per node, lanes stride over the output columns,
// W^T read through L1/L2 (coalesced over the column index), A row [x_t | h] in shared memory.
constexpr int kGenWarps = 8;
__global__ void __launch_bounds__(kGenWarps * 32)
reservoir_scan_generic(const float* __restrict__ x, int64_t x_ts, int64_t x_ns, int Fin, int FinP,
                       const float* __restrict__ wpack, const float* __restrict__ bias,
                       float alpha, float oma, int act,
                       float* __restrict__ h_state,
                       float* __restrict__ out, int64_t o_ts, int64_t o_ns,
                       int Tc, int N, int H) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = blockIdx.x * kGenWarps + warp;
    const int Ktot = FinP + H;
    float* a = smem + (size_t)warp * (Ktot + H);   // [x | h]
    float* hn = a + Ktot;                          // new state
    if (n >= N) return;                            // warps are independent: no CTA barrier below
    for (int f = lane; f < FinP; f += 32) a[f] = 0.f;
    for (int j = lane; j < H; j += 32) a[FinP + j] = h_state[(size_t)n * H + j];
    __syncwarp();
    for (int t = 0; t < Tc; ++t) {
        for (int f = lane; f < Fin; f += 32) a[f] = __ldg(x + (size_t)t * x_ts + (size_t)n * x_ns + f);
        __syncwarp();
        float ss = 0.f;
        for (int j = lane; j < H; j += 32) {
            float s = __ldg(bias + j);
            for (int k = 0; k < Ktot; ++k) s = fmaf(a[k], __ldg(wpack + (size_t)k * H + j), s);
            hn[j] = s;
            ss += s * s;
        }
        float scale = 1.f;
        if (act == SGP_ACT_SELF_NORM) scale = 1.f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);
        __syncwarp();
        for (int j = lane; j < H; j += 32) {
            const float z = (act == SGP_ACT_SELF_NORM) ? hn[j] * scale : activate(hn[j], act);
            const float v = oma * a[FinP + j] + alpha * z;
            a[FinP + j] = v;
            out[(size_t)t * o_ts + (size_t)n * o_ns + j] = v;
        }
        __syncwarp();
    }
    for (int j = lane; j < H; j += 32) h_state[(size_t)n * H + j] = a[FinP + j];
}

__global__ void reservoir_pack_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                      int Fin, int FinP, int H, float* __restrict__ wpack) {
    const int64_t total = (int64_t)(FinP + H) * H;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i / H), n = (int)(i % H);
        float v = 0.f;
        if (k < Fin) v = w_ih[(size_t)n * Fin + k];
        else if (k >= FinP) v = w_hh[(size_t)n * H + (k - FinP)];
        wpack[i] = v;
    }
}

template <int H, int TM, bool SMALL>
static int launch_tiled(const float* x, int64_t x_ts, int64_t x_ns, int Fin, int FinP,
                        const float* wpack, const float* bias, float alpha, float oma, int act,
                        float* h_state, float* out, int64_t o_ts, int64_t o_ns, int Tc, int N,
                        cudaStream_t st) {
    constexpr int BM = TM * kScanWarps;
    const int KA = SMALL ? H : FinP + H;
    const size_t smem = ((size_t)BM * (KA + 4) + (size_t)(kStages * kKC + 1) * H +
                         (SMALL ? (size_t)kSmallFin * H + (size_t)BM * kSmallFin : 0)) * sizeof(float);
    if (smem > 227 * 1024) return 1;   // caller falls back to a narrower node tile
    auto kern = reservoir_scan_tiled<H, TM, SMALL>;
    SGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (N + BM - 1) / BM;
    kern<<<grid, kScanThreads, smem, st>>>(x, x_ts, x_ns, Fin, FinP, wpack, bias, alpha, oma, act,
                                           h_state, out, o_ts, o_ns, Tc, N);
    SGP_LAUNCH_CHECK("reservoir_scan_tiled");
    return SGP_OK;
}

template <int H, bool SMALL>
static int launch_tiled_h(const float* x, int64_t x_ts, int64_t x_ns, int Fin, int FinP,
                          const float* wpack, const float* bias, float alpha, float oma, int act,
                          float* h_state, float* out, int64_t o_ts, int64_t o_ns, int Tc, int N,
                          cudaStream_t st) {
    // enough CTAs to fill the machine first, then the widest node tile that fits shared memory
    int rc = 1;
#define SGP_TRY(TM_)                                                                              \
    if (rc == 1) rc = launch_tiled<H, TM_, SMALL>(x, x_ts, x_ns, Fin, FinP, wpack, bias, alpha, oma, \
                                                  act, h_state, out, o_ts, o_ns, Tc, N, st)
    if (N >= 64 * kNumSMs) SGP_TRY(8);
    if (N >= 32 * kNumSMs) SGP_TRY(4);
    SGP_TRY(2);
#undef SGP_TRY
    return rc;
}

}  // namespace sgp

using namespace sgp;

extern "C" int sgp_reservoir_pack_rows(int Fin, int H) {
    return (Fin >= 1 && H >= 1) ? fin_padded(Fin) + H : 0;
}

extern "C" int sgp_reservoir_pack(const float* w_ih, const float* w_hh, int Fin, int H, float* wpack,
                                  void* stream) {
    SGP_REQUIRE(w_ih && w_hh && wpack, SGP_EINVAL, "sgp_reservoir_pack: null pointer");
    SGP_REQUIRE(Fin >= 1 && H >= 1, SGP_EINVAL, "sgp_reservoir_pack: Fin=%d H=%d", Fin, H);
    const int FinP = fin_padded(Fin);
    const int64_t total = (int64_t)(FinP + H) * H;
    const int grid = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
    reservoir_pack_kernel<<<grid, 256, 0, as_stream(stream)>>>(w_ih, w_hh, Fin, FinP, H, wpack);
    SGP_LAUNCH_CHECK("reservoir_pack");
    return SGP_OK;
}

extern "C" int sgp_reservoir_scan(const float* x, int64_t x_t_stride, int64_t x_n_stride, int Fin,
                                  const float* wpack, const float* bias, float alpha,
                                  float one_minus_alpha, int act, float* h_state, float* out,
                                  int64_t out_t_stride, int64_t out_n_stride, int Tc, int N, int H,
                                  void* stream) {
    SGP_REQUIRE(x && wpack && bias && h_state && out, SGP_EINVAL, "sgp_reservoir_scan: null pointer");
    SGP_REQUIRE(Fin >= 1 && H >= 1 && N >= 0 && Tc >= 0, SGP_EINVAL,
                "sgp_reservoir_scan: Fin=%d H=%d N=%d Tc=%d", Fin, H, N, Tc);
    SGP_REQUIRE(act >= SGP_ACT_TANH && act <= SGP_ACT_IDENTITY, SGP_EINVAL,
                "sgp_reservoir_scan: unknown activation code %d", act);
    if (N == 0 || Tc == 0) return SGP_OK;
    cudaStream_t st = as_stream(stream);
    const int FinP = fin_padded(Fin);
    const bool vec_ok = aligned16(out) && aligned16(h_state) && aligned16(wpack) && aligned16(bias) &&
                        out_t_stride % 4 == 0 && out_n_stride % 4 == 0;
    if ((H == 128 || H == 256) && vec_ok) {
        int rc;
#define SGP_ARGS x, x_t_stride, x_n_stride, Fin, FinP, wpack, bias, alpha, one_minus_alpha, act, h_state, \
                 out, out_t_stride, out_n_stride, Tc, N, st
        if (H == 256) rc = Fin <= kSmallFin ? launch_tiled_h<256, true>(SGP_ARGS) : launch_tiled_h<256, false>(SGP_ARGS);
        else rc = Fin <= kSmallFin ? launch_tiled_h<128, true>(SGP_ARGS) : launch_tiled_h<128, false>(SGP_ARGS);
#undef SGP_ARGS
        if (rc != 1) return rc;
    }
    const size_t smem = (size_t)kGenWarps * (FinP + 2 * (size_t)H) * sizeof(float);
    SGP_REQUIRE(smem <= 227 * 1024, SGP_EUNSUPPORTED,
                "sgp_reservoir_scan: Fin=%d H=%d needs %zu B of shared memory", Fin, H, smem);
    SGP_CUDA(cudaFuncSetAttribute(reservoir_scan_generic, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    reservoir_scan_generic<<<(N + kGenWarps - 1) / kGenWarps, kGenWarps * 32, smem, st>>>(
        x, x_t_stride, x_n_stride, Fin, FinP, wpack, bias, alpha, one_minus_alpha, act, h_state, out,
        out_t_stride, out_n_stride, Tc, N, H);
    SGP_LAUNCH_CHECK("reservoir_scan_generic");
    return SGP_OK;
}

template <int H, int TN>
static int launch_small(const float* x, int64_t x_ts, int64_t x_ns, int Fin, const SmallLayers& lay, int L, int act,
                        float* h_state, float* out, int64_t o_ts, int64_t o_ns, int Tc, int N, cudaStream_t st) {
    constexpr int LPN = H < 32 ? H : 32, NPW = TN * (32 / LPN), WARPS = 8;
    const int FinP = (Fin + 3) & ~3, ROW = FinP + L * H;
    const size_t smem = ((size_t)(FinP + H) * H + (size_t)(L - 1) * 2 * H * H + (size_t)L * H +
                         (size_t)WARPS * NPW * ROW) * sizeof(float);
    if (smem > 200 * 1024) return 1;
    auto kern = reservoir_scan_small<H, TN>;
    SGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_cta = WARPS * NPW;
    kern<<<(N + per_cta - 1) / per_cta, WARPS * 32, smem, st>>>(x, x_ts, x_ns, Fin, lay, L, act, h_state, out,
                                                                 o_ts, o_ns, Tc, N);
    SGP_LAUNCH_CHECK("reservoir_scan_small");
    return SGP_OK;
}

template <int H>
static int launch_small_h(const float* x, int64_t x_ts, int64_t x_ns, int Fin, const SmallLayers& lay, int L,
                          int act, float* h_state, float* out, int64_t o_ts, int64_t o_ns, int Tc, int N,
                          cudaStream_t st) {
    // wide node tiles only when there are enough nodes to keep every SM busy with them
    constexpr int G = 32 / (H < 32 ? H : 32);
    int rc = 1;
    if (N >= 8 * 8 * G * kNumSMs) rc = launch_small<H, 8>(x, x_ts, x_ns, Fin, lay, L, act, h_state, out, o_ts, o_ns, Tc, N, st);
    if (rc == 1 && N >= 8 * 4 * G * kNumSMs) rc = launch_small<H, 4>(x, x_ts, x_ns, Fin, lay, L, act, h_state, out, o_ts, o_ns, Tc, N, st);
    if (rc == 1) rc = launch_small<H, 2>(x, x_ts, x_ns, Fin, lay, L, act, h_state, out, o_ts, o_ns, Tc, N, st);
    return rc;
}

extern "C" int sgp_reservoir_scan_multi(const float* x, int64_t x_t_stride, int64_t x_n_stride, int Fin,
                                        const float* const* w_ih, const float* const* w_hh,
                                        const float* const* bias, const float* alpha, int act, float* h_state,
                                        float* out, int64_t out_t_stride, int64_t out_n_stride, int Tc, int N,
                                        int H, int L, void* stream) {
    SGP_REQUIRE(x && w_ih && w_hh && bias && alpha && h_state && out, SGP_EINVAL, "sgp_reservoir_scan_multi: null pointer");
    SGP_REQUIRE(Fin >= 1 && N >= 0 && Tc >= 0 && L >= 1, SGP_EINVAL, "sgp_reservoir_scan_multi: Fin=%d N=%d Tc=%d L=%d", Fin, N, Tc, L);
    SGP_REQUIRE(act >= SGP_ACT_TANH && act <= SGP_ACT_IDENTITY, SGP_EINVAL, "sgp_reservoir_scan_multi: activation %d", act);
    SGP_REQUIRE((H == 16 || H == 32 || H == 64) && L <= kSmallMaxLayers && Fin <= 64, SGP_EUNSUPPORTED,
                "sgp_reservoir_scan_multi: H=%d L=%d Fin=%d (H in {16,32,64}, L <= %d, Fin <= 64)", H, L, Fin, kSmallMaxLayers);
    if (N == 0 || Tc == 0) return SGP_OK;
    SmallLayers lay{};
    for (int l = 0; l < L; ++l) {
        SGP_REQUIRE(w_ih[l] && w_hh[l] && bias[l], SGP_EINVAL, "sgp_reservoir_scan_multi: null weights of layer %d", l);
        lay.w_ih[l] = w_ih[l]; lay.w_hh[l] = w_hh[l]; lay.bias[l] = bias[l]; lay.alpha[l] = alpha[l];
    }
    cudaStream_t st = as_stream(stream);
    int rc;
#define SGP_SMALL_ARGS x, x_t_stride, x_n_stride, Fin, lay, L, act, h_state, out, out_t_stride, out_n_stride, Tc, N, st
    if (H == 64) rc = launch_small_h<64>(SGP_SMALL_ARGS);
    else if (H == 32) rc = launch_small_h<32>(SGP_SMALL_ARGS);
    else rc = launch_small_h<16>(SGP_SMALL_ARGS);
#undef SGP_SMALL_ARGS
    SGP_REQUIRE(rc != 1, SGP_EUNSUPPORTED, "sgp_reservoir_scan_multi: H=%d x L=%d does not fit shared memory", H, L);
    return rc;
}
