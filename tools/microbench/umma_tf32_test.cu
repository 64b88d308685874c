// Bring-up test for the tcgen05 pieces the RBU-64 SpMM needs (sm_100a):
//   D[f, r] (+)= sum_k X[k, f] * B[r, k]        M = 128 features, N = 64 rows, K = 32 per chunk
//   * A operand = gathered source rows X[k][0:128], MN-major, SWIZZLE_128B, written with plain
//     16-byte stores at the swizzled offsets (the production kernel uses cp.async there)
//   * B operand = operator slab image, K-major, SWIZZLE_128B, pre-swizzled on the host
//   * kind::tf32, fp32 accumulate in TMEM, tcgen05.commit -> mbarrier, tcgen05.ld epilogue
//   * optional 3xTF32 split: hi = x & 0xffffe000, lo = x - hi
// All waits are bounded (a stuck barrier sets an error flag instead of hanging the GPU).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_tf32_test umma_tf32_test.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, KC = 32;   // per chunk
constexpr int A_STAGE = KC * M * 4;       // 16 KB
constexpr int B_STAGE = N * KC * 4;       // 8 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout_type = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;          // version = 1 (Blackwell)
    d |= (uint64_t)layout_type << 61;   // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B (MN-major tf32)
    return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int* err) {
    uint32_t a = smem_u32(bar), done = 0;
    for (long long it = 0; it < (1ll << 22); ++it) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return true;
    }
    atomicExch(err, 1);
    return false;
}

// X: [nchunks*KC][M] fp32 row-major (gathered rows), Bimg: [nchunks][split?2:1][B_STAGE bytes]
__global__ void __launch_bounds__(128, 1)
umma_test(const float* __restrict__ X, const float* __restrict__ Bimg, float* __restrict__ D,
          int nchunks, int split, int* err, int amajor, int prefill, int ts) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* a_hi = smem;                     // A_STAGE
    uint8_t* a_lo = a_hi + A_STAGE;           // A_STAGE
    uint8_t* b_hi = a_lo + A_STAGE;           // B_STAGE
    uint8_t* b_lo = b_hi + B_STAGE;           // B_STAGE
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;

    // instruction descriptor: c=f32, a=b=tf32, A MN-major, B K-major, N, M
    if (prefill) {   // sentinel in the accumulator: shows whether the MMA writes TMEM at all
        const uint32_t ta = tmem_d + ((uint32_t)(warp * 32) << 16);
        for (int j = 0; j < N; ++j)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(ta + j), "r"(__float_as_uint(1000.f + j)));
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)amajor << 15) | (0u << 16) |
                           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    uint32_t parity = 0;
    for (int c = 0; c < nchunks; ++c) {
        // ---- stage A: row k (32 rows) x 128 features, MN-major SW128_BASE32B atoms [k/4][b][k%4][128B],
        //      the 32-byte chunk index inside a 128-byte row is XORed with k%4
        for (int i = tid; i < KC * 32; i += 128) {           // i = k*32 + (16B chunk index 0..31)
            const int k = i >> 5, q = i & 31, b = q >> 3, ch = q & 7, kg = k >> 2, r = k & 3;
            float4 v = *reinterpret_cast<const float4*>(X + (size_t)(c * KC + k) * M + q * 4);
            const int off = (kg * 4 + b) * 512 + r * 128 + ((((ch >> 1) ^ r) << 5) | ((ch & 1) << 4));
            if (!amajor) {   // K-major A image: element (m = feature, k): atoms of 8 m-rows x 128 B
                const float vv[4] = {v.x, v.y, v.z, v.w};
                for (int e = 0; e < 4; ++e) {
                    const int m = q * 4 + e, mg = m >> 3, mr = m & 7;
                    const int o2 = mg * 1024 + mr * 128 + (((k >> 2) ^ mr) << 4) + (k & 3) * 4;
                    *reinterpret_cast<float*>(a_hi + o2) = vv[e];
                }
                continue;
            }
            float4 hi = v, lo = make_float4(0, 0, 0, 0);
            if (split) {
                hi.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); lo.x = v.x - hi.x;
                hi.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); lo.y = v.y - hi.y;
                hi.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); lo.z = v.z - hi.z;
                hi.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); lo.w = v.w - hi.w;
            }
            *reinterpret_cast<float4*>(a_hi + off) = (split == 2) ? v : hi;
            *reinterpret_cast<float4*>(a_lo + off) = lo;
        }
        // ---- stage B images (already swizzled on the host)
        const float* bsrc = Bimg + (size_t)c * (split ? 2 : 1) * (B_STAGE / 4);
        for (int i = tid; i < B_STAGE / 16; i += 128) {
            reinterpret_cast<float4*>(b_hi)[i] = reinterpret_cast<const float4*>(bsrc)[i];
            if (split) reinterpret_cast<float4*>(b_lo)[i] = reinterpret_cast<const float4*>(bsrc + B_STAGE / 4)[i];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (ts) {
            // TS mode: thread = TMEM lane = feature m; read X[k][m] for the chunk's 32 k from the
            // swizzled stage (hi tile holds raw/hi values), write hi -> cols [64,96), lo -> [96,128)
            const int m = tid;
            const uint32_t ta = tmem_d + ((uint32_t)(warp * 32) << 16);
            uint32_t hv[32], lv[32];
            for (int k = 0; k < KC; ++k) {
                const int q = m >> 2, b = q >> 3, ch = q & 7, kg = k >> 2, r = k & 3;
                const int off = (kg * 4 + b) * 512 + r * 128 + ((((ch >> 1) ^ r) << 5) | ((ch & 1) << 4)) + (m & 3) * 4;
                const float x = X[(size_t)(c * KC + k) * M + m];   // same value as the staged one
                (void)off;
                const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
                hv[k] = __float_as_uint(split ? h : x);
                lv[k] = __float_as_uint(split ? x - h : 0.f);
            }
#define ST32(base, arr) asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" :: "r"(base), "r"(arr[0]),"r"(arr[1]),"r"(arr[2]),"r"(arr[3]),"r"(arr[4]),"r"(arr[5]),"r"(arr[6]),"r"(arr[7]),"r"(arr[8]),"r"(arr[9]),"r"(arr[10]),"r"(arr[11]),"r"(arr[12]),"r"(arr[13]),"r"(arr[14]),"r"(arr[15]),"r"(arr[16]),"r"(arr[17]),"r"(arr[18]),"r"(arr[19]),"r"(arr[20]),"r"(arr[21]),"r"(arr[22]),"r"(arr[23]),"r"(arr[24]),"r"(arr[25]),"r"(arr[26]),"r"(arr[27]),"r"(arr[28]),"r"(arr[29]),"r"(arr[30]),"r"(arr[31]) : "memory")
            ST32(ta + 64, hv);
            ST32(ta + 96, lv);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
        }
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (ts) {
                const uint32_t idesc_ts = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | (0u << 16) |
                                          ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
                for (int ks = 0; ks < KC / 8; ++ks) {
                    const uint64_t dbh = make_desc(smem_u32(b_hi) + ks * 32, 16, 1024);
                    const uint64_t dbl = make_desc(smem_u32(b_lo) + ks * 32, 16, 1024);
                    const uint32_t ahi = tmem_d + 64 + ks * 8, alo = tmem_d + 96 + ks * 8;
                    uint32_t acc = (c | ks) ? 1u : 0u;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                                 :: "r"(tmem_d), "r"(ahi), "l"(dbh), "r"(idesc_ts), "r"(acc) : "memory");
                    if (split) {
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                                     :: "r"(tmem_d), "r"(alo), "l"(dbh), "r"(idesc_ts), "r"(1u) : "memory");
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                                     :: "r"(tmem_d), "r"(ahi), "l"(dbl), "r"(idesc_ts), "r"(1u) : "memory");
                    }
                }
            } else
            for (int ks = 0; ks < KC / 8; ++ks) {
                const uint64_t dah = amajor ? make_desc(smem_u32(a_hi) + ks * 4096, 512, 2048, 1)
                                            : make_desc(smem_u32(a_hi) + ks * 32, 16, 1024);
                const uint64_t dal = make_desc(smem_u32(a_lo) + ks * 4096, 512, 2048, 1);
                const uint64_t dbh = make_desc(smem_u32(b_hi) + ks * 32, 16, 1024);
                const uint64_t dbl = make_desc(smem_u32(b_lo) + ks * 32, 16, 1024);
                mma_tf32(tmem_d, dah, dbh, idesc, (c | ks) ? 1u : 0u);
                if (split) {
                    mma_tf32(tmem_d, dal, dbh, idesc, 1u);
                    mma_tf32(tmem_d, dah, dbl, idesc, 1u);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                         :: "r"(smem_u32(&bar)) : "memory");
        }
        // everyone waits for the MMAs of this chunk before the stage is overwritten
        if (!mbar_wait(&bar, parity, err)) break;
        parity ^= 1;
        __syncthreads();
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- epilogue: thread = TMEM lane (feature), 64 columns (rows r)
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
    for (int j = 0; j < N; j += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr + j));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int e = 0; e < 8; ++e) D[(size_t)(j + e) * M + warp * 32 + lane] = __uint_as_float(v[e]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_d), "r"(128));
}

static void swizzle_b(const std::vector<float>& Bv /*[N][K]*/, int K, int c, float* img) {
    for (int r = 0; r < N; ++r)
        for (int k = 0; k < KC; ++k) {
            const int rg = r >> 3, rr = r & 7;
            const int off = rg * 1024 + rr * 128 + (((k >> 2) ^ rr) << 4) + (k & 3) * 4;
            img[off / 4] = Bv[(size_t)r * K + c * KC + k];
        }
}

int run(int nchunks, int split, bool exact_inputs, int amajor = 1, int prefill = 0, int ts = 0) {
    const int K = nchunks * KC;
    std::vector<float> X((size_t)K * M), Bv((size_t)N * K), Dref((size_t)N * M), Dout((size_t)N * M);
    srand(1234 + nchunks + split);
    for (auto& v : X) v = exact_inputs ? (float)((rand() % 17) - 8) : (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto& v : Bv) v = exact_inputs ? (float)((rand() % 9) - 4) * 0.25f : ((rand() % 4) ? 0.f : (float)rand() / RAND_MAX);
    for (int r = 0; r < N; ++r)
        for (int f = 0; f < M; ++f) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)X[(size_t)k * M + f] * Bv[(size_t)r * K + k];
            Dref[(size_t)r * M + f] = (float)s;
        }
    const int per = (split ? 2 : 1) * (B_STAGE / 4);
    std::vector<float> img((size_t)nchunks * per, 0.f);
    for (int c = 0; c < nchunks; ++c) {
        if (!split) swizzle_b(Bv, K, c, img.data() + (size_t)c * per);
        else {
            std::vector<float> hi(Bv.size()), lo(Bv.size());
            for (size_t i = 0; i < Bv.size(); ++i) {
                uint32_t u; memcpy(&u, &Bv[i], 4); u &= 0xffffe000u; memcpy(&hi[i], &u, 4);
                lo[i] = Bv[i] - hi[i];
            }
            swizzle_b(hi, K, c, img.data() + (size_t)c * per);
            swizzle_b(lo, K, c, img.data() + (size_t)c * per + B_STAGE / 4);
        }
    }
    float *dX, *dB, *dD; int* derr;
    cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dB, img.size() * 4); cudaMalloc(&dD, Dout.size() * 4); cudaMalloc(&derr, 4);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, Dout.size() * 4); cudaMemset(derr, 0, 4);
    const size_t smem = 2 * A_STAGE + 2 * B_STAGE + 1024;
    cudaFuncSetAttribute(umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma_test<<<1, 128, smem>>>(dX, dB, dD, nchunks, split, derr, amajor, prefill, ts);
    cudaError_t e = cudaDeviceSynchronize();
    int herr = 0;
    cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(Dout.data(), dD, Dout.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (size_t i = 0; i < Dout.size(); ++i) {
        maxerr = fmax(maxerr, fabs((double)Dout[i] - Dref[i]));
        maxref = fmax(maxref, fabs((double)Dref[i]));
    }
    int nz = 0; for (auto v : Dout) nz += (v != 0.f);
    printf("ts=%d amajor=%d prefill=%d nonzero=%d ", ts, amajor, prefill, nz);
    printf("chunks=%d split=%d exact=%d: cuda=%s barrier_timeout=%d max|err|=%.3e (max|ref|=%.3e) rel=%.2e  D[0..3]=%g %g %g %g ref=%g %g %g %g\n",
           nchunks, split, (int)exact_inputs, cudaGetErrorString(e), herr, maxerr, maxref, maxerr / fmax(maxref, 1e-30),
           Dout[0], Dout[1], Dout[2], Dout[3], Dref[0], Dref[1], Dref[2], Dref[3]);
    cudaFree(dX); cudaFree(dB); cudaFree(dD); cudaFree(derr);
    return (e == cudaSuccess && !herr) ? 0 : 1;
}

int main() {
    int bad = 0;
    bad |= run(1, 0, true, 1, 1, 1);
    bad |= run(3, 0, true, 1, 0, 1);
    bad |= run(11, 1, false, 1, 0, 1);
    bad |= run(11, 2, false, 1, 0, 1);
    bad |= run(1, 0, true, 0, 0);
    bad |= run(1, 0, true, 0, 1);
    bad |= run(1, 0, true, 1, 1);
    bad |= run(1, 0, true);
    bad |= run(3, 0, true);
    bad |= run(4, 0, false);
    bad |= run(4, 1, false);
    bad |= run(11, 1, false);
    bad |= run(11, 2, false);
    bad |= run(4, 2, false);
    return bad;
}
