// Shared helpers for the sgp_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/sgp_b200.h"

namespace sgp {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
extern std::atomic<int> g_tc_cta_limit;   // persistent CTAs per tensor-core hop launch (sgp_tc_set_cta_limit)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Call after every kernel launch: counts it and converts a launch error into SGP_ECUDA.
#define SGP_LAUNCH_CHECK(name)                                                       \
    do {                                                                             \
        ::sgp::g_launches.fetch_add(1, std::memory_order_relaxed);                   \
        cudaError_t e__ = cudaGetLastError();                                        \
        if (e__ != cudaSuccess) {                                                    \
            ::sgp::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
            return SGP_ECUDA;                                                        \
        }                                                                            \
    } while (0)

#define SGP_CUDA(call)                                                               \
    do {                                                                             \
        cudaError_t e__ = (call);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            ::sgp::set_error("%s failed: %s", #call, cudaGetErrorString(e__));       \
            return SGP_ECUDA;                                                        \
        }                                                                            \
    } while (0)

#define SGP_REQUIRE(cond, code, ...)                                                 \
    do {                                                                             \
        if (!(cond)) {                                                               \
            ::sgp::set_error(__VA_ARGS__);                                           \
            return code;                                                             \
        }                                                                            \
    } while (0)

// ---- device helpers ---------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_f4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
// gathered rows are used once per warp: read-only path, do not allocate in L1
__device__ __forceinline__ float4 ldg_f4_stream(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// streaming store (written once, consumed by a later kernel / the host): do not keep in L1
__device__ __forceinline__ void st_f4(float* p, float4 v) {
    *reinterpret_cast<float4*>(p) = v;
}
// packed fp32x2 FMA: acc.xy += a * b.xy  (SASS: FFMA2 with the scalar operand broadcast)
__device__ __forceinline__ void fma2(float2& acc, float a, float2 b) {
    acc = __ffma2_rn(make_float2(a, a), b, acc);
}
__device__ __forceinline__ void fma4(float2& lo, float2& hi, float a, const float4& b) {
    lo = __ffma2_rn(make_float2(a, a), make_float2(b.x, b.y), lo);
    hi = __ffma2_rn(make_float2(a, a), make_float2(b.z, b.w), hi);
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

}  // namespace sgp
