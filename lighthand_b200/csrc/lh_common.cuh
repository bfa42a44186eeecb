// Shared device/host helpers for liblighthand_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/lighthand_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "liblighthand_b200 is written for sm_100a (B200) only"
#endif

#define LH_CHECK_LAUNCH()                                  \
    do {                                                   \
        cudaError_t e__ = cudaGetLastError();              \
        if (e__ != cudaSuccess) return (int)e__;           \
    } while (0)

namespace lh {

constexpr int kWarp = 32;

__device__ __forceinline__ float4 ldg_stream4(const float* p) {
    // read-once data: bypass L1 allocation, keep L2 for whoever reads next
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream4(float* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_cs4(float* p, const float4& v) {
    // write-once output nobody re-reads soon: evict-first streaming store
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Total order used by the decode: NaN is the maximum (numpy argmax/amax), -0 == +0.
__device__ __forceinline__ uint32_t order_key(float v) {
    if (v != v) return 0xffffffffu;
    v = v + 0.0f;  // -0 -> +0
    uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace lh
