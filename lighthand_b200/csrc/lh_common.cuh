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


// ---- mbarrier / bulk-copy (TMA) primitives shared by the K1 ring (warp_tma.cu) and the heat-map ring (heatmap.cu) ------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
// try_wait parks the warp in hardware until the phase completes or the hint (ns) runs out: few spin instructions
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "LH_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra LH_DONE;\n\t"
        "bra LH_WAIT;\n\t"
        "LH_DONE:\n\t}"
        :: "r"(bar), "r"(parity), "r"(20000u) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared (TMA engine), completion counted in bytes on `bar`; 16-byte aligned, size % 16 == 0
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- cv2.warpAffine coordinate arithmetic, shared by the warp kernels -----------------
__device__ __forceinline__ int clamp_i(int v, int lo, int hi) { return min(max(v, lo), hi); }

__device__ __forceinline__ int sat_i32(double v) {
    // cv::saturate_cast<int>(double) == cvRound (round-half-even) with saturation.  F2I.F64 saturates by itself
    // (branch-free); only NaN needs the select (F2I gives 0, the reference path INT_MIN)
    const int r = __double2int_rn(v);
    return (v != v) ? (int)0x80000000 : r;
}

// OpenCV's in-place inversion of the forward 2x3 matrix (imgwarp.cpp; Appendix A step 1).
__device__ __forceinline__ void invert_affine(const double* M, double* o) {
    double m0 = M[0], m1 = M[1], m2 = M[2], m3 = M[3], m4 = M[4], m5 = M[5];
    double D = __dsub_rn(__dmul_rn(m0, m4), __dmul_rn(m1, m3));
    D = (D != 0.0) ? __drcp_rn(D) : 0.0;      // correctly rounded 1 / D, like the division OpenCV performs
    const double a11 = __dmul_rn(m4, D), a22 = __dmul_rn(m0, D);
    m0 = a11; m1 = __dmul_rn(m1, -D); m3 = __dmul_rn(m3, -D); m4 = a22;
    const double b1 = __dsub_rn(__dmul_rn(-m0, m2), __dmul_rn(m1, m5));
    const double b2 = __dsub_rn(__dmul_rn(-m3, m2), __dmul_rn(m4, m5));
    o[0] = m0; o[1] = m1; o[2] = b1; o[3] = m3; o[4] = m4; o[5] = b2;
}

// cv2's fixed-point coordinate terms for output column x / row y (Appendix A step 2)
__device__ __forceinline__ int col_term(double m, int x) {
    return clamp_i(sat_i32(__dmul_rn(__dmul_rn(m, (double)x), 1024.0)), -(1 << 30), 1 << 30);
}
__device__ __forceinline__ int row_term(double m, double b, int y) {
    return clamp_i(sat_i32(__dmul_rn(__dadd_rn(__dmul_rn(m, (double)y), b), 1024.0)), -(1 << 30), 1 << 30) + 16;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace lh
