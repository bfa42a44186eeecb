// torchvision ColorJitter on float CHW images (0..1) + Normalize -- the augmentation branch of
// CustomDataset.__getitem__ (src/tools/dataset.py:134-146: ToTensor -> Resize -> ColorJitter(0.5,0.5,0.5,0.5) ->
// Normalize) and SURVEY.md 8 f-2.
//
// The random draw (order of the four adjustments and their factors) stays on the host, through torch's generator
// in torchvision's draw order (lighthand_b200/tools/dataset.py: color_jitter_params), so a seeded run reproduces the
// reference's parameters; the device applies them.  The four adjustments are restated from torchvision's tensor
// implementation (transforms/_functional_tensor.py: _blend, rgb_to_grayscale, _rgb2hsv, _hsv2rgb) with the same
// float32 operation order (separately rounded multiplies and adds -- torch runs them as separate kernels).
// adjust_contrast blends with the mean grey level of the WHOLE image as it is at that point of the chain, so the
// path is two passes: `jitter_mean_kernel` replays the adjustments that precede the contrast step and reduces the
// grey level per image (fixed-shape partial sums in double: deterministic), `jitter_apply_kernel` runs the chain,
// normalises and writes.  Images without a contrast step skip the first pass (their CTAs exit at once).
#include "lh_common.cuh"

namespace lh {

constexpr int kJitThreads = 256;

__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }
__device__ __forceinline__ float grey(float r, float g, float b) {
    return __fadd_rn(__fadd_rn(__fmul_rn(0.2989f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
}
// _blend(img1, img2, ratio) = (ratio * img1 + (1 - ratio) * img2).clamp(0, 1); omr = float(1.0 - ratio) from the host
__device__ __forceinline__ float blend(float a, float b, float ratio, float omr) {
    return clamp01(__fadd_rn(__fmul_rn(ratio, a), __fmul_rn(omr, b)));
}

__device__ __forceinline__ void adjust_hue(float& r, float& g, float& b, float hue) {
    // _rgb2hsv
    const float maxc = fmaxf(r, fmaxf(g, b)), minc = fminf(r, fminf(g, b));
    const bool eqc = maxc == minc;
    const float cr = __fsub_rn(maxc, minc);
    const float s = __fdiv_rn(cr, eqc ? 1.0f : maxc);
    const float div = eqc ? 1.0f : cr;
    const float rc = __fdiv_rn(__fsub_rn(maxc, r), div), gc = __fdiv_rn(__fsub_rn(maxc, g), div),
                bc = __fdiv_rn(__fsub_rn(maxc, b), div);
    const float hr = (maxc == r) ? __fsub_rn(bc, gc) : 0.0f;
    const float hg = (maxc == g && maxc != r) ? __fsub_rn(__fadd_rn(2.0f, rc), bc) : 0.0f;
    const float hb = (maxc != g && maxc != r) ? __fsub_rn(__fadd_rn(4.0f, gc), rc) : 0.0f;
    float h = __fadd_rn(__fadd_rn(hr, hg), hb);
    h = fmodf(__fadd_rn(__fdiv_rn(h, 6.0f), 1.0f), 1.0f);
    // h = (h + hue_factor) % 1.0   (python-style remainder: result in [0, 1))
    h = __fadd_rn(h, hue);
    h = __fsub_rn(h, floorf(h));
    if (h >= 1.0f) h = 0.0f;
    // _hsv2rgb
    const float v = maxc;
    const float h6 = __fmul_rn(h, 6.0f);
    const float fi = floorf(h6);
    const float f = __fsub_rn(h6, fi);
    int i = (int)fi % 6;
    if (i < 0) i += 6;
    const float p = clamp01(__fmul_rn(v, __fsub_rn(1.0f, s)));
    const float q = clamp01(__fmul_rn(v, __fsub_rn(1.0f, __fmul_rn(s, f))));
    const float t = clamp01(__fmul_rn(v, __fsub_rn(1.0f, __fmul_rn(s, __fsub_rn(1.0f, f)))));
    switch (i) {
        case 0: r = v; g = t; b = p; break;
        case 1: r = q; g = v; b = p; break;
        case 2: r = p; g = v; b = t; break;
        case 3: r = p; g = q; b = v; break;
        case 4: r = t; g = p; b = v; break;
        default: r = v; g = p; b = q; break;
    }
}

// one adjustment of the chain; fn ids are torchvision's: 0 brightness, 1 contrast, 2 saturation, 3 hue
__device__ __forceinline__ void apply_fn(int fn, const lh_jitter_params& jp, float mean, float& r, float& g, float& b) {
    if (fn == 0) {
        const float f = jp.factor[0], o = jp.one_minus[0];
        r = blend(r, 0.0f, f, o); g = blend(g, 0.0f, f, o); b = blend(b, 0.0f, f, o);
    } else if (fn == 1) {
        const float f = jp.factor[1], o = jp.one_minus[1];
        r = blend(r, mean, f, o); g = blend(g, mean, f, o); b = blend(b, mean, f, o);
    } else if (fn == 2) {
        const float f = jp.factor[2], o = jp.one_minus[2];
        const float l = grey(r, g, b);
        r = blend(r, l, f, o); g = blend(g, l, f, o); b = blend(b, l, f, o);
    } else if (fn == 3) {
        adjust_hue(r, g, b, jp.factor[3]);
    }
}

// Both kernels keep the image's parameter record in shared memory (its arrays are indexed with run-time ids: in registers
// that would mean local memory) and, when the plane size allows (VEC: H * W % 4 == 0, 16-byte aligned planes), move four
// pixels per thread and channel with one 16-byte access.  The per-pixel arithmetic is the same either way.
template <bool VEC>
__global__ void __launch_bounds__(kJitThreads) jitter_mean_kernel(const float* __restrict__ src,
                                                                  const lh_jitter_params* __restrict__ params, int hw,
                                                                  double* __restrict__ ws) {
    __shared__ double s_part[kJitThreads / 32];
    __shared__ lh_jitter_params s_jp;
    const int n = blockIdx.y;
    if (threadIdx.x == 0) s_jp = params[n];
    __syncthreads();
    const lh_jitter_params& jp = s_jp;
    int cpos = -1;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (jp.order[k] == 1) cpos = k;
    if (!jp.apply || cpos < 0) return;                        // uniform per CTA
    const float* pr = src + (size_t)n * 3 * hw;
    double acc = 0.0;
    if (VEC) {
        const float4* p0 = reinterpret_cast<const float4*>(pr);
        const float4* p1 = reinterpret_cast<const float4*>(pr + hw);
        const float4* p2 = reinterpret_cast<const float4*>(pr + 2 * (size_t)hw);
        for (int i = blockIdx.x * kJitThreads + threadIdx.x; i < (hw >> 2); i += gridDim.x * kJitThreads) {
            const float4 r4 = __ldg(p0 + i), g4 = __ldg(p1 + i), b4 = __ldg(p2 + i);
            float r[4] = {r4.x, r4.y, r4.z, r4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
            for (int k = 0; k < cpos; ++k) {
                const int fn = jp.order[k];
#pragma unroll
                for (int c = 0; c < 4; ++c) apply_fn(fn, jp, 0.0f, r[c], g[c], b[c]);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) acc += (double)grey(r[c], g[c], b[c]);
        }
    } else {
        for (int i = blockIdx.x * kJitThreads + threadIdx.x; i < hw; i += gridDim.x * kJitThreads) {
            float r = __ldg(pr + i), g = __ldg(pr + hw + i), b = __ldg(pr + 2 * (size_t)hw + i);
            for (int k = 0; k < cpos; ++k) apply_fn(jp.order[k], jp, 0.0f, r, g, b);
            acc += (double)grey(r, g, b);
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kJitThreads / 32; ++w) t += s_part[w];
        ws[(size_t)n * LH_JITTER_PARTS + blockIdx.x] = t;
    }
}

// src and dst may be the same buffer (include/lighthand_b200.h): no __restrict__ / non-coherent loads on them; a thread
// reads its three channel values before it writes them, and no other thread touches those elements.
template <bool VEC>
__global__ void __launch_bounds__(kJitThreads) jitter_apply_kernel(const float* src,
                                                                   const lh_jitter_params* __restrict__ params, int hw,
                                                                   const double* __restrict__ ws, int normalise, float m0,
                                                                   float m1, float m2, float s0, float s1, float s2,
                                                                   float* dst) {
    __shared__ lh_jitter_params s_jp;
    __shared__ float s_mean;
    const int n = blockIdx.y;
    if (threadIdx.x == 0) {
        const lh_jitter_params jp0 = params[n];
        s_jp = jp0;
        float mean = 0.0f;
        if (jp0.apply) {
            bool has_contrast = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) has_contrast |= (jp0.order[k] == 1);
            if (has_contrast) {
                double t = 0.0;
                for (int k = 0; k < LH_JITTER_PARTS; ++k) t += ws[(size_t)n * LH_JITTER_PARTS + k];     // fixed order
                mean = (float)(t / (double)hw);
            }
        }
        s_mean = mean;
    }
    __syncthreads();
    const lh_jitter_params& jp = s_jp;
    const float mean = s_mean;
    const bool apply = jp.apply != 0;
    const float* pr = src + (size_t)n * 3 * hw;
    float* po = dst + (size_t)n * 3 * hw;
    if (VEC) {
        const float4* p0 = reinterpret_cast<const float4*>(pr);
        const float4* p1 = reinterpret_cast<const float4*>(pr + hw);
        const float4* p2 = reinterpret_cast<const float4*>(pr + 2 * (size_t)hw);
        float4* o0 = reinterpret_cast<float4*>(po);
        float4* o1 = reinterpret_cast<float4*>(po + hw);
        float4* o2 = reinterpret_cast<float4*>(po + 2 * (size_t)hw);
        for (int i = blockIdx.x * kJitThreads + threadIdx.x; i < (hw >> 2); i += gridDim.x * kJitThreads) {
            const float4 r4 = p0[i], g4 = p1[i], b4 = p2[i];
            float r[4] = {r4.x, r4.y, r4.z, r4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
            if (apply) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int fn = jp.order[k];
#pragma unroll
                    for (int c = 0; c < 4; ++c) apply_fn(fn, jp, mean, r[c], g[c], b[c]);
                }
            }
            if (normalise) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    r[c] = __fdiv_rn(__fsub_rn(r[c], m0), s0); g[c] = __fdiv_rn(__fsub_rn(g[c], m1), s1);
                    b[c] = __fdiv_rn(__fsub_rn(b[c], m2), s2);
                }
            }
            __stcs(o0 + i, make_float4(r[0], r[1], r[2], r[3]));
            __stcs(o1 + i, make_float4(g[0], g[1], g[2], g[3]));
            __stcs(o2 + i, make_float4(b[0], b[1], b[2], b[3]));
        }
    } else {
        for (int i = blockIdx.x * kJitThreads + threadIdx.x; i < hw; i += gridDim.x * kJitThreads) {
            float r = pr[i], g = pr[hw + i], b = pr[2 * (size_t)hw + i];
            if (apply) {
#pragma unroll
                for (int k = 0; k < 4; ++k) apply_fn(jp.order[k], jp, mean, r, g, b);
            }
            if (normalise) {
                r = __fdiv_rn(__fsub_rn(r, m0), s0); g = __fdiv_rn(__fsub_rn(g, m1), s1); b = __fdiv_rn(__fsub_rn(b, m2), s2);
            }
            __stcs(po + i, r); __stcs(po + hw + i, g); __stcs(po + 2 * (size_t)hw + i, b);
        }
    }
}

}  // namespace lh

extern "C" int lh_color_jitter_normalize(const float* src, const lh_jitter_params* params, int N, int H, int W,
                                         const float* mean_std, float* dst, double* ws, lh_stream_t stream) {
    if (N < 0 || H <= 0 || W <= 0 || (int64_t)H * W > (int64_t)1 << 28) return LH_EINVAL;
    if (N == 0) return LH_OK;
    if (!src || !params || !dst || !ws) return LH_ENULL;
    if (N > 65535) return LH_EINVAL;
    const int hw = H * W;
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (hw % 4 == 0) && lh::aligned16(src) && lh::aligned16(dst);
    if (vec) lh::jitter_mean_kernel<true><<<dim3(LH_JITTER_PARTS, N), lh::kJitThreads, 0, st>>>(src, params, hw, ws);
    else lh::jitter_mean_kernel<false><<<dim3(LH_JITTER_PARTS, N), lh::kJitThreads, 0, st>>>(src, params, hw, ws);
    LH_CHECK_LAUNCH();
    const float* ms = mean_std;
    const int per = vec ? hw / 4 : hw;
    const int gx = min((per + lh::kJitThreads - 1) / lh::kJitThreads, 64);
    const float m0 = ms ? ms[0] : 0.f, m1 = ms ? ms[1] : 0.f, m2 = ms ? ms[2] : 0.f;
    const float s0 = ms ? ms[3] : 1.f, s1 = ms ? ms[4] : 1.f, s2 = ms ? ms[5] : 1.f;
    if (vec) lh::jitter_apply_kernel<true><<<dim3(gx, N), lh::kJitThreads, 0, st>>>(src, params, hw, ws, ms ? 1 : 0, m0, m1, m2, s0, s1, s2, dst);
    else lh::jitter_apply_kernel<false><<<dim3(gx, N), lh::kJitThreads, 0, st>>>(src, params, hw, ws, ms ? 1 : 0, m0, m1, m2, s0, s1, s2, dst);
    LH_CHECK_LAUNCH();
    return LH_OK;
}
