// K5: PCK / EPE accumulation on device.
//
//   lh_pck_epe_batch   : PCK_2d_loss / PCK_2d_loss_visible (src/utils/loss.py:83-148),
//                        EPE_train / EPE (src/utils/loss.py:28-67)          -- float32, strict '>'
//   lh_pck_curve_accum : counting core of pred_eval / pred_test
//                        (src/utils/argparser.py:326-438)                   -- float64, strict '<'
//
// Integer counts must be bit-exact against torch-CPU / numpy, so every arithmetic step that
// feeds a comparison uses explicit round-to-nearest intrinsics (no FMA contraction), IEEE sqrt
// and IEEE division.  20*J bytes per image: latency/launch-bound, not bandwidth-bound; one warp
// owns one sample (lane = joint), counts are block-reduced and flushed with one atomic per block.
#include "lh_common.cuh"

namespace lh {

constexpr int kMThreads = 128;

// bbox diagonal exactly as the per-sample Python loop does it (builtin max/min keep the
// first element unless a later one compares strictly greater/smaller; float32 throughout).
__device__ __forceinline__ float bbox_diag_f32(const float* g, int J, int gs) {
    float mxx = g[0], mnx = g[0], mxy = g[1], mny = g[1];
    for (int j = 1; j < J; ++j) {
        const float x = g[j * gs], y = g[j * gs + 1];
        if (x > mxx) mxx = x;
        if (x < mnx) mnx = x;
        if (y > mxy) mxy = y;
        if (y < mny) mny = y;
    }
    const float w = __fsub_rn(mxx, mnx), h = __fsub_rn(mxy, mny);
    return __fsqrt_rn(__fadd_rn(__fmul_rn(w, w), __fmul_rn(h, h)));
}

__device__ __forceinline__ float dist_f32(const float* p, const float* g) {
    const float dx = __fsub_rn(g[0], p[0]), dy = __fsub_rn(g[1], p[1]);
    return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
}

// One warp per sample, one lane per joint (J <= 32; larger J walks in chunks of 32): the 20*J bytes of a
// sample are two coalesced loads, the bbox extremes are warp reductions, the counts come from ballots.
// Python's builtin max()/min() keep the first element unless a later one compares strictly greater /
// smaller, so a NaN survives only if it is the FIRST joint: reduce with NaN-ignoring fmaxf/fminf and
// put the first joint's NaN back.
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__global__ void __launch_bounds__(kMThreads)
pck_epe_batch_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int gs, long long N, int J,
                     float thr, int mode, int visible, unsigned long long* counts, double* sums) {
    __shared__ unsigned s_cnt[kMThreads / 32][3];
    __shared__ double s_es[kMThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long n = (long long)blockIdx.x * (kMThreads / 32) + warp;
    unsigned wrong = 0, total = 0, ecnt = 0;
    double esum = 0.0;
    if (n < N) {
        const float* g = gt + n * J * gs;
        const float* p = pred + n * J * 2;
        const float nan = __int_as_float(0x7fc00000);
        float mxx = nan, mnx = nan, mxy = nan, mny = nan;       // NaN is ignored by fmaxf / fminf
        for (int j0 = 0; j0 < J; j0 += 32) {
            const int j = j0 + lane;
            if (j < J) {
                const float x = g[j * gs], y = g[j * gs + 1];
                mxx = fmaxf(mxx, x); mnx = fminf(mnx, x); mxy = fmaxf(mxy, y); mny = fminf(mny, y);
            }
        }
        mxx = warp_max_f(mxx); mnx = warp_min_f(mnx); mxy = warp_max_f(mxy); mny = warp_min_f(mny);
        const float fx = g[0], fy = g[1];
        if (fx != fx) { mxx = fx; mnx = fx; }
        if (fy != fy) { mxy = fy; mny = fy; }
        const float w = __fsub_rn(mxx, mnx), h = __fsub_rn(mxy, mny);
        const float diag = __fsqrt_rn(__fadd_rn(__fmul_rn(w, w), __fmul_rn(h, h)));
        for (int j0 = 0; j0 < J; j0 += 32) {
            const int j = j0 + lane;
            if (j >= J || (visible && j == 0)) continue;
            float d = dist_f32(p + 2 * j, g + j * gs);
            const float dfull = d;
            bool vis_eq1 = true, vis_ne0 = true;
            if (visible) {
                const float v = g[j * gs + gs - 1];
                vis_eq1 = (v == 1.0f);
                vis_ne0 = !(v == 0.0f);
                d = __fmul_rn(d, vis_eq1 ? 1.0f : 0.0f);  // distance * vis_index (loss.py:100)
                total += vis_eq1 ? 1u : 0u;
            } else {
                total += 1u;
            }
            const float cmp = (mode == LH_PCK_PROPORTION) ? __fdiv_rn(d, diag) : d;
            wrong += (cmp > thr) ? 1u : 0u;
            // EPE: joints 1 .. J-2 (range(1, len(distance)) drops the last joint too, loss.py:44,64)
            if (j >= 1 && j <= J - 2 && (!visible || vis_ne0)) { esum += (double)dfull; ecnt += 1u; }
        }
    }
    wrong = __reduce_add_sync(0xffffffffu, wrong);
    total = __reduce_add_sync(0xffffffffu, total);
    ecnt = __reduce_add_sync(0xffffffffu, ecnt);
    esum = warp_sum(esum);
    if (lane == 0) { s_cnt[warp][0] = wrong; s_cnt[warp][1] = total; s_cnt[warp][2] = ecnt; s_es[warp] = esum; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned w = 0, t = 0, c = 0;
        double e = 0.0;
#pragma unroll
        for (int i = 0; i < kMThreads / 32; ++i) { w += s_cnt[i][0]; t += s_cnt[i][1]; c += s_cnt[i][2]; e += s_es[i]; }
        if (w) atomicAdd(counts + 0, (unsigned long long)w);
        if (t) atomicAdd(counts + 1, (unsigned long long)t);
        if (c) { atomicAdd(counts + 2, (unsigned long long)c); atomicAdd(sums + 0, e); }
        if (blockIdx.x == 0) atomicAdd(counts + 3, (unsigned long long)N);
    }
}

__global__ void __launch_bounds__(kMThreads)
pck_curve_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int gs,
                 const float* __restrict__ bbox, const int32_t* __restrict__ category, int n_cat,
                 long long N, int J, const double* __restrict__ thr, int K, int method, int use_vis,
                 unsigned long long* hist, unsigned long long* n_valid, unsigned long long* n_all,
                 double* sum_d) {
    extern __shared__ unsigned char smem_raw[];
    double* s_thr = reinterpret_cast<double*>(smem_raw);
    unsigned* s_hist = reinterpret_cast<unsigned*>(s_thr + K);
    const int nb = n_cat * (K + 1);
    for (int i = threadIdx.x; i < K; i += kMThreads) s_thr[i] = thr[i];
    for (int i = threadIdx.x; i < nb; i += kMThreads) s_hist[i] = 0u;
    __syncthreads();

    for (long long n0 = (long long)blockIdx.x * kMThreads; n0 < N; n0 += (long long)gridDim.x * kMThreads) {
        const long long n = n0 + threadIdx.x;
        if (n >= N) continue;
        const float* g = gt + n * J * gs;
        const float* p = pred + n * J * 2;
        int cat = category ? category[n] : 0;
        if (cat < 0 || cat >= n_cat) continue;  // unknown bucket: ignored, like a name missing from meta
        const double bb = (double)(bbox ? bbox[n] : bbox_diag_f32(g, J, gs));
        unsigned nv = 0;
        double sd = 0.0;
        for (int j = 0; j < J; ++j) {
            const double dx = __dsub_rn((double)g[j * gs], (double)p[2 * j]);
            const double dy = __dsub_rn((double)g[j * gs + 1], (double)p[2 * j + 1]);
            const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            sd += d;
            if (use_vis && !(g[j * gs + gs - 1] == 1.0f)) continue;
            nv += 1u;
            const double v = method ? __ddiv_rn(d, bb) : d;
            int lo = 0, hi = K;  // smallest k with v < thr[k]; NaN never satisfies '<' -> K
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (v < s_thr[mid]) hi = mid; else lo = mid + 1;
            }
            atomicAdd(&s_hist[cat * (K + 1) + lo], 1u);
        }
        atomicAdd(n_valid + cat, (unsigned long long)nv);
        atomicAdd(n_all + cat, (unsigned long long)J);
        atomicAdd(sum_d + cat, sd);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += kMThreads)
        if (s_hist[i]) atomicAdd(hist + i, (unsigned long long)s_hist[i]);
}


// one thread per (sample, joint)
__global__ void transform_joints_kernel(const float* __restrict__ joints, int js, const float* __restrict__ valid_in,
                                        const double* __restrict__ trans, const uint8_t* __restrict__ flip, int orig_w,
                                        const int32_t* __restrict__ swap, long long total, int J, int out_w, int out_h,
                                        float* __restrict__ out, int os, float* __restrict__ valid_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long n = i / J;
    const int j = (int)(i - n * J);
    const bool f = flip ? (flip[n] != 0) : false;
    const int sj = (f && swap) ? swap[j] : j;                     // slot j receives joint sj's data on a flip
    const float* g = joints + (n * J + sj) * js;
    float x = g[0];
    const float y = g[1];
    float v = valid_in ? valid_in[n * J + sj] : (js >= 3 ? g[2] : 1.0f);
    if (f) x = __fsub_rn(__fsub_rn((float)orig_w, x), 1.0f);      // float32, like numpy on a float32 array
    const double* t = trans + n * 6;
    // np.dot(trans, [x, y, 1.]) in float64, left to right; stored into the float32 joint array
    const double dx = __dadd_rn(__dadd_rn(__dmul_rn(t[0], (double)x), __dmul_rn(t[1], (double)y)), t[2]);
    const double dy = __dadd_rn(__dadd_rn(__dmul_rn(t[3], (double)x), __dmul_rn(t[4], (double)y)), t[5]);
    const float ox = (float)dx, oy = (float)dy;
    v = v * (float)(ox >= 0.f) * (float)(ox < (float)out_w) * (float)(oy >= 0.f) * (float)(oy < (float)out_h);
    float* o = out + i * os;
    o[0] = ox; o[1] = oy;
    if (os >= 3) o[2] = v;
    if (valid_out) valid_out[i] = v;
}

__global__ void bbox_diag_kernel(const float* __restrict__ gt, int gs, long long N, int J, float* __restrict__ out) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) out[n] = bbox_diag_f32(gt + n * J * gs, J, gs);
}

}  // namespace lh

using namespace lh;

extern "C" int lh_transform_joints(const float* joints, int joint_stride, const float* valid_in, const double* trans,
                                   const uint8_t* flip, int orig_w, const int32_t* swap, int64_t N, int J, int out_w,
                                   int out_h, float* joints_out, int out_stride, float* valid_out, lh_stream_t stream) {
    if (!joints || !trans || !joints_out) return LH_ENULL;
    if (N < 0 || J <= 0 || joint_stride < 2 || out_stride < 2 || out_w <= 0 || out_h <= 0) return LH_EINVAL;
    if (N == 0) return LH_OK;
    const long long total = (long long)N * J;
    transform_joints_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        joints, joint_stride, valid_in, trans, flip, orig_w, swap, total, J, out_w, out_h, joints_out, out_stride, valid_out);
    LH_CHECK_LAUNCH();
    return LH_OK;
}

extern "C" int lh_bbox_diag(const float* gt, int gt_stride, int64_t N, int J, float* out, lh_stream_t stream) {
    if (!gt || !out) return LH_ENULL;
    if (N < 0 || J <= 0 || gt_stride < 2) return LH_EINVAL;
    if (N == 0) return LH_OK;
    bbox_diag_kernel<<<(unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(gt, gt_stride, (long long)N, J, out);
    LH_CHECK_LAUNCH();
    return LH_OK;
}

extern "C" int lh_pck_epe_batch(const float* pred_xy, const float* gt, int gt_stride, int64_t N, int J,
                                float thr_f32, int mode, int visible, unsigned long long* counts,
                                double* sums, lh_stream_t stream) {
    if (!pred_xy || !gt || !counts || !sums) return LH_ENULL;
    if (N < 0 || J < 1 || gt_stride < 2) return LH_EINVAL;
    if (visible && gt_stride < 3) return LH_EINVAL;
    if (mode != LH_PCK_PROPORTION && mode != LH_PCK_MM) return LH_EINVAL;
    if (N == 0) return LH_OK;
    const long long blocks = (N + kMThreads / 32 - 1) / (kMThreads / 32);
    if (blocks > 0x7fffffffLL) return LH_EUNSUPPORTED;
    pck_epe_batch_kernel<<<(unsigned)blocks, kMThreads, 0, (cudaStream_t)stream>>>(
        pred_xy, gt, gt_stride, N, J, thr_f32, mode, visible, counts, sums);
    LH_CHECK_LAUNCH();
    return LH_OK;
}

extern "C" int lh_pck_curve_accum(const float* pred_xy, const float* gt, int gt_stride, const float* bbox_diag,
                                  const int32_t* category, int n_cat, int64_t N, int J,
                                  const double* thr_sorted, int K, int method, int use_vis,
                                  unsigned long long* hist, unsigned long long* n_valid,
                                  unsigned long long* n_all, double* sum_d, lh_stream_t stream) {
    if (!pred_xy || !gt || !thr_sorted || !hist || !n_valid || !n_all || !sum_d) return LH_ENULL;
    if (N < 0 || J < 1 || gt_stride < 2 || K < 1 || n_cat < 1) return LH_EINVAL;
    if (use_vis && gt_stride < 3) return LH_EINVAL;
    const size_t smem = sizeof(double) * (size_t)K + sizeof(unsigned) * (size_t)n_cat * (K + 1);
    if (smem > 40 * 1024) return LH_EUNSUPPORTED;
    if (N == 0) return LH_OK;
    long long blocks = (N + kMThreads - 1) / kMThreads;
    if (blocks > 148 * 16) blocks = 148 * 16;  // grid-stride: bounded number of histogram flushes
    pck_curve_kernel<<<(unsigned)blocks, kMThreads, smem, (cudaStream_t)stream>>>(
        pred_xy, gt, gt_stride, bbox_diag, category, n_cat, N, J, thr_sorted, K, method, use_vis,
        hist, n_valid, n_all, sum_d);
    LH_CHECK_LAUNCH();
    return LH_OK;
}
