// K5: PCK / EPE accumulation on device.
//
//   lh_pck_epe_batch   : PCK_2d_loss / PCK_2d_loss_visible (src/utils/loss.py:83-148),
//                        EPE_train / EPE (src/utils/loss.py:28-67)          -- float32, strict '>'
//   lh_pck_curve_accum : counting core of pred_eval / pred_test
//                        (src/utils/argparser.py:326-438)                   -- float64, strict '<'
//
// Integer counts must be bit-exact against torch-CPU / numpy, so every arithmetic step that
// feeds a comparison uses explicit round-to-nearest intrinsics (no FMA contraction), IEEE sqrt
// and IEEE division.  20*J bytes per image: latency/launch-bound, not bandwidth-bound; one
// thread owns one sample, counts are warp-reduced and flushed with one atomic per warp.
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

__global__ void __launch_bounds__(kMThreads)
pck_epe_batch_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int gs, long long N, int J,
                     float thr, int mode, int visible, unsigned long long* counts, double* sums) {
    const long long n = (long long)blockIdx.x * kMThreads + threadIdx.x;
    unsigned wrong = 0, total = 0, ecnt = 0;
    double esum = 0.0;
    if (n < N) {
        const float* g = gt + n * J * gs;
        const float* p = pred + n * J * 2;
        const float diag = bbox_diag_f32(g, J, gs);
        for (int j = visible ? 1 : 0; j < J; ++j) {
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
    // warp-aggregate, one atomic per warp per quantity
    for (int o = 16; o > 0; o >>= 1) {
        wrong += __shfl_xor_sync(0xffffffffu, wrong, o);
        total += __shfl_xor_sync(0xffffffffu, total, o);
        ecnt += __shfl_xor_sync(0xffffffffu, ecnt, o);
    }
    esum = warp_sum(esum);
    if ((threadIdx.x & 31) == 0) {
        if (wrong) atomicAdd(counts + 0, (unsigned long long)wrong);
        if (total) atomicAdd(counts + 1, (unsigned long long)total);
        if (ecnt) { atomicAdd(counts + 2, (unsigned long long)ecnt); atomicAdd(sums + 0, esum); }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(counts + 3, (unsigned long long)N);
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

}  // namespace lh

using namespace lh;

extern "C" int lh_pck_epe_batch(const float* pred_xy, const float* gt, int gt_stride, int64_t N, int J,
                                float thr_f32, int mode, int visible, unsigned long long* counts,
                                double* sums, lh_stream_t stream) {
    if (!pred_xy || !gt || !counts || !sums) return LH_ENULL;
    if (N < 0 || J < 1 || gt_stride < 2) return LH_EINVAL;
    if (visible && gt_stride < 3) return LH_EINVAL;
    if (mode != LH_PCK_PROPORTION && mode != LH_PCK_MM) return LH_EINVAL;
    if (N == 0) return LH_OK;
    const long long blocks = (N + kMThreads - 1) / kMThreads;
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
