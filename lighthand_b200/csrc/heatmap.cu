// K2 / K3 / K4 (+ K5) and their fusion: one streaming pass over [N,J,H,W] heat-maps.
//
//   gen_targets   (src/tools/dataset.py:165-212, src/utils/dataset_loader.py:22-53)
//   JointsMSELoss (src/utils/loss.py:306-325) forward + backward
//   get_max_preds (src/utils/loss.py:327-355) [+ quarter-pixel shift, inverse affine]
//   PCK_2d_loss / EPE_train (src/utils/loss.py:50-67,116-148) on the decoded joints (optional, fused tail)
//
// One CTA of 256 threads owns one (n,j) map.  Threads are laid out 2-D over the map (a thread keeps
// one 4-float column group and walks down the rows), so the Gaussian-window test in x and the table
// column offsets are loop invariants; H*W floats are streamed with 16-byte loads/stores, 4 independent
// vectors in flight per thread.  The target value is recomputed from the joint (a <=31x31 table in
// shared memory) instead of being re-read.  The argmax keeps the exact numpy order (first index on
// ties, NaN is the maximum) but tests one NaN-propagating max of the 4 lanes of a vector against the
// running best and only then falls into the exact per-element update, so the common case costs four
// instructions per vector.  Loss partials and the argmax are reduced with warp shuffles / REDUX and
// once across the 8 warps through shared memory.  The last CTA to finish (ticket) adds the per-map
// partials in index order (deterministic loss) and, if asked, folds the PCK / EPE counts of the batch
// into the running metric accumulators -- so a whole step's post-backbone work is ONE launch.
// HBM-bound by construction: it moves H*W*4 B in (+ H*W*4 B per requested output).
#include <cmath>

#include "lh_common.cuh"

namespace lh {

constexpr int kThreads = 256;      // scalar kernel
constexpr int kVecThreads = 128;   // vector kernel: 4 warps per map, 8 CTAs per SM
constexpr int kMaxTable = 31 * 31;

struct HeatmapParams {
    const float* pred;
    const float* joints;
    const float* table;
    const float* target_in;
    const float* weight_in;     // explicit [N,J] weights for lh_jmse_loss (nullable)
    const double* inv_affine;
    float* target_out;
    float* weight_out;
    float* grad;
    double* loss_f64;
    float* loss_f32;
    double* loss_meter;
    float* coords;
    float* maxvals;
    int32_t* argidx;
    unsigned long long* metric_counts;  // fused K5 (nullable)
    double* metric_sums;
    unsigned int* ticket;       // ws + 0
    double* cta_loss;           // ws + 16: one slot per CTA (sum of its maps' squared-error sums, in map order)
    double* cta_epe;            // per CTA: sum of float32 distances of joints 1..J-2
    unsigned int* cta_wrong;    // per CTA: number of failed PCK tests
    long long n_maps;
    int J, H, W, HW;
    int joint_stride, side, style, use_weight, subpixel;
    int pck_mode;
    float pck_thr;
    float stride_x, stride_y, out_scale;
    float inv_stride_x, inv_stride_y;   // exact reciprocal when the stride is a power of two, else 0
    float grad_scale;           // 1 / (N*J*HW)
    double loss_scale;          // 0.5 / (N*J*HW)
};

struct Window {
    int ulx, uly, brx, bry;  // paste window in map coordinates [ul, br)
    int valid;
};

// Python int(): truncation toward zero; out-of-int-range and non-finite handled by the caller.
__device__ __forceinline__ int trunc_i(float v) {
    v = fminf(fmaxf(v, -1.0e9f), 1.0e9f);
    return __float2int_rz(v);
}

// x / s in float32, IEEE-rounded; when s is a power of two (the reference's stride 4) the multiplication by
// the exact reciprocal is the same value bit for bit and skips the division sequence
__device__ __forceinline__ float div_stride(float v, float s, float inv_s) {
    return inv_s != 0.f ? __fmul_rn(v, inv_s) : __fdiv_rn(v, s);
}

__device__ __forceinline__ Window joint_window(const HeatmapParams& p, float jx, float jy) {
    Window w;
    w.valid = 1;
    if (p.style == LH_TARGET_HRNET) {
        // mu = int(joint / stride + 0.5) in float32  (src/tools/dataset.py:178-179)
        const float fx = __fadd_rn(div_stride(jx, p.stride_x, p.inv_stride_x), 0.5f);
        const float fy = __fadd_rn(div_stride(jy, p.stride_y, p.inv_stride_y), 0.5f);
        const int tmp = (p.side - 1) >> 1;
        const int mx = trunc_i(fx), my = trunc_i(fy);
        w.ulx = mx - tmp; w.uly = my - tmp;
        w.brx = mx + tmp + 1; w.bry = my + tmp + 1;
        if (!(fx == fx) || !(fy == fy)) w.valid = 0;  // int(nan) raises in the reference; no target here
        if (w.ulx >= p.W || w.uly >= p.H || w.brx < 0 || w.bry < 0) w.valid = 0;  // dataset.py:183-186
    } else {
        // GenerateHeatmap: p = joint/4; gate on x > 0 only; int() truncation (dataset_loader.py:38-41)
        const float px = div_stride(jx, p.stride_x, p.inv_stride_x), py = div_stride(jy, p.stride_y, p.inv_stride_y);
        const int half = (p.side - 1) >> 1;  // 3*sigma + 1
        if (!(px > 0.0f) || !(py == py)) {
            w.valid = 0; w.ulx = w.uly = w.brx = w.bry = 0;
        } else {
            const int x = trunc_i(px), y = trunc_i(py);
            if (x < 0 || y < 0 || x >= p.W || y >= p.H) w.valid = 0;
            w.ulx = x - half; w.uly = y - half;
            w.brx = x + half + 1; w.bry = y + half + 1;
        }
    }
    return w;
}

__device__ __forceinline__ float target_at(const Window& w, const float* s_tab, int side, int y, int x) {
    if (y >= w.uly && y < w.bry && x >= w.ulx && x < w.brx)
        return s_tab[(y - w.uly) * side + (x - w.ulx)];
    return 0.0f;
}

__device__ __forceinline__ float max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// bbox diagonal exactly as the reference's per-sample Python loop (builtin max/min keep the first
// element unless a later one compares strictly greater/smaller; float32) -- src/utils/loss.py:122-127
__device__ __forceinline__ float bbox_diag_seq(const float* g, int J, int gs) {
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

// ------------------------------------------------------------------------------------------------
// Tail of one map: final argmax -> coordinates (+ sub-pixel / inverse affine), outputs, fused metric
// record.  Runs on warp 0 (lanes cooperate on the bbox diagonal); `jl` = this lane's ground-truth
// joint of sample n (lane < J), `gj` = the map's own joint.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void map_tail(const HeatmapParams& p, long long m, const float* pred, int bi, float mv,
                                         bool do_metric, float2 jl, float2 jfirst, float2 gj, int lane,
                                         double& epe_acc, unsigned int& wrong_acc) {
    float diag = 0.f;
    if (do_metric && p.pck_mode == LH_PCK_PROPORTION) {
        // builtin max()/min() of the reference keep the first element unless a later one compares strictly
        // greater / smaller: a NaN survives only as the first joint -> NaN-ignoring reduce, then patch.
        const float nan = __int_as_float(0x7fc00000);
        float mxx = nan, mnx = nan, mxy = nan, mny = nan;
        if (p.J <= 32) {
            if (lane < p.J) { mxx = mnx = jl.x; mxy = mny = jl.y; }
        } else {
            const float* g = p.joints + (m / p.J) * p.J * p.joint_stride;
            for (int j = lane; j < p.J; j += 32) {
                const float x = g[j * p.joint_stride], y = g[j * p.joint_stride + 1];
                mxx = fmaxf(mxx, x); mnx = fminf(mnx, x); mxy = fmaxf(mxy, y); mny = fminf(mny, y);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mxx = fmaxf(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
            mnx = fminf(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
            mxy = fmaxf(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
            mny = fminf(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        }
        if (jfirst.x != jfirst.x) { mxx = jfirst.x; mnx = jfirst.x; }
        if (jfirst.y != jfirst.y) { mxy = jfirst.y; mny = jfirst.y; }
        const float w = __fsub_rn(mxx, mnx), h = __fsub_rn(mxy, mny);
        diag = __fsqrt_rn(__fadd_rn(__fmul_rn(w, w), __fmul_rn(h, h)));
    }
    if (lane != 0) return;
    const int iy = bi / p.W, ix = bi - iy * p.W;
    float cx = (float)ix, cy = (float)iy;
    if (!(mv > 0.0f)) { cx = 0.f; cy = 0.f; }       // loss.py:350-354 (NaN > 0 is false)
    if (p.subpixel) {
        // upstream HRNet get_final_preds: px = floor(x + .5)
        const int px = (int)floorf(cx + 0.5f), py = (int)floorf(cy + 0.5f);
        if (px > 1 && px < p.W - 1 && py > 1 && py < p.H - 1) {
            const float dx = __fsub_rn(pred[py * p.W + px + 1], pred[py * p.W + px - 1]);
            const float dy = __fsub_rn(pred[(py + 1) * p.W + px], pred[(py - 1) * p.W + px]);
            cx += (dx > 0.f) ? 0.25f : ((dx < 0.f) ? -0.25f : (dx == dx ? 0.f : dx));
            cy += (dy > 0.f) ? 0.25f : ((dy < 0.f) ? -0.25f : (dy == dy ? 0.f : dy));
        }
    }
    cx = __fmul_rn(cx, p.out_scale);
    cy = __fmul_rn(cy, p.out_scale);
    if (p.inv_affine) {
        const double* A = p.inv_affine + (m / p.J) * 6;
        const double x = (double)cx, y = (double)cy;
        const double ox = __dadd_rn(__dadd_rn(__dmul_rn(A[0], x), __dmul_rn(A[1], y)), A[2]);
        const double oy = __dadd_rn(__dadd_rn(__dmul_rn(A[3], x), __dmul_rn(A[4], y)), A[5]);
        cx = (float)ox; cy = (float)oy;
    }
    p.coords[2 * m] = cx;
    p.coords[2 * m + 1] = cy;
    if (p.maxvals) p.maxvals[m] = mv;
    if (p.argidx) p.argidx[m] = bi;
    if (do_metric) {
        // PCK_2d_loss / EPE_train on this joint: float32, no FMA contraction, IEEE sqrt and divide
        const float dx = __fsub_rn(gj.x, cx), dy = __fsub_rn(gj.y, cy);
        const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        const float cmp = (p.pck_mode == LH_PCK_PROPORTION) ? __fdiv_rn(d, diag) : d;
        wrong_acc += (cmp > p.pck_thr) ? 1u : 0u;
        // EPE_train keeps joints 1 .. J-2 (range(1, len(distance)) drops the last one too)
        const int j = (int)(m % p.J);
        if (j >= 1 && j <= p.J - 2) epe_acc += (double)d;
    }
}

// Every CTA leaves one slot (its maps' sums, accumulated in map order); the last CTA to finish adds the
// slots in index order -- deterministic for a given grid -- and folds the batch's PCK / EPE counts into
// the running accumulators.
template <bool DO_LOSS, int NT>
__device__ __forceinline__ void final_reduce(const HeatmapParams& p, bool do_metric, double* s_a, double* s_b,
                                             unsigned int* s_c, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    double s = 0.0, es = 0.0;
    unsigned int wrong = 0;
#pragma unroll 4
    for (unsigned int i = tid; i < gridDim.x; i += NT) {   // a few rounds: the grid is at most SMs x 8 CTAs
        if (DO_LOSS) s += __ldcg(p.cta_loss + i);
        if (do_metric) { wrong += __ldcg(p.cta_wrong + i); es += __ldcg(p.cta_epe + i); }
    }
    s = warp_sum(s);
    es = warp_sum(es);
    wrong = __reduce_add_sync(0xffffffffu, wrong);
    if (lane == 0) { s_a[warp] = s; s_b[warp] = es; s_c[warp] = wrong; }
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0, etot = 0.0;
        unsigned int wtot = 0;
#pragma unroll
        for (int w = 0; w < NT / kWarp; ++w) { tot += s_a[w]; etot += s_b[w]; wtot += s_c[w]; }
        if (DO_LOSS) {
            const double loss = tot * p.loss_scale;
            if (p.loss_f64) *p.loss_f64 = loss;
            if (p.loss_f32) *p.loss_f32 = (float)loss;
            if (p.loss_meter) {  // AverageMeter.update(loss.item(), batch_size): float32 loss, stream-ordered
                const double nb = (double)(p.n_maps / p.J);
                p.loss_meter[0] += (double)(float)loss * nb;
                p.loss_meter[1] += nb;
            }
        }
        if (do_metric) {
            const unsigned long long N = (unsigned long long)(p.n_maps / p.J);
            const unsigned long long epe_cnt = p.J >= 3 ? N * (unsigned long long)(p.J - 2) : 0ull;
            atomicAdd(p.metric_counts + 0, (unsigned long long)wtot);
            atomicAdd(p.metric_counts + 1, (unsigned long long)p.n_maps);
            atomicAdd(p.metric_counts + 2, epe_cnt);
            atomicAdd(p.metric_counts + 3, N);
            atomicAdd(p.metric_sums + 0, etot);
        }
        *p.ticket = 0u;  // leave the workspace clean for the next launch
    }
}

// ------------------------------------------------------------------------------------------------
// Vectorised persistent kernel.  TSRC: 0 = no target (decode only), 1 = generate from joints, 2 = read
// target_in.  FLAGS (compile time, so absent outputs cost nothing): F_GRAD write d loss / d pred,
// F_TOUT write the target, F_DEC argmax decode, F_WGT per-map weights.
// Each CTA walks maps m = blockIdx.x, += gridDim.x; per map two barriers:
//   A  after the streaming pass: per-warp loss partial + NaN-propagating maximum
//   B  after the winner search: smallest index holding the maximum (numpy's first-index tie-break)
// Warp 0 then finishes the map (coordinates, outputs, metric record) while the other warps are already
// streaming the next one; reduction scratch is triple-buffered by map count (see s_sum below).
// ------------------------------------------------------------------------------------------------
constexpr int F_GRAD = 1, F_TOUT = 2, F_DEC = 4, F_WGT = 8;

template <bool PRED, int TSRC, int FLAGS>
struct VecBody {
    static constexpr bool GRAD = (FLAGS & F_GRAD) != 0, TOUT = (FLAGS & F_TOUT) != 0, DEC = (FLAGS & F_DEC) != 0,
                          WGT = (FLAGS & F_WGT) != 0, LOSS = PRED && TSRC != 0;
    // one 4-float vector at element offset e of the current map (row y, fixed column group)
    __device__ static __forceinline__ void run(const float4& pv, const float4& tv, int y, int e, const Window& win,
                                               const float* s_tab, int side, const int (&tx)[4], const bool (&xin)[4],
                                               bool any_x, float wgt, float gs, float* tout, float* grad,
                                               float& acc0, float& acc1, float& tmax, int& tvec) {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        if (TSRC == 1) {
            const unsigned ry = (unsigned)(y - win.uly);
            if (any_x && ry < (unsigned)(win.bry - win.uly)) {
                const float* trow = s_tab + ry * side;
#pragma unroll
                for (int k = 0; k < 4; ++k) if (xin[k]) t[k] = trow[tx[k]];
            }
            if (TOUT) stg_stream4(tout + e, make_float4(t[0], t[1], t[2], t[3]));
        } else if (TSRC == 2) {
            t[0] = tv.x; t[1] = tv.y; t[2] = tv.z; t[3] = tv.w;
        }
        if (PRED) {
            const float v[4] = {pv.x, pv.y, pv.z, pv.w};
            if (LOSS) {
                float d[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) d[k] = WGT ? (v[k] - t[k]) * wgt : (v[k] - t[k]);
                acc0 = fmaf(d[0], d[0], acc0); acc1 = fmaf(d[1], d[1], acc1);
                acc0 = fmaf(d[2], d[2], acc0); acc1 = fmaf(d[3], d[3], acc1);
                if (GRAD) stg_cs4(grad + e, make_float4(d[0] * gs, d[1] * gs, d[2] * gs, d[3] * gs));
            }
            if (DEC) {
                // vectors are visited in increasing index order: replace only on "strictly greater", or when
                // the first NaN shows up (NaN is numpy's maximum); never once tmax is NaN.
                const float vmax = max_nan(max_nan(v[0], v[1]), max_nan(v[2], v[3]));
                const bool upd = !(vmax <= tmax) && (tmax == tmax);
                tmax = upd ? vmax : tmax;
                tvec = upd ? e : tvec;
            }
        }
    }
};

template <bool PRED, int TSRC, int FLAGS>
__global__ void __launch_bounds__(kVecThreads, TSRC == 2 ? 4 : 8) heatmap_vec_kernel(const HeatmapParams p) {
    using Body = VecBody<PRED, TSRC, FLAGS>;
    constexpr bool GRAD = Body::GRAD, TOUT = Body::TOUT, DEC = Body::DEC, WGT = Body::WGT, LOSS = Body::LOSS;
    __shared__ float s_tab[TSRC == 1 ? kMaxTable : 1];
    // reduction scratch, TRIPLE-buffered by map count: after barrier A of map i warp 0 still reads map i-1's slots
    // ("finish the previous map one barrier late") while the other warps already write map i+1's
    __shared__ double s_sum[3][kVecThreads / kWarp];
    __shared__ __align__(16) float s_max[3][kVecThreads / kWarp];
    __shared__ __align__(16) int s_idx[3][kVecThreads / kWarp];
    static_assert(kVecThreads / kWarp == 4, "the float4 / int4 reads of s_max / s_idx assume 4 warps");
    __shared__ double s_fin[kVecThreads / kWarp];
    __shared__ unsigned int s_finc[kVecThreads / kWarp];
    __shared__ int s_last;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    if (TSRC == 1) {
        for (int i = tid; i < p.side * p.side; i += kVecThreads) s_tab[i] = p.table[i];
        __syncthreads();
    }
    const bool do_metric = DEC && p.metric_counts != nullptr;

    constexpr int U = 4;
    const int W = p.W, H = p.H, side = p.side;
    const int vw = W >> 2;                   // vectors per row (<= 256, checked by the launcher)
    const int rp = kVecThreads / vw;            // rows per pass
    const int r0 = tid / vw, c = tid - r0 * vw;
    const int x = 4 * c;
    const bool active = r0 < rp;
    const int e0 = r0 * W + x;               // first element of this thread, estep elements between its rows
    const int estep = rp * W;
    const int G = (H + U * rp - 1) / (U * rp);   // groups of U rows-per-thread in one map
    const float ninf = __int_as_float(0xff800000);

    __shared__ double s_cta[2];              // running sums of this CTA (thread 0): squared error, EPE distances
    __shared__ unsigned int s_ctaw;          // failed PCK tests
    if (tid == 0) { s_cta[0] = 0.0; s_cta[1] = 0.0; s_ctaw = 0u; }

    // software pipeline: the loads of work item k+1 (next row group, possibly of the next map) are issued
    // right after item k's registers are consumed, so they are in flight across the reductions, the barrier
    // and the tail of the map.  The tail of map i runs after barrier A of map i+1 (one barrier per map).
    float4 pv[U], tv[U];
#define LH_ISSUE(MM, GG)                                                                        \
    {                                                                                           \
        const size_t base_ = (size_t)(MM) * p.HW + e0 + (GG) * U * estep;                       \
        const int y_ = r0 + (GG) * U * rp;                                                      \
        _Pragma("unroll") for (int u = 0; u < U; ++u) {                                         \
            if (active && y_ + u * rp < H) {                                                    \
                if (PRED) pv[u] = ldg_stream4(p.pred + base_ + u * estep);                      \
                if (TSRC == 2) tv[u] = ldg_stream4(p.target_in + base_ + u * estep);            \
            }                                                                                   \
        }                                                                                       \
    }
    // warp 0: finish map PM (its reductions are visible: a barrier has passed); joints are re-read (L2 hits)
#define LH_FINISH_PREV(PM, PPAR)                                                                \
    {                                                                                           \
        if (LOSS && lane == 0) {                                                                \
            double sacc = 0.0;                                                                  \
            _Pragma("unroll") for (int w = 0; w < kVecThreads / kWarp; ++w) sacc += s_sum[PPAR][w]; \
            s_cta[0] += sacc;                                                                   \
        }                                                                                       \
        if (DEC) {                                                                              \
            const float4 ma = *reinterpret_cast<const float4*>(&s_max[PPAR][0]);                \
            const float Mp = max_nan(max_nan(ma.x, ma.y), max_nan(ma.z, ma.w));                 \
            const int4 ia = *reinterpret_cast<const int4*>(&s_idx[PPAR][0]);                    \
            int bi = min(min(ia.x, ia.y), min(ia.z, ia.w));                                     \
            if (bi == 0x7fffffff) bi = 0; /* HW == 0 only */                                    \
            float2 pgj = make_float2(0.f, 0.f), pjl = pgj, pjfirst = pgj;                       \
            if (do_metric) {                                                                    \
                const float* jp_ = p.joints + (PM) * p.joint_stride;                            \
                pgj = make_float2(jp_[0], jp_[1]);                                              \
                const float* gp_ = p.joints + ((PM) / p.J) * p.J * p.joint_stride;              \
                pjfirst = make_float2(gp_[0], gp_[1]);                                          \
                if (lane < p.J && p.J <= 32) pjl = make_float2(gp_[lane * p.joint_stride], gp_[lane * p.joint_stride + 1]); \
            }                                                                                   \
            double epe_ = 0.0; unsigned int wr_ = 0u;                                           \
            map_tail(p, (PM), p.pred + (size_t)(PM) * p.HW, bi, Mp, do_metric, pjl, pjfirst, pgj, lane, epe_, wr_); \
            if (lane == 0 && do_metric) { s_cta[1] += epe_; s_ctaw += wr_; }                    \
        }                                                                                       \
    }

    long long m = blockIdx.x, pm = -1;
    int g = 0, par = 0;
    if (m < p.n_maps) LH_ISSUE(m, 0)

    // state of the map being streamed
    int ulx = 0, uly = 0, brx = 0, bry = 0;
    float wgt = 1.0f, gs = p.grad_scale;
    float acc0 = 0.f, acc1 = 0.f, tmax = ninf;
    int tvec = e0;

    while (m < p.n_maps) {
        if (g == 0) {
            // ---- per-map setup -------------------------------------------------------------------------
            ulx = uly = brx = bry = 0;
            bool valid = false;
            if (TSRC == 1) {
                const float* jp = p.joints + m * p.joint_stride;
                const Window win = joint_window(p, jp[0], jp[1]);
                valid = win.valid != 0;
                if (valid) { ulx = win.ulx; uly = win.uly; brx = win.brx; bry = win.bry; }
                if (p.weight_out && tid == 0) p.weight_out[m] = valid ? 1.0f : 0.0f;
            }
            if (WGT) {
                wgt = 1.0f;
                if (p.weight_in) wgt = p.weight_in[m];
                else if (TSRC == 1) wgt = valid ? 1.0f : 0.0f;
                gs = p.grad_scale * wgt;                         // grad = (pw - tw) * w / count
            }
            acc0 = acc1 = 0.f; tmax = ninf; tvec = e0;
        }
        // ---- consume the group in registers ----------------------------------------------------------------
        if (active) {
            Window win;
            win.ulx = ulx; win.uly = uly; win.brx = brx; win.bry = bry; win.valid = 1;
            int tx[4];
            bool xin[4], any_x = false;
#pragma unroll
            for (int k = 0; k < 4; ++k) {                        // x part of the Gaussian window test
                tx[k] = x + k - ulx;
                xin[k] = (TSRC == 1) && (x + k >= ulx) && (x + k < brx);
                any_x |= xin[k];
            }
            const size_t base = (size_t)m * p.HW;
            float* tout = TOUT ? p.target_out + base : nullptr;
            float* grad = GRAD ? p.grad + base : nullptr;
            const int y = r0 + g * U * rp, e = e0 + g * U * estep;
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (y + u * rp < H)
                    Body::run(pv[u], tv[u], y + u * rp, e + u * estep, win, s_tab, side, tx, xin, any_x, wgt, gs,
                              tout, grad, acc0, acc1, tmax, tvec);
        }
        // ---- issue the next item's loads ---------------------------------------------------------------------
        long long m2 = m;
        int g2 = g + 1;
        if (g2 == G) { g2 = 0; m2 = m + gridDim.x; }
        if (m2 < p.n_maps) LH_ISSUE(m2, g2)

        if (g2 == 0 && PRED) {
            // ---- map finished: warp partials, barrier A, winner search ---------------------------------------
            float wsum = acc0 + acc1, wmax = tmax;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                if (LOSS) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
                if (DEC) wmax = max_nan(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
            }
            if (lane == 0) { if (LOSS) s_sum[par][warp] = (double)wsum; if (DEC) s_max[par][warp] = wmax; }
            __syncthreads();
            // the barrier also published the previous map's winner indices: warp 0 finishes that map now
            if (warp == 0 && pm >= 0) LH_FINISH_PREV(pm, (par + 2) % 3)
            if (DEC) {
                const float4 ma = *reinterpret_cast<const float4*>(&s_max[par][0]);
                const float M = max_nan(max_nan(ma.x, ma.y), max_nan(ma.z, ma.w));
                const bool mnan = M != M;
                int cand = 0x7fffffff;
                // threads whose own maximum IS the map maximum look up its first position inside their vector
                // (threads that own no row of a map with H < rows-per-pass keep tvec outside the map: r0 < H)
                if (active && r0 < H && (mnan ? (tmax != tmax) : (tmax == M))) {
                    const float4 q = *reinterpret_cast<const float4*>(p.pred + (size_t)m * p.HW + tvec);
                    const float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int k = 3; k >= 0; --k) if (mnan ? (v[k] != v[k]) : (v[k] == M)) cand = tvec + k;
                }
                cand = __reduce_min_sync(0xffffffffu, cand);
                if (lane == 0) s_idx[par][warp] = cand;
            }
            pm = m;
            par = (par + 1) % 3;
        }
        m = m2; g = g2;
    }

    if (!PRED) return;
    __syncthreads();                                  // publishes the last map's winner indices
    if (warp == 0 && pm >= 0) LH_FINISH_PREV(pm, (par + 2) % 3)
#undef LH_ISSUE
#undef LH_FINISH_PREV
    if ((LOSS && (p.loss_f64 || p.loss_f32 || p.loss_meter)) || do_metric) {
        if (tid == 0) {
            p.cta_loss[blockIdx.x] = s_cta[0]; p.cta_epe[blockIdx.x] = s_cta[1]; p.cta_wrong[blockIdx.x] = s_ctaw;
            __threadfence();
            const unsigned int t = atomicAdd(p.ticket, 1u);
            s_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            final_reduce<LOSS, kVecThreads>(p, do_metric, s_sum[0], s_fin, s_finc, tid);
        }
    }
}

// Generic scalar kernel for shapes the vector kernel does not take (W % 4 != 0, unaligned pointers):
// one CTA per map, exact per-element ordering key.
template <bool PRED, int TSRC>
__global__ void __launch_bounds__(kThreads) heatmap_scalar_kernel(const HeatmapParams p) {
    __shared__ float s_tab[TSRC == 1 ? kMaxTable : 1];
    __shared__ double s_sum[kThreads / kWarp];
    __shared__ uint32_t s_key[kThreads / kWarp];
    __shared__ int s_idx[kThreads / kWarp];
    __shared__ double s_fin[kThreads / kWarp];
    __shared__ unsigned int s_finc[kThreads / kWarp];
    __shared__ int s_last;

    const int tid = threadIdx.x;
    const long long m = blockIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    Window win;
    win.valid = 0; win.ulx = win.uly = win.brx = win.bry = 0;
    float2 gj = make_float2(0.f, 0.f), jl = gj, jfirst = gj;
    const bool do_loss = PRED && TSRC != 0;
    const bool do_dec = PRED && p.coords != nullptr;
    const bool do_metric = do_dec && p.metric_counts != nullptr;
    if (TSRC == 1 || do_metric) {
        const float* jp = p.joints + m * p.joint_stride;
        gj = make_float2(jp[0], jp[1]);
    }
    if (TSRC == 1) {
        for (int i = tid; i < p.side * p.side; i += kThreads) s_tab[i] = p.table[i];
        win = joint_window(p, gj.x, gj.y);
        if (!win.valid) { win.ulx = win.uly = 0; win.brx = win.bry = 0; }
        __syncthreads();
    }
    if (do_metric && warp == 0 && p.pck_mode == LH_PCK_PROPORTION) {
        const float* g = p.joints + (m / p.J) * p.J * p.joint_stride;
        jfirst = make_float2(g[0], g[1]);
        if (lane < p.J && p.J <= 32) jl = make_float2(g[lane * p.joint_stride], g[lane * p.joint_stride + 1]);
    }
    float wgt = 1.0f;
    if (p.use_weight) {
        if (p.weight_in) wgt = p.weight_in[m];
        else if (TSRC == 1) wgt = win.valid ? 1.0f : 0.0f;
    }
    if (TSRC == 1 && p.weight_out && tid == 0) p.weight_out[m] = win.valid ? 1.0f : 0.0f;

    const size_t base = (size_t)m * p.HW;
    const float* pred = PRED ? p.pred + base : nullptr;
    const float* tin = (TSRC == 2) ? p.target_in + base : nullptr;
    float* tout = p.target_out ? p.target_out + base : nullptr;
    float* grad = p.grad ? p.grad + base : nullptr;

    float acc0 = 0.f;
    uint32_t best_key = 0u;
    int best_idx = 0x7fffffff;
    const float gs = p.grad_scale * wgt;
    for (int e = tid; e < p.HW; e += kThreads) {
        float t = 0.f;
        if (TSRC == 1) {
            const int y = e / p.W, x = e - y * p.W;
            t = target_at(win, s_tab, p.side, y, x);
            if (tout) tout[e] = t;
        } else if (TSRC == 2) {
            t = tin[e];
        }
        if (PRED) {
            const float v = pred[e];
            if (do_loss) {
                const float d = (v - t) * wgt;
                acc0 = fmaf(d, d, acc0);
                if (grad) grad[e] = d * gs;
            }
            if (do_dec) {
                const uint32_t key = order_key(v);
                if (key > best_key) { best_key = key; best_idx = e; }
            }
        }
    }
    if (!PRED) return;

    double wsum = do_loss ? warp_sum((double)acc0) : 0.0;
    uint32_t kmax = 0u;
    int imin = 0x7fffffff;
    if (do_dec) {
        kmax = __reduce_max_sync(0xffffffffu, best_key);
        imin = __reduce_min_sync(0xffffffffu, best_key == kmax ? best_idx : 0x7fffffff);
    }
    if (lane == 0) { s_sum[warp] = wsum; s_key[warp] = kmax; s_idx[warp] = imin; }
    __syncthreads();
    double cta_loss = 0.0, cta_epe = 0.0;
    unsigned int cta_wrong = 0;
    if (warp == 0) {
        if (do_loss && lane == 0) {
#pragma unroll
            for (int w = 0; w < kThreads / kWarp; ++w) cta_loss += s_sum[w];
        }
        if (do_dec) {
            uint32_t bk = 0u; int bi = 0x7fffffff;
#pragma unroll
            for (int w = 0; w < kThreads / kWarp; ++w) {
                if (s_key[w] > bk || (s_key[w] == bk && s_idx[w] < bi)) { bk = s_key[w]; bi = s_idx[w]; }
            }
            if (bi == 0x7fffffff) bi = 0;
            map_tail(p, m, pred, bi, pred[bi], do_metric, jl, jfirst, gj, lane, cta_epe, cta_wrong);
        }
    }
    if ((do_loss && (p.loss_f64 || p.loss_f32 || p.loss_meter)) || do_metric) {
        __syncthreads();
        if (tid == 0) {
            p.cta_loss[blockIdx.x] = cta_loss; p.cta_epe[blockIdx.x] = cta_epe; p.cta_wrong[blockIdx.x] = cta_wrong;
            __threadfence();
            const unsigned int t = atomicAdd(p.ticket, 1u);
            s_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            if (do_loss) final_reduce<true, kThreads>(p, do_metric, s_sum, s_fin, s_finc, tid);
            else final_reduce<false, kThreads>(p, do_metric, s_sum, s_fin, s_finc, tid);
        }
    }
}

// resident CTAs per SM x SM count of the current device, per kernel instantiation (cached)
template <typename K>
static int persistent_grid(K kernel, int* cache) {
    if (*cache > 0) return *cache;
    int dev = 0, sms = 0, occ = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kVecThreads, 0) != cudaSuccess) return 0;
    if (occ < 1) occ = 1;
    *cache = sms * occ;
    return *cache;
}

template <bool PRED, int TSRC, int FLAGS>
static int launch_vec(const HeatmapParams& p, cudaStream_t st) {
    static int cache = 0;   // one device model per process (B200); benign race: same value
    const int cap = persistent_grid(heatmap_vec_kernel<PRED, TSRC, FLAGS>, &cache);
    if (cap <= 0) return (int)cudaGetLastError();
    const unsigned grid = (unsigned)(p.n_maps < (long long)cap ? p.n_maps : (long long)cap);
    heatmap_vec_kernel<PRED, TSRC, FLAGS><<<grid, kVecThreads, 0, st>>>(p);
    return LH_OK;
}

// runtime output set -> compile-time FLAGS (only the combinations the entry points can produce)
template <int TSRC>
static int dispatch_pred(const HeatmapParams& p, cudaStream_t st) {
    const bool g = p.grad != nullptr, d = p.coords != nullptr, w = p.use_weight != 0;
    const bool t = (TSRC == 1) && p.target_out != nullptr;
#define LH_CASE(G, T, D, Wt)                                                                              \
    if (g == G && t == T && d == D && w == Wt)                                                            \
        return launch_vec<true, TSRC, (G ? F_GRAD : 0) | (T ? F_TOUT : 0) | (D ? F_DEC : 0) | (Wt ? F_WGT : 0)>(p, st);
    if constexpr (TSRC == 0) {
        LH_CASE(false, false, true, false)
        return LH_EINVAL;
    } else {
    LH_CASE(true, false, true, false)  LH_CASE(true, false, false, false) LH_CASE(false, false, true, false)
    LH_CASE(false, false, false, false) LH_CASE(true, false, true, true)  LH_CASE(true, false, false, true)
    LH_CASE(false, false, true, true)  LH_CASE(false, false, false, true)
    if constexpr (TSRC == 1) {
        LH_CASE(true, true, true, false)  LH_CASE(true, true, false, false) LH_CASE(false, true, true, false)
        LH_CASE(false, true, false, false) LH_CASE(true, true, true, true)  LH_CASE(true, true, false, true)
        LH_CASE(false, true, true, true)  LH_CASE(false, true, false, true)
    }
    return LH_EINVAL;
    }
#undef LH_CASE
}

template <bool PRED, int TSRC>
static int launch(HeatmapParams& p, bool vec, cudaStream_t st) {
    if (p.n_maps == 0) return LH_OK;
    if (p.n_maps > 0x7fffffffLL) return LH_EUNSUPPORTED;
    auto pow2_inv = [](float s) { int e = 0; return std::frexp(s, &e) == 0.5f ? 1.0f / s : 0.0f; };
    p.inv_stride_x = p.stride_x > 0.f ? pow2_inv(p.stride_x) : 0.f;
    p.inv_stride_y = p.stride_y > 0.f ? pow2_inv(p.stride_y) : 0.f;
    int rc = LH_OK;
    if (vec) {
        rc = PRED ? dispatch_pred<TSRC>(p, st) : launch_vec<false, 1, F_TOUT>(p, st);
        if (rc) return rc;
    } else {
        heatmap_scalar_kernel<PRED, TSRC><<<(unsigned)p.n_maps, kThreads, 0, st>>>(p);
    }
    LH_CHECK_LAUNCH();
    return LH_OK;
}

static bool can_vec(const HeatmapParams& p) {
    if ((p.W & 3) != 0 || (p.W >> 2) > kVecThreads) return false;
    const void* ptrs[] = {p.pred, p.target_in, p.target_out, p.grad};
    for (const void* q : ptrs) if (q && !aligned16(q)) return false;
    return true;
}

static int check_geom(long long N, int J, int H, int W) {
    if (N < 0 || J <= 0 || H <= 0 || W <= 0) return LH_EINVAL;
    if ((long long)H * W > (1 << 24)) return LH_EUNSUPPORTED;  // float32 index arithmetic of the reference is exact only below 2^24
    return LH_OK;
}

static void fill_ws(HeatmapParams& p, void* ws) {
    char* b = reinterpret_cast<char*>(ws);
    p.ticket = reinterpret_cast<unsigned int*>(b);
    // one slot per CTA; the grid never exceeds n_maps
    p.cta_loss = reinterpret_cast<double*>(b + 16);
    p.cta_epe = reinterpret_cast<double*>(b + 16 + 8 * (size_t)p.n_maps);
    p.cta_wrong = reinterpret_cast<unsigned int*>(b + 16 + 16 * (size_t)p.n_maps);
}

}  // namespace lh

using namespace lh;

extern "C" size_t lh_heatmap_ws_bytes(int64_t N, int J) {
    if (N < 0 || J < 0) return 0;
    return 16 + 24 * (size_t)N * (size_t)J;
}

extern "C" int lh_gen_targets(const float* joints, int joint_stride, int N, int J, int H, int W,
                              float stride_x, float stride_y, int style, const float* gauss_table,
                              int table_side, float* target, float* target_weight, lh_stream_t stream) {
    if (!joints || !gauss_table || !target) return LH_ENULL;
    int rc = check_geom(N, J, H, W);
    if (rc) return rc;
    if (joint_stride < 2 || table_side < 1 || (table_side & 1) == 0 || table_side * table_side > kMaxTable) return LH_EINVAL;
    if (style != LH_TARGET_HRNET && style != LH_TARGET_HOURGLASS) return LH_EINVAL;
    if (!(stride_x > 0.f) || !(stride_y > 0.f)) return LH_EINVAL;
    HeatmapParams p{};
    p.joints = joints; p.table = gauss_table; p.target_out = target; p.weight_out = target_weight;
    p.n_maps = (long long)N * J; p.J = J; p.H = H; p.W = W; p.HW = H * W;
    p.joint_stride = joint_stride; p.side = table_side; p.style = style;
    p.stride_x = stride_x; p.stride_y = stride_y;
    return launch<false, 1>(p, can_vec(p), (cudaStream_t)stream);
}

extern "C" int lh_jmse_loss(const float* pred, const float* target, const float* weight, int64_t N, int J,
                            int H, int W, double* loss_f64, float* loss_f32, float* grad, void* ws,
                            lh_stream_t stream) {
    if (!pred || !target || !ws) return LH_ENULL;
    int rc = check_geom(N, J, H, W);
    if (rc) return rc;
    HeatmapParams p{};
    p.pred = pred; p.target_in = target; p.weight_in = weight; p.use_weight = weight != nullptr;
    p.grad = grad; p.loss_f64 = loss_f64; p.loss_f32 = loss_f32;
    p.n_maps = (long long)N * J; p.J = J; p.H = H; p.W = W; p.HW = H * W;
    const double cnt = (double)N * J * (double)p.HW;
    p.grad_scale = (float)(1.0 / cnt); p.loss_scale = 0.5 / cnt;
    fill_ws(p, ws);
    return launch<true, 2>(p, can_vec(p), (cudaStream_t)stream);
}

extern "C" int lh_decode(const float* heat, int64_t N, int J, int H, int W, float out_scale, int subpixel,
                         const double* inv_affine, float* coords, float* maxvals, int32_t* argidx,
                         lh_stream_t stream) {
    if (!heat || !coords) return LH_ENULL;
    int rc = check_geom(N, J, H, W);
    if (rc) return rc;
    HeatmapParams p{};
    p.pred = heat; p.coords = coords; p.maxvals = maxvals; p.argidx = argidx; p.inv_affine = inv_affine;
    p.n_maps = (long long)N * J; p.J = J; p.H = H; p.W = W; p.HW = H * W;
    p.out_scale = out_scale; p.subpixel = subpixel;
    return launch<true, 0>(p, can_vec(p), (cudaStream_t)stream);
}

extern "C" int lh_heatmap_fused(const lh_heatmap_args* a, lh_stream_t stream) {
    if (!a || !a->pred || !a->ws) return LH_ENULL;
    if (!a->target_in && (!a->joints || !a->gauss_table)) return LH_ENULL;
    int rc = check_geom(a->N, a->J, a->H, a->W);
    if (rc) return rc;
    HeatmapParams p{};
    p.pred = a->pred; p.joints = a->joints; p.table = a->gauss_table; p.target_in = a->target_in;
    p.inv_affine = a->inv_affine; p.target_out = a->target_out; p.weight_out = a->target_weight_out;
    p.grad = a->grad; p.loss_f64 = a->loss_f64; p.loss_f32 = a->loss_f32; p.loss_meter = a->loss_meter;
    p.coords = a->coords; p.maxvals = a->maxvals; p.argidx = a->argidx;
    p.n_maps = (long long)a->N * a->J; p.J = a->J; p.H = a->H; p.W = a->W; p.HW = a->H * a->W;
    p.joint_stride = a->joint_stride; p.side = a->table_side; p.style = a->style;
    p.use_weight = a->use_target_weight; p.subpixel = a->subpixel;
    p.stride_x = a->stride_x; p.stride_y = a->stride_y; p.out_scale = a->out_scale;
    if (a->metric_counts || a->metric_sums) {
        // fused K5 needs the decode, the ground-truth joints and both accumulators
        if (!a->metric_counts || !a->metric_sums || !a->joints || !a->coords) return LH_ENULL;
        if (a->joint_stride < 2) return LH_EINVAL;
        if (a->pck_mode != LH_PCK_PROPORTION && a->pck_mode != LH_PCK_MM) return LH_EINVAL;
        p.metric_counts = a->metric_counts; p.metric_sums = a->metric_sums;
        p.pck_mode = a->pck_mode; p.pck_thr = a->pck_thr;
    }
    const double cnt = (double)a->N * a->J * (double)p.HW;
    p.grad_scale = (float)(1.0 / cnt); p.loss_scale = 0.5 / cnt;
    fill_ws(p, a->ws);
    if (a->target_in) {
        p.target_out = nullptr;
        return launch<true, 2>(p, can_vec(p), (cudaStream_t)stream);
    }
    if (a->joint_stride < 2 || a->table_side < 1 || (a->table_side & 1) == 0 ||
        a->table_side * a->table_side > kMaxTable) return LH_EINVAL;
    if (!(a->stride_x > 0.f) || !(a->stride_y > 0.f)) return LH_EINVAL;
    return launch<true, 1>(p, can_vec(p), (cudaStream_t)stream);
}
