// K2 / K3 / K4 (+ K5) and their fusion: one streaming pass over [N,J,H,W] heat-maps.
//
//   gen_targets   (src/tools/dataset.py:165-212, src/utils/dataset_loader.py:22-53)
//   JointsMSELoss (src/utils/loss.py:306-325) forward + backward
//   get_max_preds (src/utils/loss.py:327-355) [+ quarter-pixel shift, inverse affine]
//   PCK_2d_loss / EPE_train (src/utils/loss.py:50-67,116-148) on the decoded joints (optional, fused tail)
//
// One WARP owns one (n,j) map (heatmap_warp_kernel below): it streams the map as warp-rows of 32 float4 (512
// contiguous bytes per load instruction), two register buffers of 4 vectors per lane, and never meets a block
// barrier.  The target value is recomputed from the joint (a <=31x31 table in shared memory) instead of being
// re-read.  The argmax keeps the exact numpy order (first index on ties, NaN is the maximum): one NaN-propagating
// max of the 4 elements of a vector against the lane's running best, REDUX on the order key across the warp, and
// only the lanes that hold the maximum look at single elements.  Loss partials and the argmax are reduced with
// shuffles / REDUX inside the warp; the same warp finishes its map (coordinates, PCK / EPE record).  Every CTA
// leaves one slot of sums; the last CTA to finish (ticket) adds the slots in index order (deterministic loss) and,
// if asked, folds the PCK / EPE counts of the batch into the running metric accumulators -- so a whole step's
// post-backbone work is ONE launch.  HBM-bound by construction: it moves H*W*4 B in (+ H*W*4 B per requested
// output).  Shapes the vector kernel does not take (W % 4 != 0, unaligned pointers) go to heatmap_scalar_kernel.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "lh_common.cuh"

namespace lh {

constexpr int kThreads = 256;      // scalar kernel
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
    unsigned long long* map_key; // per map: (order key << 32 | ~index) of split maps (zero between launches)
    unsigned int* map_cnt;      // per map: pieces arrived
    long long n_full, units;    // whole-map units, all units (work list of heatmap_warp_kernel)
    int P, WR;                  // pieces per remaining map, warp-rows (32 float4) per map
    unsigned int vinv;          // ceil(2^20 / (W/4)): v / (W/4) == (v * vinv) >> 20
    long long n_maps;
    int J, H, W, HW;
    int joint_stride, side, style, use_weight, subpixel;
    int use_vis;                // upstream HRNet: target_weight = visibility (3rd joint column) * in-bounds, drawn when > 0.5
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

// ticket = atomicAdd(ticket, 1) with acquire-release semantics at GPU scope
__device__ __forceinline__ unsigned int ticket_acq_rel(unsigned int* ticket) {
    unsigned int old;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(ticket) : "memory");
    return old;
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

// bbox diagonal of sample n across the warp (PCK "proportion" mode): the reference's builtin max() / min() keep the
// first element unless a later one compares strictly greater / smaller, so a NaN survives only as the first joint
// -> NaN-ignoring reduce, then patch.  `jl` = this lane's ground-truth joint (lane < J <= 32), else read from memory.
__device__ __forceinline__ float bbox_diag_warp(const HeatmapParams& p, long long n, float2 jl, float2 jfirst, int lane) {
    const float nan = __int_as_float(0x7fc00000);
    float mxx = nan, mnx = nan, mxy = nan, mny = nan;
    if (p.J <= 32) {
        if (lane < p.J) { mxx = mnx = jl.x; mxy = mny = jl.y; }
    } else {
        const float* g = p.joints + n * p.J * p.joint_stride;
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
    return __fsqrt_rn(__fadd_rn(__fmul_rn(w, w), __fmul_rn(h, h)));
}

// ------------------------------------------------------------------------------------------------
// Tail of one map: final argmax -> coordinates (+ sub-pixel / inverse affine), outputs, fused metric
// record.  Lane 0 of the owning warp; m = n * J + j, `diag` = the sample's bbox diagonal (bbox_diag_warp), `gj` =
// the map's own joint.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void map_tail(const HeatmapParams& p, long long m, long long n, int j, const float* pred, int bi,
                                         float mv, bool do_metric, float diag, float2 gj, int lane,
                                         double& epe_acc, unsigned int& wrong_acc) {
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
        const double* A = p.inv_affine + n * 6;
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
                atomicAdd(p.loss_meter + 0, (double)(float)loss * nb);   // single writer per launch, stream-ordered: a
                atomicAdd(p.loss_meter + 1, nb);                         // RED instead of a dependent load round trip
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

// The map's tail as a separate function (its registers -- FP64 inverse affine, bbox reduction -- stay out of the
// streaming loop's allocation).  Whole warp; the joints were fetched at the head of the unit.
__device__ __noinline__ void finish_map(const HeatmapParams& p, int m, int jm, int bi, float mv, bool do_metric, float diag,
                                        float2 gj, int lane, double& epe_acc, unsigned int& wrong_acc) {
    if (bi == 0x7fffffff) bi = 0;    // HW == 0 only
    map_tail(p, m, (m - jm) / p.J, jm, p.pred + (size_t)m * p.HW, bi, mv, do_metric, diag, gj, lane, epe_acc, wrong_acc);
}

// ------------------------------------------------------------------------------------------------
// Vectorised persistent kernel, ONE WARP PER MAP.  TSRC: 0 = no target (decode only), 1 = generate from joints,
// 2 = read target_in.  FLAGS (compile time, so absent outputs cost nothing): F_GRAD write d loss / d pred, F_TOUT
// write the target, F_DEC argmax decode, F_WGT per-map weights.
//
// A warp streams a map as "warp-rows" of 32 float4 (512 contiguous bytes per load instruction), kU of them in flight
// per lane; loss partial, maximum and first index are reduced with shuffles / REDUX inside the warp and the same warp
// finishes the map (coordinates, PCK / EPE record).  No block barrier and no shared-memory exchange per map -- the
// CTA-per-map version of round 1 spent its time in exactly that chain (stream -> reduce -> barrier -> winner search
// -> barrier) and lost a row group whenever H was not a multiple of its rows-per-pass (56 x 56).  Any H x W with
// W % 4 == 0 streams at the same rate here.
//
// Work list (static, interleaved over the CTAs): with Wt resident warps, the first floor(n_maps / Wt) * Wt maps are
// whole-map units; the REMAINING maps are cut into P = floor(Wt / remaining) pieces of consecutive warp-rows so that
// the last round keeps every warp busy for 1/P of a map instead of leaving most of them idle (batch 32 x 21 joints
// on 3552 warps: 672 maps in 5 pieces each; batch 256: 3552 + 1824 whole maps, P = 1).  Pieces of one map meet through the workspace: a 64-bit
// atomicMax on (order key << 32 | ~index) -- maximum value first, then the smallest index = numpy's tie-break -- and a
// counter; the piece that arrives last finishes the map and leaves both words zeroed.  Loss partials never meet per
// map: every unit adds to its warp's float64 sum, warps are combined per CTA in warp order, CTAs by the last CTA in
// index order -> deterministic loss for a given grid.
// ------------------------------------------------------------------------------------------------
constexpr int F_GRAD = 1, F_TOUT = 2, F_DEC = 4, F_WGT = 8;
constexpr int kWarpsPerCta = 8;
constexpr int kWarpThreads = 32 * kWarpsPerCta;

#ifndef LH_HM_MINB
#define LH_HM_MINB 3
#endif
template <bool PRED, int TSRC, int FLAGS>
__global__ void __launch_bounds__(kWarpThreads, TSRC == 2 ? 3 : LH_HM_MINB) heatmap_warp_kernel(const __grid_constant__ HeatmapParams p) {
    constexpr bool GRAD = (FLAGS & F_GRAD) != 0, TOUT = (FLAGS & F_TOUT) != 0, DEC = PRED && (FLAGS & F_DEC) != 0,
                   WGT = (FLAGS & F_WGT) != 0, LOSS = PRED && TSRC != 0;
#ifndef LH_HM_KU
#define LH_HM_KU 4
#endif
    constexpr int kU = (TSRC == 2) ? LH_HM_KU / 2 : LH_HM_KU;   // float4 per buffer; two buffers (one in flight, one being consumed)
    __shared__ float s_tab[TSRC == 1 ? kMaxTable : 1];
    __shared__ double s_wl[kWarpsPerCta], s_we[kWarpsPerCta];
    __shared__ unsigned int s_ww[kWarpsPerCta];
    __shared__ double s_fa[kWarpsPerCta], s_fb[kWarpsPerCta];
    __shared__ unsigned int s_fc[kWarpsPerCta];
    __shared__ int s_last;

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    if (TSRC == 1) {
        for (int i = tid; i < p.side * p.side; i += kWarpThreads) s_tab[i] = p.table[i];
        __syncthreads();
    }
    const bool do_metric = DEC && p.metric_counts != nullptr;
    const int W = p.W, side = p.side, V = p.HW >> 2, vw = W >> 2;
    const float ninf = __int_as_float(0xff800000);
    const int Wt = (int)gridDim.x * kWarpsPerCta;                       // (the launcher keeps units below 2^31)
    const int wid = warp * (int)gridDim.x + (int)blockIdx.x;            // consecutive units land on different SMs
    const int n_full = (int)p.n_full, units = (int)p.units;

    double loss_acc = 0.0, epe_acc = 0.0;          // this warp's running sums (lane 0)
    unsigned int wrong_acc = 0u;

    for (int u = wid; u < units; u += Wt) {
        int m, k = 0, Pm = 1;
        if (u < n_full) {
            m = u;
        } else {
            const int r = u - n_full, q = r / p.P;
            m = n_full + q; k = r - q * p.P; Pm = p.P;
        }
        const int wr0 = (int)(((long long)k * p.WR) / Pm), wr1 = (int)(((long long)(k + 1) * p.WR) / Pm);

        const size_t base = (size_t)m * p.HW;
        const float* pred = PRED ? p.pred + base : nullptr;
        const float* tin = (TSRC == 2) ? p.target_in + base : nullptr;
        float* tout = TOUT ? p.target_out + base : nullptr;
        float* grad = GRAD ? p.grad + base : nullptr;
        const int vend = min(wr1 * 32, V);                      // this lane streams vectors v = wr * 32 + lane < vend

        // ---- software pipeline: two buffers of kU vectors; the first chunk's loads go out BEFORE anything that
        // depends on the joints (a warp owns little more than one map: every dependent global round trip at the head or
        // the tail of a unit would be paid in full) ---------------------------------------------------------------
        float4 pa[kU], pb[kU], ta[kU], tb[kU];
#define LH_LOAD(PV, TV, WRB)                                                        \
        _Pragma("unroll") for (int i = 0; i < kU; ++i) {                            \
            const int v_ = ((WRB) + i) * 32 + lane;                                 \
            if (v_ < vend) {                                                        \
                if (PRED) PV[i] = ldg_stream4(pred + 4 * v_);                       \
                if (TSRC == 2) TV[i] = ldg_stream4(tin + 4 * v_);                   \
            }                                                                       \
        }
        LH_LOAD(pa, ta, wr0)

        // ---- per-map setup: ONE coalesced load brings the sample's J ground-truth joints (lane = joint); the map's own
        // joint (window, PCK / EPE record) and the first joint (bbox quirk) come out of it by shuffle ----------------
        float2 jl = make_float2(0.f, 0.f), gj = jl, jfirst = jl;
        float vis = 1.0f;                                        // the map's joint visibility (use_vis only)
        const bool need_j = (TSRC == 1) || do_metric;
        const int jm = m % p.J;
        if (need_j) {
            const float* jp = p.joints + (long long)m * p.joint_stride;
            if (TSRC == 1 && p.use_vis) vis = jp[2];
            if (p.J <= 32) {
                if (lane < p.J) {
                    const float* gp = p.joints + ((long long)(m - jm) + lane) * p.joint_stride;
                    jl = make_float2(gp[0], gp[1]);
                }
                gj = make_float2(__shfl_sync(0xffffffffu, jl.x, jm), __shfl_sync(0xffffffffu, jl.y, jm));
                jfirst = make_float2(__shfl_sync(0xffffffffu, jl.x, 0), __shfl_sync(0xffffffffu, jl.y, 0));
            } else {
                const float* gp = p.joints + (long long)(m - jm) * p.joint_stride;
                gj = make_float2(jp[0], jp[1]);
                jfirst = make_float2(gp[0], gp[1]);
            }
        }
        // the sample's bbox diagonal (PCK normaliser) now, while the first chunk is in flight -- not in the unit's tail
        float diag = 0.f;
        if (do_metric && p.pck_mode == LH_PCK_PROPORTION) diag = bbox_diag_warp(p, (m - jm) / p.J, jl, jfirst, lane);
        int ulx = 0, uly = 0, wwin = 0, hwin = 0;
        float wgt = 1.0f, gs = p.grad_scale;
        if (TSRC == 1) {
            const Window win = joint_window(p, gj.x, gj.y);
            // reference: target_weight = in-bounds (computed, then discarded).  use_vis (upstream HRNet semantics the
            // north star names): target_weight = visibility * in-bounds, and the Gaussian is drawn only when it is > 0.5
            const float tw = win.valid ? vis : 0.0f;
            const bool valid = p.use_vis ? (tw > 0.5f) : (win.valid != 0);
            if (valid) { ulx = win.ulx; uly = win.uly; wwin = win.brx - win.ulx; hwin = win.bry - win.uly; }
            if (p.weight_out && lane == 0 && k == 0) p.weight_out[m] = tw;
            if (WGT && !p.weight_in) wgt = tw;
        }
        if (WGT) {
            if (p.weight_in) wgt = p.weight_in[m];
            gs = p.grad_scale * wgt;                             // grad = (pw - tw) * w / count
        }
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f, tmax = ninf;
        int tvec = -1;                                           // element offset of the vector holding tmax (-1: none seen)

#define LH_CONSUME(PV, TV, WRB)                                                                             \
        _Pragma("unroll") for (int i = 0; i < kU; ++i) {                                                    \
            const int v = ((WRB) + i) * 32 + lane;                                                          \
            if (v < vend) {                                                                                 \
                const int e = 4 * v;                                                                        \
                float t[4] = {0.f, 0.f, 0.f, 0.f};                                                          \
                if (TSRC == 1) {                                                                            \
                    const int y = (int)(((unsigned)v * p.vinv) >> 20);   /* v / vw (exact: launcher checks) */ \
                    const int x = 4 * (v - y * vw);                                                         \
                    const unsigned ry = (unsigned)(y - uly);                                                \
                    if (ry < (unsigned)hwin && x + 3 >= ulx && x < ulx + wwin) {                            \
                        const float* trow = s_tab + ry * side;                                              \
                        _Pragma("unroll") for (int c = 0; c < 4; ++c) {                                     \
                            const unsigned xx = (unsigned)(x + c - ulx);                                    \
                            if (xx < (unsigned)wwin) t[c] = trow[xx];                                       \
                        }                                                                                   \
                    }                                                                                       \
                    if (TOUT) stg_stream4(tout + e, make_float4(t[0], t[1], t[2], t[3]));                   \
                } else if (TSRC == 2) {                                                                     \
                    t[0] = TV[i].x; t[1] = TV[i].y; t[2] = TV[i].z; t[3] = TV[i].w;                         \
                }                                                                                           \
                if (PRED) {                                                                                 \
                    const float q[4] = {PV[i].x, PV[i].y, PV[i].z, PV[i].w};                                \
                    if (LOSS) {                                                                             \
                        float d[4];                                                                         \
                        _Pragma("unroll") for (int c = 0; c < 4; ++c) d[c] = WGT ? (q[c] - t[c]) * wgt : (q[c] - t[c]); \
                        acc0 = fmaf(d[0], d[0], acc0); acc1 = fmaf(d[1], d[1], acc1);                       \
                        acc2 = fmaf(d[2], d[2], acc2); acc3 = fmaf(d[3], d[3], acc3);                       \
                        if (GRAD) stg_cs4(grad + e, make_float4(d[0] * gs, d[1] * gs, d[2] * gs, d[3] * gs)); \
                    }                                                                                       \
                    if (DEC) {                                                                              \
                        /* a lane visits its vectors in increasing index order: replace only on "strictly greater", or */ \
                        /* when the first NaN shows up (NaN is numpy's maximum); never once tmax is NaN */     \
                        const float vmax = max_nan(max_nan(q[0], q[1]), max_nan(q[2], q[3]));               \
                        const bool upd = (tvec < 0) || (!(vmax <= tmax) && (tmax == tmax));                 \
                        tmax = upd ? vmax : tmax;                                                           \
                        tvec = upd ? e : tvec;                                                              \
                    }                                                                                       \
                }                                                                                           \
            }                                                                                               \
        }

        // ---- stream the unit's warp-rows: chunk c+1 in flight while chunk c is consumed -------------------------
        for (int wr = wr0; wr < wr1; wr += 2 * kU) {
            LH_LOAD(pb, tb, wr + kU)
            LH_CONSUME(pa, ta, wr)
            LH_LOAD(pa, ta, wr + 2 * kU)
            LH_CONSUME(pb, tb, wr + kU)
        }
#undef LH_LOAD
#undef LH_CONSUME
        if (!PRED) continue;

        // ---- warp reductions ------------------------------------------------------------------------------------
        if (LOSS) {
            const double ws = warp_sum((double)(acc0 + acc1) + (double)(acc2 + acc3));
            loss_acc += ws;                                      // every lane holds the sum; lane 0's copy is the one used
        }
        if (DEC) {
            const uint32_t key = tvec >= 0 ? order_key(tmax) : 0u;
            const uint32_t kmax = __reduce_max_sync(0xffffffffu, key);
            // lanes whose own maximum IS the unit's maximum look up its first position inside their vector
            int cand = 0x7fffffff;
            float cval = 0.f;
            if (tvec >= 0 && key == kmax) {
                const float4 q4 = *reinterpret_cast<const float4*>(pred + tvec);
                const float q[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
                for (int c = 3; c >= 0; --c) if (order_key(q[c]) == kmax) { cand = tvec + c; cval = q[c]; }
            }
            int bi = __reduce_min_sync(0xffffffffu, cand);
            const unsigned owner = __ballot_sync(0xffffffffu, cand == bi);
            float mv = __shfl_sync(0xffffffffu, cval, __ffs(owner) - 1);
            bool finish = true;
            if (Pm > 1) {
                // pieces of one map meet in the workspace; the last one to arrive finishes the map
                int last = 0;
                if (lane == 0) {
                    const unsigned long long k64 = ((unsigned long long)kmax << 32) | (unsigned long long)(0xffffffffu - (unsigned)bi);
                    atomicMax(p.map_key + m, k64);
                    __threadfence();
                    const unsigned int old = atomicAdd(p.map_cnt + m, 1u);
                    if (old == (unsigned)(Pm - 1)) {
                        __threadfence();
                        const unsigned long long kk = atomicExch(p.map_key + m, 0ull);   // leave the workspace zeroed
                        p.map_cnt[m] = 0u;
                        bi = (int)(0xffffffffu - (unsigned)(kk & 0xffffffffull));
                        mv = __ldcg(pred + bi);
                        last = 1;
                    }
                }
                last = __shfl_sync(0xffffffffu, last, 0);
                bi = __shfl_sync(0xffffffffu, bi, 0);
                mv = __shfl_sync(0xffffffffu, mv, 0);
                finish = last != 0;
            }
            if (finish && lane == 0) finish_map(p, m, jm, bi, mv, do_metric, diag, gj, lane, epe_acc, wrong_acc);
        }
    }

    if (!PRED) return;
    if ((LOSS && (p.loss_f64 || p.loss_f32 || p.loss_meter)) || do_metric) {
        if (lane == 0) { s_wl[warp] = loss_acc; s_we[warp] = epe_acc; s_ww[warp] = wrong_acc; }
        __syncthreads();
        if (tid == 0) {
            double cl = 0.0, ce = 0.0;
            unsigned int cw = 0u;
#pragma unroll
            for (int w = 0; w < kWarpsPerCta; ++w) { cl += s_wl[w]; ce += s_we[w]; cw += s_ww[w]; }
            p.cta_loss[blockIdx.x] = cl; p.cta_epe[blockIdx.x] = ce; p.cta_wrong[blockIdx.x] = cw;
            // release our slot / acquire everybody else's in ONE acq_rel atomic: __threadfence() is a sequentially
            // consistent fence (MEMBAR.SC) and the pair fence + atomic + fence sat on the kernel's critical path --
            // only the LAST CTA's tail is exposed, but it is exposed in full (fused pass)
            const unsigned int t = ticket_acq_rel(p.ticket);
            s_last = (t == gridDim.x - 1);
        }
        __syncthreads();      // orders the other threads' slot loads (ld.cg, L2) after thread 0's acquire
        if (s_last) final_reduce<LOSS, kWarpThreads>(p, do_metric, s_fa, s_fb, s_fc, tid);
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
    float tw = 0.0f;            // target weight of this map (TSRC == 1)
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
        tw = win.valid ? (p.use_vis ? p.joints[m * p.joint_stride + 2] : 1.0f) : 0.0f;
        if (p.use_vis && !(tw > 0.5f)) win.valid = 0;
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
        else if (TSRC == 1) wgt = tw;
    }
    if (TSRC == 1 && p.weight_out && tid == 0) p.weight_out[m] = tw;

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
            const long long n = m / p.J;
            float diag = 0.f;
            if (do_metric && p.pck_mode == LH_PCK_PROPORTION) diag = bbox_diag_warp(p, n, jl, jfirst, lane);
            map_tail(p, m, n, (int)(m - n * p.J), pred, bi, pred[bi], do_metric, diag, gj, lane, cta_epe, cta_wrong);
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
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kWarpThreads, 0) != cudaSuccess) return 0;
    if (occ < 1) occ = 1;
    *cache = sms * occ;
    return *cache;
}

constexpr int kCtaSlotsExtra = 1024;   // the workspace holds n_maps + kCtaSlotsExtra per-CTA slots (lh_heatmap_ws_bytes)
constexpr int kMaxPieces = 16;

template <bool PRED, int TSRC, int FLAGS>
static int launch_vec(const HeatmapParams& p0, cudaStream_t st) {
    static int cache = 0;   // one device model per process (B200); benign race: same value
    const int cap = persistent_grid(heatmap_warp_kernel<PRED, TSRC, FLAGS>, &cache);
    if (cap <= 0) return (int)cudaGetLastError();
    HeatmapParams p = p0;
    const int V = p.HW >> 2;
    p.WR = (V + 31) / 32;
    p.vinv = ((1u << 20) + (unsigned)(p.W >> 2) - 1u) / (unsigned)(p.W >> 2);
    // work list: whole maps first, the remainder cut into pieces so that the last round fills the machine
    long long grid = p.n_maps * (long long)kMaxPieces;        // warps that could be given work at all ...
    grid = (grid + kWarpsPerCta - 1) / kWarpsPerCta;          // ... as CTAs
    if (grid > cap) grid = cap;
    if (grid > p.n_maps + kCtaSlotsExtra) grid = p.n_maps + kCtaSlotsExtra;   // per-CTA slots in the workspace
    if (grid < 1) grid = 1;
    const long long Wt = grid * kWarpsPerCta;
    p.n_full = (p.n_maps / Wt) * Wt;
    const long long rem = p.n_maps - p.n_full;
    constexpr bool kDecode = PRED && (FLAGS & F_DEC) != 0;
    long long P = 1;
    if (rem > 0 && (!kDecode || p.map_key != nullptr)) {      // pieces of a decoded map meet in the workspace
        // one piece per idle warp of the last round, never more: a piece pays the unit's fixed latency (first loads,
        // joints, two workspace atomics) whatever its size -- measured: 16 pieces per remaining map (shortest tail on
        // paper, 9/16 of a map-time) cost 59 us instead of 25 us for the C2 eval pass, the per-piece latency chain
        // serialises on each warp; and in the last round the active warps have the whole HBM bandwidth to themselves
        P = Wt / rem;
        if (P > kMaxPieces) P = kMaxPieces;
        if (P > p.WR) P = p.WR;
        if (P < 1) P = 1;
    }
    p.P = (int)P;
    p.units = p.n_full + rem * P;
    if (p.units >= (1LL << 31)) return LH_EUNSUPPORTED;
    const long long need = (p.units + kWarpsPerCta - 1) / kWarpsPerCta;
    if (grid > need) grid = need;
    // NOTE: n_full / P were derived from Wt of the capped grid; a smaller final grid only changes who takes which unit
    heatmap_warp_kernel<PRED, TSRC, FLAGS><<<(unsigned)grid, kWarpThreads, 0, st>>>(p);
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
    // float4 rows; (v * vinv) >> 20 == v / (W/4) needs v * (W/4) < 2^20 for every vector index v < H*W/4
    if ((p.W & 3) != 0 || (long long)(p.HW >> 2) * (p.W >> 2) >= (1 << 20)) return false;
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
    // one slot per CTA: the grid never exceeds n_maps + kCtaSlotsExtra
    const size_t slots = (size_t)p.n_maps + kCtaSlotsExtra;
    p.cta_loss = reinterpret_cast<double*>(b + 16);
    p.cta_epe = reinterpret_cast<double*>(b + 16 + 8 * slots);
    p.cta_wrong = reinterpret_cast<unsigned int*>(b + 16 + 16 * slots);
    p.map_key = reinterpret_cast<unsigned long long*>(b + 16 + 24 * slots);
    p.map_cnt = reinterpret_cast<unsigned int*>(b + 16 + 24 * slots + 8 * (size_t)p.n_maps);
}

}  // namespace lh

using namespace lh;

extern "C" size_t lh_heatmap_ws_bytes(int64_t N, int J) {
    if (N < 0 || J < 0) return 0;
    const size_t maps = (size_t)N * (size_t)J;
    // ticket | per-CTA slots (loss f64, epe f64, wrong u32) | per-map combine words of split maps (key u64, count u32)
    return 16 + 24 * (maps + kCtaSlotsExtra) + 16 * maps;
}

extern "C" int lh_gen_targets(const float* joints, int joint_stride, int N, int J, int H, int W,
                              float stride_x, float stride_y, int style, int use_vis, const float* gauss_table,
                              int table_side, float* target, float* target_weight, lh_stream_t stream) {
    if (!joints || !gauss_table || !target) return LH_ENULL;
    int rc = check_geom(N, J, H, W);
    if (rc) return rc;
    if (joint_stride < 2 || table_side < 1 || (table_side & 1) == 0 || table_side * table_side > kMaxTable) return LH_EINVAL;
    if (use_vis && joint_stride < 3) return LH_EINVAL;           // visibility = third joint column
    if (style != LH_TARGET_HRNET && style != LH_TARGET_HOURGLASS) return LH_EINVAL;
    if (!(stride_x > 0.f) || !(stride_y > 0.f)) return LH_EINVAL;
    HeatmapParams p{};
    p.joints = joints; p.table = gauss_table; p.target_out = target; p.weight_out = target_weight;
    p.n_maps = (long long)N * J; p.J = J; p.H = H; p.W = W; p.HW = H * W;
    p.joint_stride = joint_stride; p.side = table_side; p.style = style; p.use_vis = use_vis != 0;
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
    p.use_weight = a->use_target_weight; p.subpixel = a->subpixel; p.use_vis = a->use_vis != 0;
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
    if (a->use_vis && a->joint_stride < 3) return LH_EINVAL;
    if (!(a->stride_x > 0.f) || !(a->stride_y > 0.f)) return LH_EINVAL;
    return launch<true, 1>(p, can_vec(p), (cudaStream_t)stream);
}
