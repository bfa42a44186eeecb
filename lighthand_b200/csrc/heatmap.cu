// K2 / K3 / K4 and their fusion: one streaming pass over [N,J,H,W] heat-maps.
//
//   gen_targets   (src/tools/dataset.py:165-212, src/utils/dataset_loader.py:22-53)
//   JointsMSELoss (src/utils/loss.py:306-325) forward + backward
//   get_max_preds (src/utils/loss.py:327-355) [+ quarter-pixel shift, inverse affine]
//
// One CTA of 256 threads owns one (n,j) map: H*W floats are streamed with 16-byte loads/stores,
// 4 independent vectors in flight per thread; the target value is recomputed from the joint
// (a <=19x19 table in shared memory) instead of being re-read; the loss partial and the argmax
// are reduced with warp shuffles / REDUX, then once across the 8 warps through shared memory.
// HBM-bound by construction: fused it moves H*W*4 B in (+ H*W*4 B per requested output).
#include "lh_common.cuh"

namespace lh {

constexpr int kThreads = 256;
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
    unsigned int* ticket;       // ws + 0
    double* partials;           // ws + 16, one per map
    long long n_maps;
    int J, H, W, HW;
    int joint_stride, side, style, use_weight, subpixel;
    float stride_x, stride_y, out_scale;
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

__device__ __forceinline__ Window joint_window(const HeatmapParams& p, long long m) {
    Window w;
    const float* jp = p.joints + m * p.joint_stride;
    const float jx = jp[0], jy = jp[1];
    w.valid = 1;
    if (p.style == LH_TARGET_HRNET) {
        // mu = int(joint / stride + 0.5) in float32  (src/tools/dataset.py:178-179)
        const float fx = __fadd_rn(__fdiv_rn(jx, p.stride_x), 0.5f);
        const float fy = __fadd_rn(__fdiv_rn(jy, p.stride_y), 0.5f);
        const int tmp = (p.side - 1) >> 1;
        const int mx = trunc_i(fx), my = trunc_i(fy);
        w.ulx = mx - tmp; w.uly = my - tmp;
        w.brx = mx + tmp + 1; w.bry = my + tmp + 1;
        if (!(fx == fx) || !(fy == fy)) w.valid = 0;  // int(nan) raises in the reference; no target here
        if (w.ulx >= p.W || w.uly >= p.H || w.brx < 0 || w.bry < 0) w.valid = 0;  // dataset.py:183-186
    } else {
        // GenerateHeatmap: p = joint/4; gate on x > 0 only; int() truncation (dataset_loader.py:38-41)
        const float px = __fdiv_rn(jx, p.stride_x), py = __fdiv_rn(jy, p.stride_y);
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

// TSRC: 0 = no target (decode only), 1 = generate from joints, 2 = read target_in.
template <bool PRED, int TSRC, bool VEC>
__global__ void __launch_bounds__(kThreads) heatmap_kernel(const HeatmapParams p) {
    __shared__ float s_tab[TSRC == 1 ? kMaxTable : 1];
    __shared__ double s_sum[kThreads / kWarp];
    __shared__ uint32_t s_key[kThreads / kWarp];
    __shared__ int s_idx[kThreads / kWarp];
    __shared__ int s_last;

    const int tid = threadIdx.x;
    const long long m = blockIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    Window win;
    win.valid = 0; win.ulx = win.uly = win.brx = win.bry = 0;
    if (TSRC == 1) {
        for (int i = tid; i < p.side * p.side; i += kThreads) s_tab[i] = p.table[i];
        win = joint_window(p, m);
        if (!win.valid) { win.ulx = win.uly = 0; win.brx = win.bry = 0; }
        __syncthreads();
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
    const bool do_loss = PRED && TSRC != 0;
    const bool do_dec = PRED && p.coords != nullptr;

    float acc0 = 0.f, acc1 = 0.f;
    uint32_t best_key = 0u;
    int best_idx = 0x7fffffff;
    const float gs = p.grad_scale * wgt;  // grad = (pw - tw) * w / count

    if (VEC) {
        constexpr int U = 4;
        const int nvec = p.HW >> 2;
        for (int b0 = 0; b0 < nvec; b0 += kThreads * U) {
            float4 pv[U], tv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = b0 + u * kThreads + tid;
                if (PRED && i < nvec) pv[u] = ldg_stream4(pred + 4 * (size_t)i);
                if (TSRC == 2 && i < nvec) tv[u] = ldg_stream4(tin + 4 * (size_t)i);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = b0 + u * kThreads + tid;
                if (i >= nvec) continue;
                const int e = 4 * i;
                float t[4] = {0.f, 0.f, 0.f, 0.f};
                if (TSRC == 1) {
                    const int y = e / p.W, x = e - y * p.W;
                    if (y >= win.uly && y < win.bry && x + 3 >= win.ulx && x < win.brx) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) t[k] = target_at(win, s_tab, p.side, y, x + k);
                    }
                    if (tout) stg_stream4(tout + e, make_float4(t[0], t[1], t[2], t[3]));
                } else if (TSRC == 2) {
                    t[0] = tv[u].x; t[1] = tv[u].y; t[2] = tv[u].z; t[3] = tv[u].w;
                }
                if (PRED) {
                    const float v[4] = {pv[u].x, pv[u].y, pv[u].z, pv[u].w};
                    if (do_loss) {
                        float d[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) d[k] = (v[k] - t[k]) * wgt;
                        acc0 = fmaf(d[0], d[0], acc0); acc1 = fmaf(d[1], d[1], acc1);
                        acc0 = fmaf(d[2], d[2], acc0); acc1 = fmaf(d[3], d[3], acc1);
                        if (grad) stg_cs4(grad + e, make_float4(d[0] * gs, d[1] * gs, d[2] * gs, d[3] * gs));
                    }
                    if (do_dec) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t key = order_key(v[k]);
                            if (key > best_key) { best_key = key; best_idx = e + k; }
                        }
                    }
                }
            }
        }
    } else {
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
    }

    if (!PRED) return;

    // ---- block reductions ------------------------------------------------------------------
    double wsum = do_loss ? warp_sum((double)acc0 + (double)acc1) : 0.0;
    uint32_t kmax = 0u;
    int imin = 0x7fffffff;
    if (do_dec) {
        kmax = __reduce_max_sync(0xffffffffu, best_key);
        imin = __reduce_min_sync(0xffffffffu, best_key == kmax ? best_idx : 0x7fffffff);
    }
    if (lane == 0) { s_sum[warp] = wsum; s_key[warp] = kmax; s_idx[warp] = imin; }
    __syncthreads();

    if (tid == 0) {
        if (do_loss) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < kThreads / kWarp; ++w) s += s_sum[w];
            p.partials[m] = s;
        }
        if (do_dec) {
            uint32_t bk = 0u; int bi = 0x7fffffff;
#pragma unroll
            for (int w = 0; w < kThreads / kWarp; ++w) {
                if (s_key[w] > bk || (s_key[w] == bk && s_idx[w] < bi)) { bk = s_key[w]; bi = s_idx[w]; }
            }
            // every element has key >= 1 (only -NaN-free minimum -inf maps to 0x007fffff), so bi is set
            // unless HW == 0; keep a defined value anyway.
            if (bi == 0x7fffffff) bi = 0;
            const float mv = pred[bi];
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
        }
    }

    // ---- deterministic final loss: the last CTA to finish adds the per-map partials in index order
    if (do_loss && (p.loss_f64 || p.loss_f32 || p.loss_meter)) {
        if (tid == 0) {
            __threadfence();
            const unsigned int t = atomicAdd(p.ticket, 1u);
            s_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            double s = 0.0;
            for (long long i = tid; i < p.n_maps; i += kThreads) s += __ldcg(p.partials + i);
            s = warp_sum(s);
            if (lane == 0) s_sum[warp] = s;
            __syncthreads();
            if (tid == 0) {
                double tot = 0.0;
#pragma unroll
                for (int w = 0; w < kThreads / kWarp; ++w) tot += s_sum[w];
                const double loss = tot * p.loss_scale;
                if (p.loss_f64) *p.loss_f64 = loss;
                if (p.loss_f32) *p.loss_f32 = (float)loss;
                if (p.loss_meter) {  // AverageMeter.update(loss.item(), batch_size): float32 loss, stream-ordered
                    const double nb = (double)(p.n_maps / p.J);
                    p.loss_meter[0] += (double)(float)loss * nb;
                    p.loss_meter[1] += nb;
                }
                *p.ticket = 0u;  // leave the workspace clean for the next launch
            }
        }
    }
}

template <bool PRED, int TSRC>
static int launch(const HeatmapParams& p, bool vec, cudaStream_t st) {
    if (p.n_maps == 0) return LH_OK;
    if (p.n_maps > 0x7fffffffLL) return LH_EUNSUPPORTED;
    dim3 grid((unsigned)p.n_maps), block(kThreads);
    if (vec) heatmap_kernel<PRED, TSRC, true><<<grid, block, 0, st>>>(p);
    else     heatmap_kernel<PRED, TSRC, false><<<grid, block, 0, st>>>(p);
    LH_CHECK_LAUNCH();
    return LH_OK;
}

static bool can_vec(const HeatmapParams& p) {
    if ((p.W & 3) != 0) return false;
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
    p.ticket = reinterpret_cast<unsigned int*>(ws);
    p.partials = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + 16);
}

}  // namespace lh

using namespace lh;

extern "C" size_t lh_heatmap_ws_bytes(int64_t N, int J) {
    if (N < 0 || J < 0) return 0;
    return 16 + sizeof(double) * (size_t)N * (size_t)J;
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
