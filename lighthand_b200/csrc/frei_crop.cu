// FreiHAND image path: the multi-stage crop of src/utils/image_ops.py:97-131 and the
// tail of HandMeshTSVDataset.__getitem__ (src/datasets/frei_dataloader.py:134-147, 276-283, 338), batched.
//
//   stage R (rotated samples only)  myimrotate  image_ops.py:25-44   cv2.warpAffine of the zero-padded window about
//                                   its centre on cv2's 1/32-px fixed-point grid; only the part that survives
//                                   new_img[pad:-pad, pad:-pad] is computed, into a float32 scratch canvas.  The
//                                   values are sums of (uint8 x 10-bit dyadic weights): exact in float32, i.e.
//                                   bit-identical to cv2's float64 result.
//   stage S                         myimresize  image_ops.py:46-56   cv2.resize(INTER_LINEAR) of the (rotated)
//                                   canvas to res x res: half-pixel coordinates and floor in double like OpenCV,
//                                   blend in float32; then rgb_processing's pixel noise + clip + /255, CHW.
//   stage T                         transforms.Resize + Normalize + channels (2,1,0): lh_resize_normalize_f32.
//
// The integer window bookkeeping of crop() (transform(..., invert=1) through torch.inverse, int truncation, the
// rotation padding) is host logic: it arrives per sample in lh_frei_desc (lighthand_b200/utils/image_ops.py).
// The un-rotated window is never materialised: stage S reads the uint8 image through the window's zero padding.
#include "lh_common.cuh"

namespace lh {

constexpr int kFreiThreads = 256;

// pixel (cx, cy) of the zero float64 canvas crop() fills from the image (image_ops.py:107-122)
__device__ __forceinline__ void canvas_px(const uint8_t* img, int pitch, int sh, int sw, const lh_frei_desc& d,
                                          int cx, int cy, float (&v)[3]) {
    v[0] = v[1] = v[2] = 0.f;
    if (cx < 0 || cy < 0 || cx >= d.cw || cy >= d.ch) return;
    const int x = d.ulx + cx, y = d.uly + cy;
    if (x < 0 || y < 0 || x >= sw || y >= sh) return;
    const uint8_t* q = img + (long long)y * pitch + 3 * x;
    v[0] = (float)__ldg(q); v[1] = (float)__ldg(q + 1); v[2] = (float)__ldg(q + 2);
}

// three channels of canvas pixel (cx, cy) known to lie inside both the canvas and the image
__device__ __forceinline__ void image_px(const uint8_t* org, int pitch, int cx, int cy, float (&v)[3]) {
    const uint8_t* q = org + (long long)cy * pitch + 3 * cx;
    v[0] = (float)__ldg(q); v[1] = (float)__ldg(q + 1); v[2] = (float)__ldg(q + 2);
}

// byte k of w as a float: PRMT puts it under the exponent of 2^23, one FADD takes the 2^23 off again -- two full-rate
// instructions instead of a shift / mask and a quarter-rate I2F, and exact
__device__ __forceinline__ float byte_f(uint32_t w, uint32_t k) {
    return __uint_as_float(__byte_perm(w, 0x4b000000u, 0x7650u + k)) - 8388608.0f;
}

// pixels (cx, cy) and (cx + 1, cy), both inside the image, rows 4-byte aligned: the six bytes through three aligned 32-bit
// loads + two funnel shifts instead of six byte loads.  `row` = the image row, `woff` = (3 * cx) & ~3, `sh` = (3 * cx) << 3.
__device__ __forceinline__ void row_px2(const uint8_t* row, unsigned woff, unsigned sh, float (&v0)[3], float (&v1)[3]) {
    const uint32_t* q = reinterpret_cast<const uint32_t*>(row + woff);
    const uint32_t a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    const uint32_t lo = __funnelshift_r(a, b, sh), hi = __funnelshift_r(b, c, sh);   // R0 G0 B0 R1 | G1 B1 - -
    v0[0] = byte_f(lo, 0); v0[1] = byte_f(lo, 1); v0[2] = byte_f(lo, 2);
    v1[0] = byte_f(lo, 3); v1[1] = byte_f(hi, 0); v1[2] = byte_f(hi, 1);
}
__device__ __forceinline__ void image_px2(const uint8_t* org, int pitch, int cx, int cy, float (&v0)[3], float (&v1)[3]) {
    const unsigned off = (unsigned)(3 * cx);
    row_px2(org + (long long)cy * pitch, off & ~3u, off << 3, v0, v1);
}

constexpr int kFreiTileW = 32, kFreiTileH = kFreiThreads / 32;

// CTA = a 32 x 32 tile of the rotated canvas of one rotated sample (blockIdx.z indexes rot_index), a thread renders 4 pixels
// of one column: the FP64 row and column terms of cv2's coordinate grid are computed once per tile (64 threads), not once
// per pixel, and the taps of the four pixels are in flight together.  (Round 2: 32 x 8 tiles spent their time in that
// prologue -- descriptor load -> FP64 terms -> barrier -> taps -> store is one latency chain per CTA, 25 k CTAs per batch.)
constexpr int kRotTileH = 32;

__global__ void __launch_bounds__(kFreiThreads) frei_rotate_kernel(const uint8_t* __restrict__ src, long long img_stride,
                                                                   int sh, int sw, int pitch,
                                                                   const lh_frei_desc* __restrict__ desc,
                                                                   const int32_t* __restrict__ rot_index,
                                                                   float* __restrict__ scratch) {
    __shared__ int s_ax[kFreiTileW], s_bx[kFreiTileW], s_x0[kRotTileH], s_y0[kRotTileH];
    const int n = rot_index ? rot_index[blockIdx.z] : blockIdx.z;
    const lh_frei_desc& d = desc[n];
    if (!d.rotated) return;
    const int pad = d.pad, cw = d.cw, ch = d.ch, iw = cw - 2 * pad, ih = ch - 2 * pad;
    const int tx0 = blockIdx.x * kFreiTileW, ty0 = blockIdx.y * kRotTileH;
    if (tx0 >= iw || ty0 >= ih) return;                                  // uniform per CTA
    const int tid = threadIdx.x;
    if (tid < kFreiTileW) {
        s_ax[tid] = col_term(d.iM[0], tx0 + tid + pad);                  // position in the padded canvas warpAffine renders
        s_bx[tid] = col_term(d.iM[3], tx0 + tid + pad);
    } else if (tid < kFreiTileW + kRotTileH) {
        const int r = tid - kFreiTileW;
        s_x0[r] = row_term(d.iM[1], d.iM[2], ty0 + r + pad);
        s_y0[r] = row_term(d.iM[4], d.iM[5], ty0 + r + pad);
    }
    __syncthreads();
    const int lx = tid & 31, wy = tid >> 5;
    const int x = tx0 + lx;
    if (x >= iw) return;
    const int ax = s_ax[lx], bx = s_bx[lx];
    const int ulx = d.ulx, uly = d.uly;
    const uint8_t* img = src + (long long)n * img_stride;
    const bool words = ((reinterpret_cast<uintptr_t>(img) | (uintptr_t)pitch) & 3u) == 0;      // uniform
    float* out = scratch + d.scratch_off + ((size_t)(ty0 + wy) * iw + x) * 3;
    constexpr int kStep = kFreiThreads / 32;
#pragma unroll
    for (int k = 0; k < kRotTileH / kStep; ++k, out += (size_t)kStep * iw * 3) {
        const int ly = wy + k * kStep;
        if (ty0 + ly >= ih) break;
        const int X = (s_x0[ly] + ax) >> 5, Y = (s_y0[ly] + bx) >> 5;
        const int ix = clamp_i(X >> 5, -32768, 32767), iy = clamp_i(Y >> 5, -32768, 32767);   // cv2 stores shorts
        const float fx = (float)(X & 31) * 0.03125f, fy = (float)(Y & 31) * 0.03125f;
        const float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx, w10 = fy * (1.f - fx), w11 = fy * fx;
        float p00[3], p01[3], p10[3], p11[3];
        // all four taps inside the canvas and inside the image: no per-tap tests
        const int gx = ulx + ix, gy = uly + iy;
        if (ix >= 0 && iy >= 0 && ix + 1 < cw && iy + 1 < ch && gx >= 0 && gy >= 0 && gx + 1 < sw && gy + 1 < sh) {
            if (words && 3 * gx + 12 <= pitch) {          // the third word stays inside the row
                image_px2(img, pitch, gx, gy, p00, p01); image_px2(img, pitch, gx, gy + 1, p10, p11);
            } else {
                image_px(img, pitch, gx, gy, p00); image_px(img, pitch, gx + 1, gy, p01);
                image_px(img, pitch, gx, gy + 1, p10); image_px(img, pitch, gx + 1, gy + 1, p11);
            }
        } else {
            canvas_px(img, pitch, sh, sw, d, ix, iy, p00);
            canvas_px(img, pitch, sh, sw, d, ix + 1, iy, p01);
            canvas_px(img, pitch, sh, sw, d, ix, iy + 1, p10);
            canvas_px(img, pitch, sh, sw, d, ix + 1, iy + 1, p11);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)    // every product and partial sum is exact in float32 (<= 18 bits)
            out[c] = p00[c] * w00 + p01[c] * w01 + p10[c] * w10 + p11[c] * w11;
    }
}

// cv2.resize's source index / weight for destination index d (resize.cpp, INTER_LINEAR, CV_64F): double
// arithmetic without contraction, edge clamp
__device__ __forceinline__ void cv2_linear_axis(int d, int ssize, double scale, int& s0, int& s1, float& w0, float& w1) {
    double f = __dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5);
    int s = (int)floor(f);
    f = __dsub_rn(f, (double)s);
    if (s < 0) { f = 0.0; s = 0; }
    if (s >= ssize - 1) { f = 0.0; s = ssize - 1; }
    s0 = s; s1 = min(s + 1, ssize - 1);
    w0 = (float)__dsub_rn(1.0, f); w1 = (float)f;
}

// CTA = a 32 x 8 tile of the res x res output of one sample; cv2.resize's double-precision source indices and
// weights are computed once per tile column / row (40 threads)
__global__ void __launch_bounds__(kFreiThreads) frei_resize_kernel(const uint8_t* __restrict__ src, long long img_stride,
                                                                   int sh, int sw, int pitch,
                                                                   const lh_frei_desc* __restrict__ desc,
                                                                   const float* __restrict__ scratch, int res,
                                                                   float* __restrict__ dst) {
    __shared__ int s_x0[kFreiTileW], s_x1[kFreiTileW], s_y0[kFreiTileH], s_y1[kFreiTileH];
    __shared__ float s_a0[kFreiTileW], s_a1[kFreiTileW], s_b0[kFreiTileH], s_b1[kFreiTileH];
    const int n = blockIdx.z;
    const lh_frei_desc& d = desc[n];
    const int rotated = d.rotated;
    const int pad = rotated ? d.pad : 0;
    const int iw = d.cw - 2 * pad, ih = d.ch - 2 * pad;
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * kFreiTileW, ty0 = blockIdx.y * kFreiTileH;
    if (iw > 0 && ih > 0) {
        if (tid < kFreiTileW) cv2_linear_axis(tx0 + tid, iw, d.scale_x, s_x0[tid], s_x1[tid], s_a0[tid], s_a1[tid]);
        else if (tid < kFreiTileW + kFreiTileH)
            cv2_linear_axis(ty0 + tid - kFreiTileW, ih, d.scale_y, s_y0[tid - kFreiTileW], s_y1[tid - kFreiTileW],
                            s_b0[tid - kFreiTileW], s_b1[tid - kFreiTileW]);
    }
    __syncthreads();
    const int lx = tid & 31, ly = tid >> 5;
    const int x = tx0 + lx, y = ty0 + ly;
    if (x >= res || y >= res) return;
    float* o = dst + (size_t)n * 3 * res * res + (size_t)y * res + x;
    if (iw <= 0 || ih <= 0) { o[0] = 0.f; o[(size_t)res * res] = 0.f; o[2 * (size_t)res * res] = 0.f; return; }
    const int x0 = s_x0[lx], x1 = s_x1[lx], y0 = s_y0[ly], y1 = s_y1[ly];
    const float a0 = s_a0[lx], a1 = s_a1[lx], b0 = s_b0[ly], b1 = s_b1[ly];
    float p00[3], p01[3], p10[3], p11[3];
    if (rotated) {
        const float* r0 = scratch + d.scratch_off + (size_t)y0 * iw * 3;
        const float* r1 = scratch + d.scratch_off + (size_t)y1 * iw * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            p00[c] = __ldg(r0 + 3 * x0 + c); p01[c] = __ldg(r0 + 3 * x1 + c);
            p10[c] = __ldg(r1 + 3 * x0 + c); p11[c] = __ldg(r1 + 3 * x1 + c);
        }
    } else {
        const uint8_t* img = src + (long long)n * img_stride;
        const int gx0 = d.ulx + x0, gx1 = d.ulx + x1, gy0 = d.uly + y0, gy1 = d.uly + y1;   // x0 <= x1 < cw, y0 <= y1 < ch
        if (gx0 >= 0 && gy0 >= 0 && gx1 < sw && gy1 < sh) {
            image_px(img, pitch, gx0, gy0, p00); image_px(img, pitch, gx1, gy0, p01);
            image_px(img, pitch, gx0, gy1, p10); image_px(img, pitch, gx1, gy1, p11);
        } else {
            canvas_px(img, pitch, sh, sw, d, x0, y0, p00);
            canvas_px(img, pitch, sh, sw, d, x1, y0, p01);
            canvas_px(img, pitch, sh, sw, d, x0, y1, p10);
            canvas_px(img, pitch, sh, sw, d, x1, y1, p11);
        }
    }
    const float pn[3] = {d.pn[0], d.pn[1], d.pn[2]};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float top = fmaf(p01[c], a1, p00[c] * a0), bot = fmaf(p11[c], a1, p10[c] * a0);   // rows first ...
        float v = fmaf(bot, b1, top * b0);                                                       // ... then columns
        v = fminf(255.0f, fmaxf(0.0f, v * pn[c]));            // pixel noise + clip (frei_dataloader.py:142-144)
        o[(size_t)c * res * res] = v * (1.0f / 255.0f);       // CHW float32 / 255 (:146), <= 1 ulp from the division
    }
}

// transforms.Resize((oh, ow)) on a float CHW tensor (bilinear, align_corners = False, no antialias, ATen's
// float32 operation order) + Normalize (one fused multiply-add, <= 2 ulp from (x - mean) / std) + optional channel
// reversal (frei_dataloader.py:279-283, 338)
__global__ void __launch_bounds__(kFreiThreads) resize_f32_kernel(const float* __restrict__ src, int sh, int sw,
                                                                  float* __restrict__ dst, int oh, int ow, float sy, float sx,
                                                                  float a0, float a1, float a2, float b0, float b1,
                                                                  float b2, int chan_order, int vec) {
    // a thread renders 4 consecutive output columns of one row, all 3 channels: the row terms are computed once
    // and every channel leaves as one 16-byte streaming store
    const int n = blockIdx.z;
    const int ox = (blockIdx.x * 64 + (threadIdx.x & 63)) * 4, oy = blockIdx.y * (kFreiThreads / 64) + (threadIdx.x >> 6);
    if (ox >= ow || oy >= oh) return;
    float fy = fmaf(sy, __fadd_rn((float)oy, 0.5f), -0.5f);      // ATen: one fused multiply-add; sy = in / out (float32)
    fy = fy < 0.f ? 0.f : fy;
    const int y0 = min((int)fy, sh - 1), y1 = min(y0 + 1, sh - 1);
    const float ly1 = fminf(fmaxf(__fsub_rn(fy, (float)y0), 0.f), 1.f), ly0 = __fsub_rn(1.0f, ly1);
    int x0[4], x1[4];
    float lx0[4], lx1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float fx = fmaf(sx, __fadd_rn((float)(ox + k), 0.5f), -0.5f);
        fx = fx < 0.f ? 0.f : fx;
        x0[k] = min((int)fx, sw - 1); x1[k] = min(x0[k] + 1, sw - 1);
        lx1[k] = fminf(fmaxf(__fsub_rn(fx, (float)x0[k]), 0.f), 1.f); lx0[k] = __fsub_rn(1.0f, lx1[k]);
    }
    const float na[3] = {a0, a1, a2}, nb[3] = {b0, b1, b2};      // Normalize as v * a + b, a = 1/std, b = -mean/std
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float* r0 = src + (((size_t)n * 3 + c) * sh + y0) * sw;
        const float* r1 = src + (((size_t)n * 3 + c) * sh + y1) * sw;
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float top = __fadd_rn(__fmul_rn(lx0[k], __ldg(r0 + x0[k])), __fmul_rn(lx1[k], __ldg(r0 + x1[k])));
            const float bot = __fadd_rn(__fmul_rn(lx0[k], __ldg(r1 + x0[k])), __fmul_rn(lx1[k], __ldg(r1 + x1[k])));
            v[k] = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
            v[k] = fmaf(v[k], na[c], nb[c]);
        }
        const int co = (chan_order == LH_CHAN_BGR) ? 2 - c : c;
        float* o = dst + (((size_t)n * 3 + co) * oh + oy) * ow + ox;
        if (vec) {
            stg_cs4(o, make_float4(v[0], v[1], v[2], v[3]));
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (ox + k < ow) __stcs(o + k, v[k]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Stages S + T in ONE kernel (round 2): the res x res float image of rgb_processing never goes to HBM.
// CTA = two vertically adjacent 32 x 32 tiles of the final oh x ow output of one sample (descriptor and column tables
// are set up once, each tile has its own rectangle buffer so only one barrier separates its two stages).  Per tile it works out which rectangle of the res x res
// intermediate the tile's torch-bilinear taps touch (<= 32 x 32 when res / out <= 0.9375: 224 -> 256 needs 30 x 30), renders
// exactly that rectangle with stage S's arithmetic (cv2.resize in double coordinates, float32 blend, noise, clip, /255)
// into shared memory, and resamples it with stage T's arithmetic (ATen's float32 operation order, one FMA normalises,
// optional channel reversal).  Same floats as the two-kernel path, bit for bit -- the intermediate values are the same
// float32 numbers, they just live in shared memory -- for 15 % more stage-S pixels (tile halos) and 308 MB per
// 256-image batch less HBM traffic (154 MB written + 154 MB read back) and one launch less.
// ------------------------------------------------------------------------------------------------
constexpr int kFuseTile = 32;

// shared-memory load at a 32-bit shared address + compile-time byte offset (stage T reads 12 taps per output row off four
// row/column sums; spelled out so that the adds are not redone per tap)
template <int OFF>
__device__ __forceinline__ float lds_off(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF));
    return v;
}
// the value must live in a register from here on (keeps ptxas from re-deriving it from special registers and kernel
// parameters at every use, which is what it does under the register cap)
__device__ __forceinline__ void keep_reg(uint32_t& v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void keep_reg(int& v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void keep_reg(float*& v) { asm volatile("" : "+l"(v)); }
__device__ __forceinline__ void keep_reg(const uint8_t*& v) { asm volatile("" : "+l"(v)); }
constexpr int kFuseTilesPerCta = 2;      // vertically adjacent tiles per CTA: the column tables and the descriptor are set up once

#ifndef LH_FREI_MINB
#define LH_FREI_MINB 6
#endif
#ifndef LH_FREI_BUFS
#define LH_FREI_BUFS 2
#endif
// one 16-byte record per rectangle row (stage S) and per output row (stage T): a single LDS.128 per row and stage instead
// of four table reads with four address computations -- the kernel is instruction-bound (profiles/r02_frei_ncu_summary.txt)
struct __align__(16) FuseSRow { int y0, y1; float b0, b1; };
struct __align__(16) FuseTRow { int off0, off1; float l1, pad; };   // off = (tap row - rectangle top) * pitch, in bytes

__global__ void __launch_bounds__(kFreiThreads, LH_FREI_MINB) frei_fused_kernel(const uint8_t* __restrict__ src, long long img_stride,
                                                                  int sh, int sw, int pitch,
                                                                  const lh_frei_desc* __restrict__ desc,
                                                                  const float* __restrict__ scratch, int res,
                                                                  float* __restrict__ dst, int oh, int ow, float sy, float sx,
                                                                  float na0, float na1, float na2, float nb0, float nb1,
                                                                  float nb2, int chan_order) {
    constexpr int kRows = kFuseTile * kFuseTilesPerCta;
    constexpr int kBufs = LH_FREI_BUFS;                                      // 1: a barrier between tiles; 2: one rectangle buffer per tile
    constexpr int kRowStep = kFreiThreads / 32, kPasses = kFuseTile / kRowStep;
    constexpr int kPitchS = kFuseTile + 1, kPlane = kFuseTile * kPitchS;     // floats per rectangle row / channel plane
    __shared__ float s_im[kBufs][3 * kPlane];                                // the rectangle of the intermediate image
    __shared__ FuseSRow s_srow[kRows];
    __shared__ FuseTRow s_trow[kRows];
    __shared__ int s_ox0[kFuseTile], s_ox1[kFuseTile], s_oy0[kRows], s_oy1[kRows];             // stage T taps (absolute)
    __shared__ float s_olx[kFuseTile], s_oly[kRows];
    __shared__ int s_x0[kFuseTile], s_x1[kFuseTile];                                           // stage S taps per rectangle column
    __shared__ float s_a0[kFuseTile], s_a1[kFuseTile];

    const int n = blockIdx.z;
    const lh_frei_desc& d = desc[n];
    const int rotated = d.rotated;
    const int pad = rotated ? d.pad : 0;
    const int iw = d.cw - 2 * pad, ih = d.ch - 2 * pad;
    const int tid = threadIdx.x;
    const int tx0 = blockIdx.x * kFuseTile, ty0 = blockIdx.y * kRows;
    const int ncol = min(kFuseTile, ow - tx0), nrows = min(kRows, oh - ty0);

    // ---- stage T tables: torch's source index / lambda for the CTA's output columns and rows ----------------------
    if (tid < kFuseTile + kRows) {
        const bool is_row = tid >= kFuseTile;
        const int t = is_row ? tid - kFuseTile : tid;
        const int o = (is_row ? ty0 : tx0) + min(t, (is_row ? nrows : ncol) - 1);
        float f = fmaf(is_row ? sy : sx, __fadd_rn((float)o, 0.5f), -0.5f);   // ATen: one fused multiply-add
        f = f < 0.f ? 0.f : f;
        const int i0 = min((int)f, res - 1), i1 = min(i0 + 1, res - 1);
        const float l1 = fminf(fmaxf(__fsub_rn(f, (float)i0), 0.f), 1.f);
        if (is_row) { s_oy0[t] = i0; s_oy1[t] = i1; s_oly[t] = l1; }
        else        { s_ox0[t] = i0; s_ox1[t] = i1; s_olx[t] = l1; }
    }
    __syncthreads();
    // the rectangle of the intermediate image the taps touch: columns once, rows per tile (rw, rh <= 32: the launcher checks the scale)
    const int rx_lo = s_ox0[0], rw = s_ox1[ncol - 1] - rx_lo + 1;

    // ---- stage S tables for the rectangles' columns / rows, stage T row records ------------------------------------
    if (tid < kFuseTile) {
        if (tid < rw && iw > 0 && ih > 0)
            cv2_linear_axis(rx_lo + tid, iw, d.scale_x, s_x0[tid], s_x1[tid], s_a0[tid], s_a1[tid]);
    } else if (tid < kFuseTile + 2 * kRows) {
        const bool is_t = tid >= kFuseTile + kRows;
        const int t = tid - kFuseTile - (is_t ? kRows : 0), tile = t / kFuseTile, r = t - tile * kFuseTile;
        const int row0 = tile * kFuseTile, nr = min(kFuseTile, nrows - row0);
        if (nr > 0) {
            const int ry_lo = s_oy0[row0];
            if (is_t) {
                if (r < nr) {
                    FuseTRow tr;
                    tr.off0 = (s_oy0[t] - ry_lo) * kPitchS * 4; tr.off1 = (s_oy1[t] - ry_lo) * kPitchS * 4; tr.l1 = s_oly[t]; tr.pad = 0.f;
                    s_trow[t] = tr;
                }
            } else if (iw > 0 && ih > 0) {
                const int rh = s_oy1[row0 + nr - 1] - ry_lo + 1;
                if (r < rh) {
                    FuseSRow sr;
                    cv2_linear_axis(ry_lo + r, ih, d.scale_y, sr.y0, sr.y1, sr.b0, sr.b1);
                    s_srow[t] = sr;
                }
            }
        }
    }
    __syncthreads();

    const uint8_t* img = src + (long long)n * img_stride;
    const bool words = ((reinterpret_cast<uintptr_t>(img) | (uintptr_t)pitch) & 3u) == 0;      // uniform
    const float pn[3] = {d.pn[0], d.pn[1], d.pn[2]};
    const int ulx = d.ulx, uly = d.uly;
    const int lx = tid & 31, wy = tid >> 5;
    // this lane's rectangle column (stage S) and output column (stage T): the same for every tile of the CTA
    const bool s_col = lx < rw && iw > 0 && ih > 0;
    const int x0 = s_col ? s_x0[lx] : 0, x1 = s_col ? s_x1[lx] : 0;
    const float a0 = s_col ? s_a0[lx] : 0.f, a1 = s_col ? s_a1[lx] : 0.f;
    const int gx0 = ulx + x0, gx1 = ulx + x1;
    const bool col_in = gx0 >= 0 && gx1 < sw;
    const bool pair_ok = words && col_in && gx1 == gx0 + 1 && 3 * gx0 + 12 <= pitch;
    const unsigned boff = (unsigned)(3 * gx0);
    unsigned bsh = boff << 3;
    const uint8_t* imgw = img + (pair_ok ? (boff & ~3u) : 0u);                 // this column's first word of image row 0
    keep_reg(imgw); keep_reg(bsh);
    const float* rot0 = rotated ? scratch + d.scratch_off + 3 * x0 : nullptr;  // rotated canvas: this column's two taps
    const float* rot1 = rotated ? scratch + d.scratch_off + 3 * x1 : nullptr;
    const bool t_col = lx < ncol;
    const int ox0 = t_col ? s_ox0[lx] - rx_lo : 0, ox1 = t_col ? s_ox1[lx] - rx_lo : 0;
    const float lx1 = t_col ? s_olx[lx] : 0.f, lx0 = __fsub_rn(1.0f, lx1);
    // 32-bit indexing inside the sample's three output planes (the launcher checks 3 * oh * ow < 2^31)
    const int plane = oh * ow;
    const int c0 = (chan_order == LH_CHAN_BGR) ? 2 * plane : 0, c2 = 2 * plane - c0;   // plane offsets of channels 0 and 2
    float* const osample = dst + (size_t)n * 3 * (size_t)plane;
    int ofirst = (ty0 + wy) * ow + tx0 + lx;
    float* osample_k = osample;
    keep_reg(osample_k); keep_reg(ofirst);

#pragma unroll 1
    for (int tile = 0; tile < kFuseTilesPerCta; ++tile) {
        const int row0 = tile * kFuseTile, nrow = min(kFuseTile, nrows - row0);
        if (nrow <= 0) break;                                                                   // uniform
        const int rh = s_oy1[row0 + nrow - 1] - s_oy0[row0] + 1;
        float* const im = s_im[tile % kBufs];

        // ---- stage S: the rectangle of rgb_processing's image into shared memory (lane = column, 8 rows per pass) ---
        if (lx < rw) {
            float* sp = im + wy * kPitchS + lx;
            const FuseSRow* srp = s_srow + row0 + wy;
#pragma unroll
            for (int k = 0; k < kPasses; ++k, sp += kRowStep * kPitchS, srp += kRowStep) {
                if (wy + k * kRowStep >= rh) break;
                float v3[3] = {0.f, 0.f, 0.f};
                if (s_col) {
                    const FuseSRow sr = *srp;
                    float p00[3], p01[3], p10[3], p11[3];
                    if (rotated) {
                        const unsigned r0 = (unsigned)(sr.y0 * iw * 3), r1 = (unsigned)(sr.y1 * iw * 3);
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            p00[c] = __ldg(rot0 + r0 + c); p01[c] = __ldg(rot1 + r0 + c);
                            p10[c] = __ldg(rot0 + r1 + c); p11[c] = __ldg(rot1 + r1 + c);
                        }
                    } else {
                        const int gy0 = uly + sr.y0, gy1 = uly + sr.y1;
                        if (col_in && gy0 >= 0 && gy1 < sh) {
                            if (pair_ok) {
                                row_px2(imgw + (unsigned)(gy0 * pitch), 0u, bsh, p00, p01);
                                row_px2(imgw + (unsigned)(gy1 * pitch), 0u, bsh, p10, p11);
                            } else {
                                image_px(img, pitch, gx0, gy0, p00); image_px(img, pitch, gx1, gy0, p01);
                                image_px(img, pitch, gx0, gy1, p10); image_px(img, pitch, gx1, gy1, p11);
                            }
                        } else {
                            canvas_px(img, pitch, sh, sw, d, x0, sr.y0, p00);
                            canvas_px(img, pitch, sh, sw, d, x1, sr.y0, p01);
                            canvas_px(img, pitch, sh, sw, d, x0, sr.y1, p10);
                            canvas_px(img, pitch, sh, sw, d, x1, sr.y1, p11);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float top = fmaf(p01[c], a1, p00[c] * a0), bot = fmaf(p11[c], a1, p10[c] * a0);
                        float v = fmaf(bot, sr.b1, top * sr.b0);
                        v = fminf(255.0f, fmaxf(0.0f, v * pn[c]));
                        v3[c] = v * (1.0f / 255.0f);
                    }
                }
                sp[0] = v3[0]; sp[kPlane] = v3[1]; sp[2 * kPlane] = v3[2];
            }
        }
        __syncthreads();

        // ---- stage T: torch bilinear from the rectangle (lane = output column, 4 rows per thread) --------------------
        if (t_col) {
            const uint32_t imb = smem_u32(im);
            uint32_t qa = imb + 4u * (unsigned)ox0, qb = imb + 4u * (unsigned)ox1;
            keep_reg(qa); keep_reg(qb);
            const FuseTRow* trp = s_trow + row0 + wy;
            int oi = ofirst + row0 * ow;
#pragma unroll
            for (int k = 0; k < kPasses; ++k, trp += kRowStep, oi += kRowStep * ow) {
                if (wy + k * kRowStep >= nrow) break;
                const FuseTRow tr = *trp;
                const float ly1 = tr.l1, ly0 = __fsub_rn(1.0f, ly1);
                const uint32_t q00 = qa + (unsigned)tr.off0, q01 = qb + (unsigned)tr.off0;      // off: bytes
                const uint32_t q10 = qa + (unsigned)tr.off1, q11 = qb + (unsigned)tr.off1;
                constexpr int P = kPlane * 4;
                const float t0 = __fadd_rn(__fmul_rn(lx0, lds_off<0>(q00)), __fmul_rn(lx1, lds_off<0>(q01)));
                const float b0 = __fadd_rn(__fmul_rn(lx0, lds_off<0>(q10)), __fmul_rn(lx1, lds_off<0>(q11)));
                const float t1 = __fadd_rn(__fmul_rn(lx0, lds_off<P>(q00)), __fmul_rn(lx1, lds_off<P>(q01)));
                const float b1 = __fadd_rn(__fmul_rn(lx0, lds_off<P>(q10)), __fmul_rn(lx1, lds_off<P>(q11)));
                const float t2 = __fadd_rn(__fmul_rn(lx0, lds_off<2 * P>(q00)), __fmul_rn(lx1, lds_off<2 * P>(q01)));
                const float b2 = __fadd_rn(__fmul_rn(lx0, lds_off<2 * P>(q10)), __fmul_rn(lx1, lds_off<2 * P>(q11)));
                const float v0 = fmaf(__fadd_rn(__fmul_rn(ly0, t0), __fmul_rn(ly1, b0)), na0, nb0);
                const float v1 = fmaf(__fadd_rn(__fmul_rn(ly0, t1), __fmul_rn(ly1, b1)), na1, nb1);
                const float v2 = fmaf(__fadd_rn(__fmul_rn(ly0, t2), __fmul_rn(ly1, b2)), na2, nb2);
                __stcs(osample_k + (oi + c0), v0);
                __stcs(osample_k + (oi + plane), v1);
                __stcs(osample_k + (oi + c2), v2);
            }
        }
        if (kBufs == 1 && tile + 1 < kFuseTilesPerCta) __syncthreads();      // the next tile overwrites the rectangle
    }
}

}  // namespace lh

extern "C" int lh_frei_rgb_processing(const uint8_t* src, int64_t src_img_stride, int src_h, int src_w, int src_pitch,
                                      const lh_frei_desc* desc, int N, int res, const int32_t* rot_index, int n_rot,
                                      float* scratch, int max_rot_w, int max_rot_h, float* out, lh_stream_t stream) {
    if (N < 0 || res <= 0 || src_h <= 0 || src_w <= 0 || src_pitch < 3 * src_w || max_rot_w < 0 || max_rot_h < 0 ||
        (max_rot_w > 0) != (max_rot_h > 0) || n_rot < 0 || n_rot > N)
        return LH_EINVAL;
    if (N == 0) return LH_OK;
    if (!src || !desc || !out || (max_rot_w > 0 && !scratch)) return LH_ENULL;
    if (N > 65535 || res > 32768 || max_rot_w > 32768 || max_rot_h > 32768) return LH_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (max_rot_w > 0) {
        const int mw = max_rot_w, mh = max_rot_h;
        const int nz = rot_index ? n_rot : N;      // with the list only the rotated samples get CTAs
        if (nz > 0)
            lh::frei_rotate_kernel<<<dim3((mw + lh::kFreiTileW - 1) / lh::kFreiTileW, (mh + lh::kRotTileH - 1) / lh::kRotTileH, nz),
                                     lh::kFreiThreads, 0, st>>>(src, (long long)src_img_stride, src_h, src_w, src_pitch, desc,
                                                                rot_index, scratch);
        LH_CHECK_LAUNCH();
    }
    const int rows = lh::kFreiThreads / 32;
    lh::frei_resize_kernel<<<dim3((res + 31) / 32, (res + rows - 1) / rows, N), lh::kFreiThreads, 0, st>>>(src, (long long)src_img_stride, src_h, src_w,
                                                                   src_pitch, desc, scratch, res, out);
    LH_CHECK_LAUNCH();
    return LH_OK;
}

extern "C" int lh_resize_normalize_f32(const float* src, int N, int src_h, int src_w, const float* mean_std,
                                       int chan_order, float* dst, int out_h, int out_w, lh_stream_t stream) {
    if (N < 0 || src_h <= 0 || src_w <= 0 || out_h <= 0 || out_w <= 0) return LH_EINVAL;
    if (chan_order != LH_CHAN_RGB && chan_order != LH_CHAN_BGR) return LH_EINVAL;
    if (N == 0) return LH_OK;
    if (!src || !dst) return LH_ENULL;
    if (N > 65535 || out_h > 65535 * 4 || (int64_t)src_h * src_w > (int64_t)1 << 30) return LH_EINVAL;
    const float* ms = mean_std;
    const int rows = lh::kFreiThreads / 64;
    const int vec = (out_w % 4 == 0) && lh::aligned16(dst);
    float a[3] = {1.f, 1.f, 1.f}, b[3] = {0.f, 0.f, 0.f};
    if (ms)
        for (int c = 0; c < 3; ++c) { a[c] = 1.0f / ms[3 + c]; b[c] = -ms[c] / ms[3 + c]; }
    const float sy = (float)src_h / (float)out_h, sx = (float)src_w / (float)out_w;     // ATen: float32 in / out
    lh::resize_f32_kernel<<<dim3((out_w + 255) / 256, (out_h + rows - 1) / rows, N), lh::kFreiThreads, 0,
                            (cudaStream_t)stream>>>(src, src_h, src_w, dst, out_h, out_w, sy, sx, a[0], a[1], a[2], b[0],
                                                    b[1], b[2], chan_order, vec);
    LH_CHECK_LAUNCH();
    return LH_OK;
}

// rgb_processing + Resize + Normalize (+ channel order) in one pass over the output (stages S and T fused; the rotation
// stage R still renders the rotated samples' canvases first).  Same results as lh_frei_rgb_processing followed by
// lh_resize_normalize_f32, bit for bit.  LH_EUNSUPPORTED when res / out > 0.9375 (the tile's rectangle of the intermediate
// image would not fit the shared-memory tile): use the two calls then.
extern "C" int lh_frei_fused(const uint8_t* src, int64_t src_img_stride, int src_h, int src_w, int src_pitch,
                             const lh_frei_desc* desc, int N, int res, const int32_t* rot_index, int n_rot,
                             float* scratch, int max_rot_w, int max_rot_h, const float* mean_std, int chan_order,
                             float* out, int out_h, int out_w, lh_stream_t stream) {
    if (N < 0 || res <= 0 || src_h <= 0 || src_w <= 0 || src_pitch < 3 * src_w || max_rot_w < 0 || max_rot_h < 0 ||
        (max_rot_w > 0) != (max_rot_h > 0) || n_rot < 0 || n_rot > N || out_h <= 0 || out_w <= 0)
        return LH_EINVAL;
    if (chan_order != LH_CHAN_RGB && chan_order != LH_CHAN_BGR) return LH_EINVAL;
    if (N == 0) return LH_OK;
    if (!src || !desc || !out || (max_rot_w > 0 && !scratch)) return LH_ENULL;
    if (N > 65535 || res > 32768 || max_rot_w > 32768 || max_rot_h > 32768) return LH_EINVAL;
    if (32.0 * res / out_w + 2.0 > lh::kFuseTile || 32.0 * res / out_h + 2.0 > lh::kFuseTile) return LH_EUNSUPPORTED;
    if ((int64_t)3 * out_h * out_w >= ((int64_t)1 << 31) || (int64_t)src_h * src_pitch >= ((int64_t)1 << 31)) return LH_EUNSUPPORTED;   // 32-bit in-sample offsets
    cudaStream_t st = (cudaStream_t)stream;
    if (max_rot_w > 0) {
        const int nz = rot_index ? n_rot : N;
        if (nz > 0)
            lh::frei_rotate_kernel<<<dim3((max_rot_w + lh::kFreiTileW - 1) / lh::kFreiTileW,
                                          (max_rot_h + lh::kRotTileH - 1) / lh::kRotTileH, nz),
                                     lh::kFreiThreads, 0, st>>>(src, (long long)src_img_stride, src_h, src_w, src_pitch, desc,
                                                                rot_index, scratch);
        LH_CHECK_LAUNCH();
    }
    float a[3] = {1.f, 1.f, 1.f}, b[3] = {0.f, 0.f, 0.f};
    if (mean_std)
        for (int c = 0; c < 3; ++c) { a[c] = 1.0f / mean_std[3 + c]; b[c] = -mean_std[c] / mean_std[3 + c]; }
    const float sy = (float)res / (float)out_h, sx = (float)res / (float)out_w;     // ATen: float32 in / out
    constexpr int kCtaRows = lh::kFuseTile * lh::kFuseTilesPerCta;
    lh::frei_fused_kernel<<<dim3((out_w + lh::kFuseTile - 1) / lh::kFuseTile, (out_h + kCtaRows - 1) / kCtaRows, N),
                            lh::kFreiThreads, 0, st>>>(src, (long long)src_img_stride, src_h, src_w, src_pitch, desc, scratch, res,
                                                       out, out_h, out_w, sy, sx, a[0], a[1], a[2], b[0], b[1], b[2], chan_order);
    LH_CHECK_LAUNCH();
    return LH_OK;
}
