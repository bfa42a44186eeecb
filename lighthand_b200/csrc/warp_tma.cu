// K1, TMA-staged persistent kernel: batched cv2.warpAffine (INTER_LINEAR, BORDER_CONSTANT 0) of uint8 RGB
// images + flip + mean/std normalise -> float32 network input.
//
//   generate_patch_image  src/utils/preprocessing.py:144-162   (cv2.warpAffine; flip :153-155)
//   ToTensor / Normalize  src/tools/dataset.py:147-157
//
// Same arithmetic as warp_pairs_kernel (warp_normalize.cu; SURVEY.md Appendix A: OpenCV's double-precision matrix
// inversion, 1/32-px fixed-point grid, (sum p*w + 2^14) >> 15), different machine mapping:
//
//   * persistent CTAs (4 per SM), static round-robin over the 32x32 output tiles;
//   * one PRODUCER warp per CTA computes a tile's coordinate tables and source footprint (FP64, off the critical
//     path) and fetches the footprint with ONE 3-D TMA load (cp.async.bulk.tensor over a u8 tensor map
//     [N][src_h][3*src_w]; out-of-image bytes arrive as zeros = cv2's constant border, no predicates anywhere)
//     into a kStages-deep (3) ring of shared-memory stages; full/empty mbarriers hand the stages over;
//   * four CONSUMER warps render (lane = output column, a warp takes 8 tile rows) straight from the RAW bytes:
//     three aligned LDS.32 per tap row, two funnel shifts + two PRMT give {R0 R1 G0 G1} / {B0 B1 - -}, and each
//     channel's whole 2-D blend is two IDP.2A against the 16-bit weight pairs {32 wx0 wy, 32 wx1 wy} (cv2's own weights);
//     the >>15 and the int->float conversion are one funnel shift under the exponent of 2^23 and one FADD.
//
// No staging instructions, no per-tile barrier, no global-load latency on the rendering warps.
#include <cstdlib>
#include <cuda.h>   // CUtensorMap + enums (types only: the encoder is resolved through cudaGetDriverEntryPoint)

#include "warp_params.cuh"

#ifndef LH_TMA_STAGES
#define LH_TMA_STAGES 3
#endif
#ifndef LH_TMA_MINB
#define LH_TMA_MINB 6
#endif
#ifndef LH_RS_I2F
#define LH_RS_I2F 1     // resize ring: byte -> float conversions through I2F.U8 (XU pipe): 0 none, 1 half, 2 all -- measured 52.6 / 51.5 / 53.6 us
#endif
#ifndef LH_TMA_CTAS
#define LH_TMA_CTAS 4     // resident CTAs per SM the persistent grid is sized for (measured optimum)
#endif

namespace lh {

constexpr int kBoxW = 160;                 // bytes per staged source row = 53 RGB pixels (multiple of 16 for TMA)
constexpr int kBoxHS = 36, kBoxHL = 48;    // rows of the small (un-rotated tiles) / large (45 degrees at 1 px/px: 32*sqrt(2) + 2) box
constexpr int kStages = LH_TMA_STAGES;
constexpr int kConsWarps = 4;
constexpr int kTmaThreads = 32 * (kConsWarps + 1);
constexpr int kTmaTile = 32;               // output tile: 32 x 32
constexpr int kRawBytes = kBoxW * kBoxHL;  // 7680
// stage = raw bytes | 5 tables of 32 words (warp: ad bd x0 y0 -; resize: cx lx1 r0 r1 ly1) | StageInfo; a tap may read two words
// past its row: they land in the tables
struct __align__(16) StageInfo { int lox, hix, loy, hiy, flags, n, tx0, ty0, base, pad0, pad1, pad2; };
constexpr int kTabWords = 5 * 32;
constexpr int kStageBytes2 = ((kRawBytes + kTabWords * 4 + (int)sizeof(StageInfo)) + 127) & ~127;   // 8448
static_assert(kRawBytes % 128 == 0, "TMA destinations must stay 128-byte aligned");

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        :: "r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// footprint too large for the box (strong down-scaling): the tap words straight from global
static __device__ __noinline__ uint4 tma_taps_from_global(const uint8_t* img, int pitch_b, int sw, int sh, int ix, int iy) {
    uint32_t px[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int xx = ix + (t & 1), yy = iy + (t >> 1);
        uint32_t v = 0;
        if (xx >= 0 && xx < sw && yy >= 0 && yy < sh) {
            const uint8_t* q = img + (long long)yy * pitch_b + 3 * xx;
            v = __ldg(q) | (__ldg(q + 1) << 8) | (__ldg(q + 2) << 16);
        }
        px[t] = v;
    }
    return make_uint4(__byte_perm(px[0], px[1], 0x5140), __byte_perm(px[0], px[1], 0x0062),
                      __byte_perm(px[2], px[3], 0x5140), __byte_perm(px[2], px[3], 0x0062));
}

// One consumer warp renders rows 8*warp .. 8*warp+7 of the tile described by `inf` (lane = output column).
// GEO != 0: compile-time square CHW RGB output of GEO x GEO (a multiple of the tile): every store of the warp-tile is
// one base register + an immediate, no row / column guards.  GEO == 0: any size / layout / channel order.
template <int NORM, int GEO, bool STAGED>
__device__ __forceinline__ void render_rows(const WarpParams& p, const StageInfo& inf, const uint8_t* raw,
                                            const int* tab, const float* s_lut, int lane, int warp) {
    const int ow = GEO ? GEO : p.ow, oh = GEO ? GEO : p.oh;
    const int ox = inf.tx0 + lane;
    const int r0 = warp * 8;
    const int oy = inf.ty0 + r0;
    if (!GEO && (ox >= ow || oy >= oh)) return;
    const int n = inf.n;
    const int ad = tab[lane], bd = tab[32 + lane];
    const int4 xa = *reinterpret_cast<const int4*>(tab + 64 + r0), xb = *reinterpret_cast<const int4*>(tab + 64 + r0 + 4);
    const int4 ya = *reinterpret_cast<const int4*>(tab + 96 + r0), yb = *reinterpret_cast<const int4*>(tab + 96 + r0 + 4);
    const int x0v[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
    const int y0v[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
    // shared-window address of source pixel (0, 0); a multiple of 4 (the stage is 128-byte aligned, the box starts on a
    // 16-byte boundary, rows are kBoxW apart), so masking / shifting the sum below acts on the pixel offset alone
    const uint32_t org = smem_u32(raw) + (uint32_t)inf.base;
    const float na0 = p.mean[0], na1 = p.mean[1], na2 = p.mean[2];   // NORM 1: a, b of v*a + b
    const float nb0 = p.stdv[0], nb1 = p.stdv[1], nb2 = p.stdv[2];
    // GEO == 0: 32-bit element indices (the launcher keeps one launch below 2^31 output elements); BGR swaps the planes
    unsigned i0 = 0, i1 = 0, i2 = 0, step = 0;
    float* d0 = nullptr;
    if (GEO) {
        d0 = p.dst + (((size_t)n * 3 * GEO + oy) * GEO + ox);
    } else {
        if (p.layout == LH_LAYOUT_CHW) {
            const unsigned plane = (unsigned)(oh * ow);
            i0 = ((unsigned)(n * 3) * oh + oy) * ow + ox;
            i1 = i0 + plane; i2 = i1 + plane;
            step = (unsigned)ow;
        } else {
            i0 = (((unsigned)n * oh + oy) * ow + ox) * 3u;
            i1 = i0 + 1; i2 = i0 + 2;
            step = 3u * (unsigned)ow;
        }
        if (p.chan_order == LH_CHAN_BGR) { const unsigned t = i0; i0 = i2; i2 = t; }
    }
    const uint8_t* img = STAGED ? nullptr : p.src + (long long)n * p.img_stride;
    const int rows = GEO ? 8 : min(8, oh - oy);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (!GEO && j >= rows) break;
        const int X = x0v[j] + ad, Y = y0v[j] + bd;
        const int ix = X >> 10, iy = Y >> 10;
        // horizontal weights stay scaled by 32 (ax32 = 32 * ax is a plain mask of X: one instruction less than shift + mask);
        // the sum then carries cv2's own scale: (sum p * 32 wx wy + 2^14) >> 15, products still fit 16 bits (<= 32768)
        const uint32_t ax32 = (uint32_t)X & 0x3e0u, ay = (uint32_t)(Y >> 5) & 31u;
        uint32_t t0x, t0y, t1x, t1y;     // {R0 R1 G0 G1}, {B0 B1 - -} of the upper / lower tap row
        if (STAGED) {
            const uint32_t o = (uint32_t)(iy * kBoxW) + ((uint32_t)(ix * 3) + org);   // byte address of the tap origin
            const uint32_t* q = reinterpret_cast<const uint32_t*>(__cvta_shared_to_generic(o & ~3u));
            const uint32_t sh = o << 3;                               // funnel shifts use (o & 3) * 8
            const uint32_t a0 = q[0], b0 = q[1], c0 = q[2];
            const uint32_t a1 = q[kBoxW / 4], b1 = q[kBoxW / 4 + 1], c1 = q[kBoxW / 4 + 2];
            const uint32_t lo0 = __funnelshift_r(a0, b0, sh), hi0 = __funnelshift_r(b0, c0, sh);   // bytes o..o+3, o+4..o+7
            const uint32_t lo1 = __funnelshift_r(a1, b1, sh), hi1 = __funnelshift_r(b1, c1, sh);
            t0x = __byte_perm(lo0, hi0, 0x4130); t0y = __byte_perm(lo0, hi0, 0x0052);
            t1x = __byte_perm(lo1, hi1, 0x4130); t1y = __byte_perm(lo1, hi1, 0x0052);
        } else {
            const uint4 t = tma_taps_from_global(img, p.pitch, p.sw, p.sh, ix, iy);
            t0x = t.x; t0y = t.y; t1x = t.z; t1y = t.w;
        }
        const uint32_t wxp = ax32 * 65535u + 1024u;              // 16-bit pair {32 (32 - ax), 32 ax}
        const uint32_t w0 = wxp * (32u - ay), w1 = wxp * ay;     // cv2's weights 32 wx wy: {w00, w01}, {w10, w11}, each <= 32768
        const uint32_t sr = __dp2a_lo(w1, t1x, __dp2a_lo(w0, t0x, 16384u));
        const uint32_t sg = __dp2a_hi(w1, t1x, __dp2a_hi(w0, t0x, 16384u));
        const uint32_t sb = __dp2a_lo(w1, t1y, __dp2a_lo(w0, t0y, 16384u));
        float o0, o1, o2;
        if (NORM == 2) {
            o0 = s_lut[sr >> 15]; o1 = s_lut[256 + (sg >> 15)]; o2 = s_lut[512 + (sb >> 15)];
        } else {
            // (s >> 15) under the exponent of 2^23 in one funnel shift, minus 2^23: the exact float of the uint8 result
            const float vr = __uint_as_float(__funnelshift_r(sr, 0x2580u, 15)) - 8388608.0f;
            const float vg = __uint_as_float(__funnelshift_r(sg, 0x2580u, 15)) - 8388608.0f;
            const float vb = __uint_as_float(__funnelshift_r(sb, 0x2580u, 15)) - 8388608.0f;
            if (NORM == 1) { o0 = fmaf(vr, na0, nb0); o1 = fmaf(vg, na1, nb1); o2 = fmaf(vb, na2, nb2); }
            else { o0 = vr; o1 = vg; o2 = vb; }
        }
        if (GEO) {
            __stcs(d0 + j * GEO, o0); __stcs(d0 + GEO * GEO + j * GEO, o1); __stcs(d0 + 2 * GEO * GEO + j * GEO, o2);
        } else {
            __stcs(p.dst + i0, o0); __stcs(p.dst + i1, o1); __stcs(p.dst + i2, o2);
            i0 += step; i1 += step; i2 += step;
        }
    }
}

template <int NORM, int GEO>
__global__ void __launch_bounds__(kTmaThreads, LH_TMA_MINB)
warp_tma_kernel(const __grid_constant__ CUtensorMap tm_small, const __grid_constant__ CUtensorMap tm_large,
                const WarpParams p, int tiles_x, int tiles_y, int total_tiles) {
    extern __shared__ uint8_t s_dyn[];
    // 128-byte aligned carve-up: stages | full[kStages] empty[kStages] | lut
    uint8_t* s_base = s_dyn + ((128u - (smem_u32(s_dyn) & 127u)) & 127u);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_base + kStages * kStageBytes2);
    float* s_lut = reinterpret_cast<float*>(s_base + kStages * kStageBytes2 + 64);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(s_bar + s), 1);                       // full: the producer's arrive (+ TMA bytes)
            mbar_init(smem_u32(s_bar + kStages + s), kConsWarps);    // empty: one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (NORM == 2) {
        for (int i = tid; i < 768; i += kTmaThreads) s_lut[i] = p.lut[i];
    }
    __syncthreads();

    if (warp == kConsWarps) {
        // ---------------- producer: coordinate tables + footprint + one TMA load per tile ----------------
        const int tiles_img = tiles_x * tiles_y;
        const int mirror = 1024 * (p.sw - 1) + 31;
        int s = 0;
        uint32_t ph = 1;                                             // a fresh barrier passes a wait on parity 1
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int n = t / tiles_img, r = t - n * tiles_img;
            const int ty = r / tiles_x, tx = r - ty * tiles_x;
            const int tx0 = tx * kTmaTile, ty0 = ty * kTmaTile;
            const bool flip = p.flip ? (p.flip[n] != 0) : false;
            double im[6];
            invert_affine(p.M + (long long)n * 6, im);
            int ad = col_term(im[0], tx0 + lane);
            if (flip) ad = -ad;                                      // X' = 32(sw-1) - X  <=>  X0' = mirror - X0, adelta' = -adelta
            const int bd = col_term(im[3], tx0 + lane);
            int x0 = row_term(im[1], im[2], ty0 + lane);
            if (flip) x0 = mirror - x0;
            const int y0 = row_term(im[4], im[5], ty0 + lane);
            // the maps are monotone in x and in y: the footprint is spanned by the 4 tile corners
            const int cl = min(kTmaTile, p.ow - tx0) - 1, rl = min(kTmaTile, p.oh - ty0) - 1;
            const int a0 = __shfl_sync(0xffffffffu, ad, 0), a1 = __shfl_sync(0xffffffffu, ad, cl);
            const int b0 = __shfl_sync(0xffffffffu, bd, 0), b1 = __shfl_sync(0xffffffffu, bd, cl);
            const int xa = __shfl_sync(0xffffffffu, x0, 0), xb = __shfl_sync(0xffffffffu, x0, rl);
            const int ya = __shfl_sync(0xffffffffu, y0, 0), yb = __shfl_sync(0xffffffffu, y0, rl);
            const int X00 = (xa + a0) >> 10, X01 = (xa + a1) >> 10, X10 = (xb + a0) >> 10, X11 = (xb + a1) >> 10;
            const int Y00 = (ya + b0) >> 10, Y01 = (ya + b1) >> 10, Y10 = (yb + b0) >> 10, Y11 = (yb + b1) >> 10;
            const int ixmin = min(min(X00, X01), min(X10, X11)), ixmax = max(max(X00, X01), max(X10, X11));
            const int iymin = min(min(Y00, Y01), min(Y10, Y11)), iymax = max(max(Y00, Y01), max(Y10, Y11));
            StageInfo inf;
            // the box is placed on the footprint wherever it lies: bytes outside the image arrive as zeros, so tiles that
            // stick out of the source (or miss it altogether) need no clamping and no separate code path
            inf.lox = ixmin; inf.hix = ixmax; inf.loy = iymin; inf.hiy = iymax;
            // TMA fetches 16-byte granules: the box starts at the 16-byte boundary at or below column lox
            const int c0 = (3 * inf.lox) & ~15;
            const int wbytes = 3 * (inf.hix + 2) - c0;               // columns lox .. hix + 1
            const int nrows = inf.hiy - inf.loy + 2;                 // rows loy .. hiy + 1
            const bool staged = (wbytes <= kBoxW) & (nrows <= kBoxHL);
            inf.flags = staged ? 1 : 0;
            inf.n = n; inf.tx0 = tx0; inf.ty0 = ty0;
            inf.base = -inf.loy * kBoxW - c0;
            inf.pad0 = inf.pad1 = inf.pad2 = 0;

            uint8_t* stage = s_base + s * kStageBytes2;
            int* tab = reinterpret_cast<int*>(stage + kRawBytes);
            mbar_wait(smem_u32(s_bar + kStages + s), ph);            // the consumers are done with this stage
            tab[lane] = ad; tab[32 + lane] = bd; tab[64 + lane] = x0; tab[96 + lane] = y0;
            if (lane == 0) *reinterpret_cast<StageInfo*>(tab + kTabWords) = inf;
            __syncwarp();
            if (lane == 0) {
                const uint32_t full = smem_u32(s_bar + s);
                if (staged) {
                    const bool small = nrows <= kBoxHS;
                    mbar_expect_tx(full, (uint32_t)(kBoxW * (small ? kBoxHS : kBoxHL)));
                    tma_load_3d(smem_u32(stage), small ? &tm_small : &tm_large, full, c0, inf.loy, n);
                } else {
                    mbar_arrive(full);
                }
            }
            if (++s == kStages) { s = 0; ph ^= 1u; }
        }
    } else {
        // ---------------- consumers: render from the raw staged bytes ----------------
        int s = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const uint8_t* stage = s_base + s * kStageBytes2;
            const int* tab = reinterpret_cast<const int*>(stage + kRawBytes);
            mbar_wait(smem_u32(s_bar + s), ph);
            const StageInfo inf = *reinterpret_cast<const StageInfo*>(tab + kTabWords);   // 16-byte aligned: 3 x LDS.128
            if (inf.flags & 1) render_rows<NORM, GEO, true>(p, inf, stage, tab, s_lut, lane, warp);
            else               render_rows<NORM, GEO, false>(p, inf, stage, tab, s_lut, lane, warp);
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(s_bar + kStages + s));
            if (++s == kStages) { s = 0; ph ^= 1u; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// LH_INTERP_TORCH on the same ring: ToTensor -> Resize(bilinear, align_corners=False, no antialias) -> Normalize
// (src/tools/dataset.py:147-157, the LightHand99K dataset path).  Producer: torch's source index / lambda per output column
// and row of the tile (float32, ATen's fused scale * (dst + 0.5) - 0.5), the tile's source window, one TMA load.  Consumers
// (lane = output column, 8 rows per warp) interpolate each needed SOURCE row horizontally ONCE -- H(y) = lx0 p(y,x0) +
// lx1 p(y,x1) per channel, from the raw bytes: 3 LDS + 2 funnel shifts + 7 PRMT + 6 FADD -- into a per-warp slice of shared
// memory, then blend (H(y0), H(y1)) per output row: up-sampling 224 -> 256 needs 1.14 horizontal passes per output pixel instead of 2, i.e. 6.9 byte->float
// conversions instead of 12 (the record-staging kernel resize_pairs_kernel is bound by exactly those).  Same float operations as
// resize_pairs_kernel (value in byte units, one FMA normalises): identical results.  Right / bottom edge: torch clamps the
// second tap to the last pixel; here the lane's PRMT selector repeats the first tap's bytes / the row table repeats the row.
// ------------------------------------------------------------------------------------------------
constexpr int kHRows = 14;     // source rows 8 output rows can touch: 8 * 1.43 + 2

template <int GEO>
__global__ void __launch_bounds__(kTmaThreads, LH_TMA_MINB)
resize_tma_kernel(const __grid_constant__ CUtensorMap tm_small, const __grid_constant__ CUtensorMap tm_large,
                  const WarpParams p, int tiles_x, int tiles_y, int total_tiles) {
    extern __shared__ uint8_t s_dyn[];
    uint8_t* s_base = s_dyn + ((128u - (smem_u32(s_dyn) & 127u)) & 127u);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_base + kStages * kStageBytes2);
    float* s_h = reinterpret_cast<float*>(s_base + kStages * kStageBytes2 + 64);   // [kConsWarps][kHRows][3][32] horizontal rows
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(s_bar + s), 1);
            mbar_init(smem_u32(s_bar + kStages + s), kConsWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int ow = GEO ? GEO : p.ow, oh = GEO ? GEO : p.oh;

    if (warp == kConsWarps) {
        // ---------------- producer ----------------
        const float sy = __fdiv_rn((float)p.sh, (float)oh), sx = __fdiv_rn((float)p.sw, (float)ow);
        auto src_index = [](float scale, int d) {
            const float f = fmaf(scale, __fadd_rn((float)d, 0.5f), -0.5f);
            return f < 0.f ? 0.f : f;
        };
        const int tiles_img = tiles_x * tiles_y;
        int s = 0;
        uint32_t ph = 1;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const int n = t / tiles_img, r = t - n * tiles_img;
            const int ty = r / tiles_x, tx = r - ty * tiles_x;
            const int tx0 = tx * kTmaTile, ty0 = ty * kTmaTile;
            const int cl = min(kTmaTile, ow - tx0) - 1, rl = min(kTmaTile, oh - ty0) - 1;
            const float fx = src_index(sx, tx0 + min(lane, cl));
            const int x0 = min((int)fx, p.sw - 1);
            const float lx1 = fminf(fmaxf(__fsub_rn(fx, (float)x0), 0.f), 1.f);
            const float fy = src_index(sy, ty0 + min(lane, rl));
            const int y0 = min((int)fy, p.sh - 1), y1 = min(y0 + 1, p.sh - 1);
            const float ly1 = fminf(fmaxf(__fsub_rn(fy, (float)y0), 0.f), 1.f);
            StageInfo inf;
            inf.lox = __shfl_sync(0xffffffffu, x0, 0); inf.hix = __shfl_sync(0xffffffffu, x0, cl);     // window columns lox .. hix + 1
            inf.loy = __shfl_sync(0xffffffffu, y0, 0); inf.hiy = __shfl_sync(0xffffffffu, y1, rl);     // window rows loy .. hiy
            const int c0 = (3 * inf.lox) & ~15;
            const int nrows = inf.hiy - inf.loy + 1;
            inf.flags = 1; inf.n = n; inf.tx0 = tx0; inf.ty0 = ty0;
            inf.base = -inf.loy * kBoxW - c0;
            inf.pad0 = inf.pad1 = inf.pad2 = 0;
            uint8_t* stage = s_base + s * kStageBytes2;
            int* tab = reinterpret_cast<int*>(stage + kRawBytes);
            mbar_wait(smem_u32(s_bar + kStages + s), ph);
            tab[lane] = 3 * x0; tab[32 + lane] = __float_as_int(lx1);
            tab[64 + lane] = y0; tab[96 + lane] = y1; tab[128 + lane] = __float_as_int(ly1);
            if (lane == 0) *reinterpret_cast<StageInfo*>(tab + kTabWords) = inf;
            __syncwarp();
            if (lane == 0) {
                const uint32_t full = smem_u32(s_bar + s);
                const bool small = nrows <= kBoxHS;
                mbar_expect_tx(full, (uint32_t)(kBoxW * (small ? kBoxHS : kBoxHL)));
                tma_load_3d(smem_u32(stage), small ? &tm_small : &tm_large, full, c0, inf.loy, n);
            }
            if (++s == kStages) { s = 0; ph ^= 1u; }
        }
    } else {
        // ---------------- consumers ----------------
        const float na0 = p.mean[0], na1 = p.mean[1], na2 = p.mean[2];   // a, b of v * a + b (v in byte units)
        const float nb0 = p.stdv[0], nb1 = p.stdv[1], nb2 = p.stdv[2];
        int s = 0;
        uint32_t ph = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            const uint8_t* stage = s_base + s * kStageBytes2;
            const int* tab = reinterpret_cast<const int*>(stage + kRawBytes);
            mbar_wait(smem_u32(s_bar + s), ph);
            const StageInfo inf = *reinterpret_cast<const StageInfo*>(tab + kTabWords);
            const int ox = inf.tx0 + lane, r0 = warp * 8, oy = inf.ty0 + r0;
            if (GEO || (ox < ow && oy < oh)) {
                const int cx = tab[lane];
                const float lx1 = __int_as_float(tab[32 + lane]), lx0 = __fsub_rn(1.0f, lx1);
                // second tap = the next pixel's bytes 3..5 of the window, or the first tap again at the right edge (torch's clamp)
                const uint32_t sel1 = (cx == 3 * (p.sw - 1)) ? 0x3210u : 0x6543u;
                const uint32_t org = smem_u32(stage) + (uint32_t)inf.base + (uint32_t)cx;
                const int rows = GEO ? 8 : min(8, oh - oy);
                float* d0 = nullptr;
                unsigned i0 = 0, i1 = 0, i2 = 0, step = 0;
                if (GEO) {
                    d0 = p.dst + (((size_t)inf.n * 3 * GEO + oy) * GEO + ox);
                } else {
                    if (p.layout == LH_LAYOUT_CHW) {
                        const unsigned plane = (unsigned)(oh * ow);
                        i0 = ((unsigned)(inf.n * 3) * oh + oy) * ow + ox;
                        i1 = i0 + plane; i2 = i1 + plane;
                        step = (unsigned)ow;
                    } else {
                        i0 = (((unsigned)inf.n * oh + oy) * ow + ox) * 3u;
                        i1 = i0 + 1; i2 = i0 + 2;
                        step = 3u * (unsigned)ow;
                    }
                    if (p.chan_order == LH_CHAN_BGR) { const unsigned tt = i0; i0 = i2; i2 = tt; }
                }
                // pass 1: every source row the warp's 8 output rows touch, horizontally interpolated ONCE (byte units, three
                // channels) into the warp's slice of shared memory -- no row-to-row dependence, no data-dependent branches
                const int ys = tab[64 + r0], ye = tab[96 + r0 + rows - 1];           // warp-uniform: first .. last source row
                float* hbuf = s_h + warp * (kHRows * 3 * 32) + lane;
                for (int y = ys; y <= ye; ++y) {
                    const uint32_t o = (uint32_t)(y * kBoxW) + org;
                    const uint32_t* q = reinterpret_cast<const uint32_t*>(__cvta_shared_to_generic(o & ~3u));
                    const uint32_t sh = o << 3;
                    const uint32_t a = q[0], b = q[1], c = q[2];
                    const uint32_t lo = __funnelshift_r(a, b, sh), hi = __funnelshift_r(b, c, sh);   // R0 G0 B0 R1 | G1 B1 - -
                    const uint32_t t1 = __byte_perm(lo, hi, sel1);                                   // R1 G1 B1 -  (or R0 G0 B0 -)
                    auto bytef = [](uint32_t w, uint32_t k) {   // byte k of w as float without I2F: under the exponent of 2^23
                        return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u | k)) - 8388608.0f;
                    };
                    auto bytei = [](uint32_t w, uint32_t k) { return (float)((w >> (8 * k)) & 0xffu); };   // I2F.U8 (XU pipe)
                    float* h = hbuf + (y - ys) * (3 * 32);
#if LH_RS_I2F == 2
                    h[0] = fmaf(lx1, bytei(t1, 0), lx0 * bytei(lo, 0));
                    h[32] = fmaf(lx1, bytei(t1, 1), lx0 * bytei(lo, 1));
                    h[64] = fmaf(lx1, bytei(t1, 2), lx0 * bytei(lo, 2));
#elif LH_RS_I2F == 1
                    h[0] = fmaf(lx1, bytef(t1, 0), lx0 * bytef(lo, 0));
                    h[32] = fmaf(lx1, bytei(t1, 1), lx0 * bytef(lo, 1));
                    h[64] = fmaf(lx1, bytei(t1, 2), lx0 * bytei(lo, 2));
#else
                    h[0] = fmaf(lx1, bytef(t1, 0), lx0 * bytef(lo, 0));
                    h[32] = fmaf(lx1, bytef(t1, 1), lx0 * bytef(lo, 1));
                    h[64] = fmaf(lx1, bytef(t1, 2), lx0 * bytef(lo, 2));
#endif
                }
                // pass 2: the 8 output rows, independent of one another (a lane reads back only what it wrote: no barrier)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (!GEO && j >= rows) break;
                    const float* h0 = hbuf + (tab[64 + r0 + j] - ys) * (3 * 32);
                    const float* h1 = hbuf + (tab[96 + r0 + j] - ys) * (3 * 32);
                    const float ly1 = __int_as_float(tab[128 + r0 + j]), ly0 = __fsub_rn(1.0f, ly1);
                    const float vr = fmaf(ly1, h1[0], ly0 * h0[0]), vg = fmaf(ly1, h1[32], ly0 * h0[32]),
                                vb = fmaf(ly1, h1[64], ly0 * h0[64]);
                    const float o0 = fmaf(vr, na0, nb0), o1 = fmaf(vg, na1, nb1), o2 = fmaf(vb, na2, nb2);
                    if (GEO) {
                        __stcs(d0 + j * GEO, o0); __stcs(d0 + GEO * GEO + j * GEO, o1); __stcs(d0 + 2 * GEO * GEO + j * GEO, o2);
                    } else {
                        __stcs(p.dst + i0, o0); __stcs(p.dst + i1, o1); __stcs(p.dst + i2, o2);
                        i0 += step; i1 += step; i2 += step;
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(s_bar + kStages + s));
            if (++s == kStages) { s = 0; ph ^= 1u; }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps (thread-local cache: a training loop presents the same few buffers again and again)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

struct MapKey {
    const void* src; long long stride; int n, sh, sw, pitch;
    bool operator==(const MapKey& o) const {
        return src == o.src && stride == o.stride && n == o.n && sh == o.sh && sw == o.sw && pitch == o.pitch;
    }
};
struct MapEntry { MapKey key; CUtensorMap small, large; bool valid; };

static bool make_maps(const MapKey& k, CUtensorMap* small, CUtensorMap* large) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)3 * k.sw, (cuuint64_t)k.sh, (cuuint64_t)k.n};
    const cuuint64_t strides[2] = {(cuuint64_t)k.pitch, (cuuint64_t)k.stride};
    const cuuint32_t es[3] = {1, 1, 1};
    for (int i = 0; i < 2; ++i) {
        const cuuint32_t box[3] = {(cuuint32_t)kBoxW, (cuuint32_t)(i ? kBoxHL : kBoxHS), 1};
        const CUresult r = enc(i ? large : small, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(k.src), dims, strides,
                               box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return false;
    }
    return true;
}

template <int NORM, int GEO>
static int launch_norm(const CUtensorMap& ms, const CUtensorMap& ml, const WarpParams& p, cudaStream_t st) {
    static int cap = 0;     // resident CTAs of this instantiation on the device model (B200); benign race: same value
    static int sms_cached = 0;
    const size_t smem = 128 + (size_t)kStages * kStageBytes2 + 64 + (NORM == 2 ? 768 * sizeof(float) : 0);
    if (cap <= 0) {
        int dev = 0, sms = 0, occ = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return (int)cudaGetLastError();
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return (int)cudaGetLastError();
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(warp_tma_kernel<NORM, GEO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return (int)cudaGetLastError();
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, warp_tma_kernel<NORM, GEO>, kTmaThreads, smem) != cudaSuccess)
            return (int)cudaGetLastError();
        sms_cached = sms;
        cap = sms * (occ < 1 ? 1 : occ);
    }
    const int tiles_x = (p.ow + kTmaTile - 1) / kTmaTile, tiles_y = (p.oh + kTmaTile - 1) / kTmaTile;
    const long long total = (long long)tiles_x * tiles_y * p.N;
    if (total >= (1LL << 31)) return LH_EUNSUPPORTED;
    // FOUR CTAs per SM, not the eight that fit: measured on B200 (C2 / C3 / C5, 3 stages) 8: 49.0 / 38.8 / 99.1 us, 6: 48.0 / 37.2 / 96.1,
    // 5: 45.7 / 36.2 / 90.0, 4: 43.9 / 35.2 / 87.2, 3: 48.5 / 38.8 / 88.4, 2: 57.2 / 44.4 / 102 -- the kernel is bound by its store
    // stream, fewer tiles in flight keep it more sequential, and below 4 the 16 rendering warps per SM run out of issue slots.
    // LH_K1_CTAS in the environment overrides it for such sweeps.
    long long want = cap;
    {
        static const int env_ctas = [] { const char* e = getenv("LH_K1_CTAS"); return e ? atoi(e) : 0; }();
        const long long per_sm = env_ctas > 0 ? env_ctas : LH_TMA_CTAS;
        if (per_sm * sms_cached < want) want = per_sm * sms_cached;
    }
    const unsigned grid = (unsigned)(total < want ? total : want);
    warp_tma_kernel<NORM, GEO><<<grid, kTmaThreads, smem, st>>>(ms, ml, p, tiles_x, tiles_y, (int)total);
    return LH_OK;
}

template <int NORM>
static int launch_geo(const CUtensorMap& ms, const CUtensorMap& ml, const WarpParams& p, cudaStream_t st) {
    // the network-input sizes of the reference's configurations get compile-time store addressing
    const bool plain = p.layout == LH_LAYOUT_CHW && p.chan_order == LH_CHAN_RGB && p.oh == p.ow;
    if (NORM != 0 && plain && p.ow == 256) return launch_norm<NORM, 256>(ms, ml, p, st);
    if (NORM != 0 && plain && p.ow == 224) return launch_norm<NORM, 224>(ms, ml, p, st);
    if (NORM != 0 && plain && p.ow == 512) return launch_norm<NORM, 512>(ms, ml, p, st);
    return launch_norm<NORM, 0>(ms, ml, p, st);
}

template <int GEO>
static int launch_resize_geo(const CUtensorMap& ms, const CUtensorMap& ml, const WarpParams& p, cudaStream_t st) {
    static int cap = 0, sms_cached = 0;
    const size_t smem = 128 + (size_t)kStages * kStageBytes2 + 64 + (size_t)kConsWarps * kHRows * 3 * 32 * sizeof(float);
    if (cap <= 0) {
        int dev = 0, sms = 0, occ = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return (int)cudaGetLastError();
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return (int)cudaGetLastError();
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, resize_tma_kernel<GEO>, kTmaThreads, smem) != cudaSuccess)
            return (int)cudaGetLastError();
        sms_cached = sms;
        cap = sms * (occ < 1 ? 1 : occ);
    }
    const int tiles_x = (p.ow + kTmaTile - 1) / kTmaTile, tiles_y = (p.oh + kTmaTile - 1) / kTmaTile;
    const long long total = (long long)tiles_x * tiles_y * p.N;
    if (total >= (1LL << 31)) return LH_EUNSUPPORTED;
    long long want = cap;
    {
        static const int env_ctas = [] { const char* e = getenv("LH_K1_CTAS"); return e ? atoi(e) : 0; }();
        const long long per_sm = env_ctas > 0 ? env_ctas : LH_TMA_CTAS;
        if (per_sm * sms_cached < want) want = per_sm * sms_cached;
    }
    const unsigned grid = (unsigned)(total < want ? total : want);
    resize_tma_kernel<GEO><<<grid, kTmaThreads, smem, st>>>(ms, ml, p, tiles_x, tiles_y, (int)total);
    return LH_OK;
}

static const struct MapEntry* lookup_maps(const WarpParams& p);

// LH_INTERP_TORCH through the ring (p.mean / p.stdv = a, b of v * a + b in byte units).  kTmaNotApplicable: alignment, flip /
// gain, or a source window that does not fit the box (down-sampling by more than ~1.4) -- the caller keeps its other kernels.
int launch_resize_tma(const WarpParams& p, cudaStream_t st) {
    if (((reinterpret_cast<uintptr_t>(p.src) | (uintptr_t)p.pitch | (uintptr_t)p.img_stride) & 15u) != 0) return kTmaNotApplicable;
    if (p.flip || p.gain) return kTmaNotApplicable;
    // window of a 32 x 32 output tile: columns 32 * sw / ow + 2 (x 3 bytes, + 15 of box alignment), rows 32 * sh / oh + 2
    if (3.0 * (32.0 * p.sw / p.ow + 3.0) + 15.0 > kBoxW || 32.0 * p.sh / p.oh + 3.0 > kBoxHL) return kTmaNotApplicable;
    const MapEntry* e = lookup_maps(p);
    if (!e) return kTmaNotApplicable;
    const bool plain = p.layout == LH_LAYOUT_CHW && p.chan_order == LH_CHAN_RGB && p.oh == p.ow;
    int rc;
    if (plain && p.ow == 256)      rc = launch_resize_geo<256>(e->small, e->large, p, st);
    else if (plain && p.ow == 224) rc = launch_resize_geo<224>(e->small, e->large, p, st);
    else                           rc = launch_resize_geo<0>(e->small, e->large, p, st);
    if (rc) return rc;
    LH_CHECK_LAUNCH();
    return LH_OK;
}

int launch_warp_tma(const WarpParams& p, int norm, cudaStream_t st) {
    // what a tiled tensor map can describe: 16-byte aligned base and strides
    if (((reinterpret_cast<uintptr_t>(p.src) | (uintptr_t)p.pitch | (uintptr_t)p.img_stride) & 15u) != 0) return kTmaNotApplicable;
    if (p.gain || p.lut_per_image) return kTmaNotApplicable;      // per-image tables stay on the record kernel
    const MapEntry* e = lookup_maps(p);
    if (!e) return kTmaNotApplicable;
    int rc;
    if (norm == 2)      rc = launch_geo<2>(e->small, e->large, p, st);
    else if (norm == 1) rc = launch_geo<1>(e->small, e->large, p, st);
    else                rc = launch_geo<0>(e->small, e->large, p, st);
    if (rc) return rc;
    LH_CHECK_LAUNCH();
    return LH_OK;
}

// tensor maps of this source batch (thread-local cache: a training loop presents the same few buffers again and again)
static const MapEntry* lookup_maps(const WarpParams& p) {
    const MapKey key{p.src, p.img_stride, p.N, p.sh, p.sw, p.pitch};
    constexpr int kCache = 8;
    thread_local MapEntry cache[kCache];
    thread_local int next = 0;
    for (int i = 0; i < kCache; ++i) if (cache[i].valid && cache[i].key == key) return &cache[i];
    MapEntry& slot = cache[next];
    next = (next + 1) % kCache;
    slot.valid = false;
    if (!make_maps(key, &slot.small, &slot.large)) return nullptr;
    slot.key = key; slot.valid = true;
    return &slot;
}

}  // namespace lh
