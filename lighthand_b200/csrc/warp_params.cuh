// Kernel parameter block shared by the K1 kernels (warp_normalize.cu, warp_tma.cu).
#pragma once
#include "lh_common.cuh"

namespace lh {

struct WarpParams {
    const uint8_t* src;
    long long img_stride;
    int sh, sw, pitch;
    const double* M;
    const uint8_t* flip;
    const double* gain;
    const float* lut;
    int lut_per_image;
    int normalise;
    float mean[3], stdv[3];
    int chan_order, layout;
    float* dst;
    int N, oh, ow;
    int vec_store;
};

// The TMA-staged persistent K1 kernel (warp_tma.cu).  Returns LH_OK when it took the launch, kTmaNotApplicable when
// the geometry / alignment is outside what a tensor map can describe (the caller then uses warp_pairs_kernel),
// anything else is an error.  `norm`: 0 raw 0..255, 1 one FMA (p.mean = a, p.stdv = b), 2 table (p.lut).
constexpr int kTmaNotApplicable = 0x7fff0001;
int launch_warp_tma(const WarpParams& p, int norm, cudaStream_t st);
// LH_INTERP_TORCH (ToTensor -> Resize -> Normalize) on the same TMA ring; p.mean / p.stdv hold a, b of v * a + b (byte units)
int launch_resize_tma(const WarpParams& p, cudaStream_t st);

}  // namespace lh
