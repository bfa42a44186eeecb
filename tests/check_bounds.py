#!/usr/bin/env python
"""TEST INFRASTRUCTURE (uses the CPU oracle as the checker; not collected by pytest).
Bounds self-check of the fast warp kernel (compute-sanitizer is closed on the GPU pool).  Build the debug library first:
    cd lighthand_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC \
        -DLH_DEBUG_BOUNDS -c warp_normalize.cu -o /tmp/wn_dbg.o && \
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants_dbg.so build/capi.o build/heatmap.o \
        build/metrics.o /tmp/wn_dbg.o build/pipeline.o build/frei_crop.o build/color_jitter.o -cudart static
then  LIGHTHAND_B200_LIB=$PWD/variants_dbg.so python tests/check_bounds.py
Sweeps rotations / scales / flips / crops leaving the image at several sizes and reports how many shared-memory
indices fell outside the staged window (must be 0) and that the pixels still equal the CPU model."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lighthand_b200 import _lib, ops  # noqa: E402
from oracle import restate as R  # noqa: E402


def main():
    lib = _lib.load()
    cnt = ctypes.CDLL(_lib.LIB_PATH).lh_debug_oob_count
    cnt.restype = ctypes.c_ulonglong
    rng = np.random.default_rng(0)
    n_img, bad_px = 0, 0
    for (sh, sw), (oh, ow) in (((256, 256), (256, 256)), ((224, 224), (256, 256)), ((512, 512), (128, 128)),
                               ((100, 132), (96, 100)), ((64, 64), (224, 224))):
        n = 24
        imgs = rng.integers(0, 256, (n, sh, sw, 3), dtype=np.uint8)
        Ms, flips = [], []
        for i in range(n):
            side = rng.uniform(0.3, 1.6) * min(sh, sw)
            cx, cy = rng.uniform(-0.2, 1.2) * sw, rng.uniform(-0.2, 1.2) * sh
            Ms.append(R.gen_trans_from_patch_cv(cx, cy, side, side, ow, oh, rng.uniform(0.7, 1.3), rng.uniform(-180, 180)).astype(np.float64))
            flips.append(rng.uniform() < 0.5)
        out = ops.warp_normalize(torch.from_numpy(imgs).cuda(), torch.from_numpy(np.stack(Ms)).cuda(), (oh, ow),
                                 flip=torch.tensor(flips, dtype=torch.uint8).cuda(), mean=None, std=None,
                                 layout=ops.LH_LAYOUT_HWC).cpu().numpy()
        for i in range(n):
            want = R.warp_affine_model(imgs[i], Ms[i], oh, ow, flip=flips[i]).astype(np.float32)
            bad_px += int((out[i] != want).sum())
        n_img += n
    oob = int(cnt())
    print(f"images {n_img}  mismatching pixels {bad_px}  out-of-window shared-memory indices {oob}")
    return 0 if (oob == 0 and bad_px == 0) else 1


if __name__ == "__main__":
    sys.exit(main())
