#!/usr/bin/env python
"""Per-phase clock64() breakdown of the fast warp kernel (debug build with -DLH_PHASE_TIMING, loaded through
LIGHTHAND_B200_LIB).  Prints the mean cycles a CTA spends in prologue / staging / render under real contention."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lighthand_b200 import _lib, ops  # noqa: E402
import bench  # noqa: E402
from lighthand_b200.utils.preprocessing import batch_patch_matrices  # noqa: E402


def main():
    _lib.load()
    dbg = ctypes.CDLL(_lib.LIB_PATH).lh_debug_phase_cycles
    N, S = 256, 256
    imgs, joints, bbox, scale, rot, flip = bench.synth_host(N, S, 7)
    trans, _ = batch_patch_matrices(bbox, flip, scale, rot, (S, S), (S, S))
    d_img = torch.from_numpy(imgs).cuda()
    d_M = torch.from_numpy(np.ascontiguousarray(trans.astype(np.float64))).cuda()
    d_f = torch.from_numpy(np.ascontiguousarray(flip.astype(np.uint8))).cuda()
    out = torch.empty((N, 3, S, S), device="cuda")
    buf = (ctypes.c_ulonglong * 4)()
    for _ in range(3):
        ops.warp_normalize(d_img, d_M, (S, S), flip=d_f, out=out, exact_norm=False)
    dbg(buf, 1)
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.warp_normalize(d_img, d_M, (S, S), flip=d_f, out=out, exact_norm=False)
    e1.record()
    dbg(buf, 0)
    ctas = reps * (S // 32) * (S // 32) * N
    names = ["prologue (to barrier 1)", "staging (to barrier 2)", "render (warp 0)", "whole CTA (warp 0)"]
    print(f"kernel {e0.elapsed_time(e1) / reps * 1e3:.1f} us, {ctas // reps} CTAs per launch")
    for nme, v in zip(names, buf):
        print(f"  {nme:28s} {v / ctas:9.0f} cycles per CTA  = {v / ctas / 1.965e3:6.2f} us")


if __name__ == "__main__":
    main()
