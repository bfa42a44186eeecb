#!/usr/bin/env python
"""Developer tool: device time of K1 (lh_warp_normalize, one-FMA normalisation) on bench.py's synthetic C2 / C3 / C5
augmentation mix, graph-replayed over a ring of cold batches.  LIGHTHAND_B200_LIB / LH_K1 select the variant."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from lighthand_b200 import ops  # noqa: E402
from lighthand_b200.utils.preprocessing import batch_patch_matrices  # noqa: E402
from tools.microbench import timeit  # noqa: E402


def main():
    res = {"lib": os.environ.get("LIGHTHAND_B200_LIB", "default"), "LH_K1": os.environ.get("LH_K1", "")}
    for name in sys.argv[1:] or ["C2"]:
        wl = bench.WORKLOADS[name]
        B, S = wl["batch"], wl["S"]
        R = 4
        sets = []
        for r in range(R):
            imgs, joints, bbox, scale, rot, flip = bench.synth_host(B, S, seed=123400 + r)
            trans, _ = batch_patch_matrices(bbox, flip, scale, rot, (S, S), (S, S))
            sets.append((torch.from_numpy(imgs).cuda(), torch.from_numpy(np.ascontiguousarray(trans.astype(np.float64))).cuda(),
                         torch.from_numpy(flip.astype(np.uint8)).cuda()))
        out = torch.empty((B, 3, S, S), device="cuda")
        us = timeit(lambda i: ops.warp_normalize(sets[i][0], sets[i][1], (S, S), flip=sets[i][2], out=out, exact_norm=False), R)
        nbytes = B * S * S * 15
        res[name] = {"us": round(us, 2), "GBps": round(nbytes / us / 1e3, 1), "frac_6516": round(nbytes / us / 1e3 / 6516.7, 3)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
