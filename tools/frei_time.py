"""Timing of the FreiHAND image path (rows a-3 / f-2): rotate + fused resize kernels for three augmentation mixes.
Usage: python tools/frei_time.py   (one JSON line; graph-free, CUDA events over 30 launches after 3 warm-ups)"""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from lighthand_b200 import ops
from lighthand_b200.datasets import frei_dataloader as F
from lighthand_b200.utils import image_ops


def timeit(fn, iters=30, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def main():
    dev = torch.device("cuda:0")
    N, S, R = 256, 256, 4
    g = torch.Generator(device=dev).manual_seed(0)
    fimgs = [torch.randint(0, 256, (N, 224, 224, 3), device=dev, dtype=torch.uint8, generator=g) for _ in range(R)]
    out = torch.empty((N, 3, S, S), device=dev)
    res = {}
    for mix in ("train", "none", "all"):
        rng = np.random.RandomState(0)
        augm = [F.augm_params(True, rng) for _ in range(N)]
        rots = [a[2] for a in augm]
        if mix == "none":
            rots = [0.0] * N
        elif mix == "all":
            rots = [r if r != 0 else 17.0 for r in rots]
        centers = 112 + rng.uniform(-10, 10, (N, 2))
        recs, sfl, mpx = image_ops.crop_descriptors(centers, [a[3] for a in augm], [224, 224], rots, [a[1] for a in augm])
        desc = image_ops.descriptors_to_device(recs, dev)
        rot = torch.tensor([i for i, r in enumerate(recs) if r.rotated], dtype=torch.int32, device=dev)
        us = timeit(lambda i: ops.frei_fused(fimgs[i % R], desc, 224, sfl, mpx, (S, S), chan_order=ops.LH_CHAN_BGR,
                                             rot_index=rot, out=out))
        nbytes = N * (224 * 224 * 3 + S * S * 12)
        res[mix] = {"us": round(us, 1), "n_rot": int(rot.numel()), "GBps": round(nbytes / us / 1e3, 1),
                    "Mimg_s": round(N / us, 3)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
