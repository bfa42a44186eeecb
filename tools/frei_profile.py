#!/usr/bin/env python
"""Developer helper: a few plain launches of the FreiHAND image path (no CUDA graph) for ncu."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lighthand_b200 import ops  # noqa: E402
from lighthand_b200.datasets import frei_dataloader as F  # noqa: E402
from lighthand_b200.utils import image_ops  # noqa: E402

N = 256
rng = np.random.RandomState(0)
imgs = torch.randint(0, 256, (N, 224, 224, 3), device="cuda", dtype=torch.uint8)
augm = [F.augm_params(True, rng) for _ in range(N)]
centers = 112 + rng.uniform(-10, 10, (N, 2))
recs, sfl, mpx = image_ops.crop_descriptors(centers, [a[3] for a in augm], [224, 224], [a[2] for a in augm], [a[1] for a in augm])
desc = image_ops.descriptors_to_device(recs, "cuda")
rot = torch.tensor([i for i, r in enumerate(recs) if r.rotated], dtype=torch.int32, device="cuda")
for _ in range(3):
    rgb = ops.frei_rgb_processing(imgs, desc, 224, sfl, mpx, rot_index=rot)
    out = ops.resize_normalize_f32(rgb, (256, 256), chan_order=ops.LH_CHAN_BGR)
torch.cuda.synchronize()
print("rotated", sum(r.rotated for r in recs), "max_wh", mpx, "scratch MB", sfl * 4 / 1e6)
