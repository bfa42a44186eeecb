#!/usr/bin/env python
"""Developer tool: device time of the post-backbone pass variants (graph-replayed over a ring of cold batches).
LH_HM=stream selects the streaming warp-per-map kernel instead of the TMA ring kernel."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lighthand_b200 import ops  # noqa: E402
from tools.microbench import timeit  # noqa: E402


def main():
    res = {"LH_HM": os.environ.get("LH_HM", ""), "lib": os.environ.get("LIGHTHAND_B200_LIB", "default")}
    for (N, H, sigma) in ((256, 64, 2), (256, 56, 2), (128, 128, 3)):
        J, R = 21, 4
        S = 4 * H
        g = torch.Generator(device="cuda").manual_seed(0)
        joints = [torch.rand((N, J, 2), device="cuda", generator=g) * S for _ in range(R)]
        pred = [torch.randn((N, J, H, H), device="cuda", generator=g) for _ in range(R)]
        bufs = {"grad": torch.empty_like(pred[0])}
        nb = N * J * H * H * 4
        out = {}
        us = timeit(lambda i: ops.decode(pred[i], 4.0), R)
        out["decode"] = [round(us, 2), round(nb / us / 1e3)]
        us = timeit(lambda i: ops.heatmap_fused(pred[i], joints[i], sigma=sigma, want_grad=False, buffers=dict(bufs)), R)
        out["eval"] = [round(us, 2), round(nb / us / 1e3)]
        us = timeit(lambda i: ops.heatmap_fused(pred[i], joints[i], sigma=sigma, buffers=dict(bufs)), R)
        out["train"] = [round(us, 2), round(2 * nb / us / 1e3)]
        res[f"{N}x{H}"] = out
    print(json.dumps(res))


if __name__ == "__main__":
    main()
