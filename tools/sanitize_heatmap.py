#!/usr/bin/env python
"""Developer tool: a few small launches of the fused heat-map pass (whole-map units and split maps, train / eval / decode) for
`compute-sanitizer --tool racecheck|memcheck`; prints a checksum so a clean run is visibly the same computation."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lighthand_b200 import ops  # noqa: E402


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    acc = 0.0
    for (n, h) in ((3, 64), (40, 56), (2, 16)):
        joints = torch.rand((n, 21, 2), device="cuda", generator=g) * 4 * h
        pred = torch.randn((n, 21, h, h), device="cuda", generator=g)
        counts, sums = ops.new_metric_state("cuda")
        out = ops.heatmap_fused(pred, joints, out_scale=4.0, metrics=(counts, sums))
        ev = ops.heatmap_fused(pred, joints, want_grad=False, out_scale=4.0)
        co, mv, ai = ops.decode(pred, 4.0)
        torch.cuda.synchronize()
        assert torch.equal(co, out["coords"]) and torch.equal(ai, ev["argidx"])
        acc += float(out["loss"].item()) + float(out["coords"].sum().item()) + float(counts.sum().item())
    print("checksum %.6f" % acc)


if __name__ == "__main__":
    main()
