"""GenerateHeatmap with the reference's call surface (src/utils/dataset_loader.py:22-53)."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from .loss import _to_cuda


class GenerateHeatmap():
    """``GenerateHeatmap(output_res, num_parts)(p)`` with p = joint_2d / 4 (heat-map pixels).
    One sample [J,2] or a batch [N,J,2]; CPU/numpy input returns a numpy float32 array like the
    reference, CUDA input stays on the device."""

    def __init__(self, output_res, num_parts):
        self.output_res = output_res
        self.num_parts = num_parts
        self.sigma = self.output_res / 64

    def __call__(self, p):
        on_dev = isinstance(p, torch.Tensor) and p.is_cuda
        t = _to_cuda(p)
        single = t.dim() == 2
        if single:
            t = t[None]
        if t.shape[1] != self.num_parts:
            raise ops.LighthandError("GenerateHeatmap: wrong number of parts")
        hms, _ = ops.gen_targets(t, (self.output_res, self.output_res), (1.0, 1.0),
                                 style=ops.LH_TARGET_HOURGLASS, want_weight=False)
        if single:
            hms = hms[0]
        return hms if on_dev else hms.cpu().numpy()
