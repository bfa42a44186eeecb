"""AverageMeter (reference: src/utils/metric_logger.py:7-28) and its device-resident counterpart."""
from __future__ import annotations

import torch

from .. import ops


class AverageMeter(object):
    """Same fields and update rules as the reference (host floats)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update_p(self, val, point):
        self.val = val
        self.sum += val
        self.count += point
        self.avg = self.sum / self.count

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class DeviceMeter:
    """The three AverageMeters of the valid loop (loss, PCK, EPE; src/utils/method.py:241-250) kept as
    one packed device buffer so an epoch costs ONE device->host read and, sharded over ranks, one
    all-reduce (lighthand_b200.distributed.all_reduce_metrics).

        counts int64[4]  = n_wrong, n_total, epe_count, n_images
        sums   float64[4] = epe_sum, sum(loss_b*n_b), sum(n_b), reserved
    """

    def __init__(self, device="cuda"):
        self.counts, self.sums = ops.new_metric_state(device)

    def reset(self):
        self.counts.zero_()
        self.sums.zero_()

    @property
    def loss_meter(self) -> torch.Tensor:
        return self.sums[1:3]

    def result(self, counts: torch.Tensor = None, sums: torch.Tensor = None) -> dict:
        """One synchronising read; values as the reference's meters report them.  ``counts`` / ``sums``:
        all-reduced copies of the accumulators (the local ones stay per-rank; reading twice is idempotent)."""
        c = (self.counts if counts is None else counts).cpu().tolist()
        s = (self.sums if sums is None else sums).cpu().tolist()
        n_wrong, n_total, epe_cnt, n_img = c
        return {
            "pck": (n_total - n_wrong) / n_total if n_total else 0.0,   # pck_losses.avg
            "epe": s[0] / epe_cnt if epe_cnt else 0.0,                   # epe_losses.avg
            "loss": s[1] / s[2] if s[2] else 0.0,                        # log_losses.avg
            "n_images": n_img, "n_total": n_total, "n_wrong": n_wrong, "epe_count": epe_cnt, "epe_sum": s[0],
        }
