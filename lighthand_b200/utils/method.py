"""Device-resident train / valid loops with the reference's call surface (src/utils/method.py:12-308).

``Runner_t(...).run(end)`` / ``Runner_v(train_runner, model, loader, phase, best_loss).run(end)`` keep the
reference's constructor arguments and return values:

    TRAIN  -> (model, optimizer, batch_time)
    VALID  -> (log_losses.avg, count, pck_losses.avg * 100, batch_time)

but one iteration no longer crosses the host <-> device boundary four times (``.cuda()`` x2, ``.item()``,
``.cpu()`` + numpy decode + ``torch.tensor(...).cuda()``, reference lines 165-172 / 225-250).  After the
unchanged backbone produced ``pred``, ONE fused launch (``ops.heatmap_fused``) yields the JointsMSELoss value,
its gradient w.r.t. ``pred``, the decoded joints (x4) and -- in the valid phase -- the PCK@0.2 / EPE counts,
accumulated in a ``DeviceMeter``; the host reads the three averages once per ``args.logging_steps`` iterations
(for the progress bar) and once at the end.

The loaders yield ``(images, gt_2d_joints, gt_heatmaps)`` exactly like the reference's datasets.  The heat-maps a
loader provides are used as the loss target (the reference's semantics); with ``gt_heatmaps=None`` in the batch,
or ``regenerate_targets=True``, they are recomputed on the device from the joints (bit-identical to
``generate_target``) and never touch HBM.  The matplotlib dumps of the reference are not reproduced.
"""
from __future__ import annotations

import time
from time import ctime

import torch

from .. import ops
from .metric_logger import AverageMeter, DeviceMeter


class _MeterView:
    """AverageMeter-shaped read-only view of one DeviceMeter average (refreshed by Runner._sync)."""

    def __init__(self):
        self.avg = 0.0
        self.val = 0.0

    def reset(self):
        self.avg = 0.0
        self.val = 0.0


class Runner_t(object):
    def __init__(self, args, model, epoch, train_loader, phase, batch_time, logger, data_len, len_total, count,
                 writer, optimizer, regenerate_targets=False, sigma=2, multiply=4):
        super(Runner_t, self).__init__()
        self.args = args
        self.logger = logger
        self.len_data = data_len
        self.len_total = len_total
        self.count = count
        self.phase = phase
        self.writer = writer
        self.train_loader = train_loader
        self.batch_time = batch_time if batch_time is not None else AverageMeter()
        self.now_loader = train_loader
        self.bar = None                       # the reference's progress bar (src/utils/bar.py) is console UX: optional
        self.model = model
        self.optimizer = optimizer
        self.epoch = epoch
        self.regenerate_targets = regenerate_targets
        self.sigma = sigma
        self.multiply = multiply
        self.device = next(model.parameters()).device
        self.meter = DeviceMeter(self.device)
        self.log_losses = _MeterView()
        self.pck_losses = _MeterView()
        self.epe_losses = _MeterView()
        self.best_loss = float("inf")
        self._buffers = {}

    # ------------------------------------------------------------------------------------------
    def _sync(self):
        """One device->host read of the packed accumulators (64 bytes)."""
        r = self.meter.result()
        self.log_losses.avg = r["loss"]
        self.pck_losses.avg = r["pck"]
        self.epe_losses.avg = r["epe"]
        return r

    def _log(self, iteration, eta_seconds, end):
        steps = max(1, int(getattr(self.args, "logging_steps", 100)))
        last = iteration == len(self.now_loader) - 1
        if not (last or iteration % (steps * 5) == 0):
            return
        self._sync()
        if self.logger is None:
            return
        tt = " ".join(ctime(eta_seconds + end).split(" ")[1:-1])
        if self.phase == "TRAIN":
            self.logger.debug("dataset_length: {} epoch: {} iter: {} /{}, count: {}/{} lr: {:.6f} loss: {:.8f}".format(
                self.len_data, self.epoch, iteration, len(self.now_loader), self.count, getattr(self.args, "count", 0),
                self.optimizer.param_groups[0]["lr"], self.log_losses.avg))
        elif last:
            self.logger.debug("Test =>> epoch: {} iter: {} /{} epe: {:.2f}mm, count: {} / {}, total_pck: {:.2f} %, "
                              "best_loss: {:.7f} , expected_date: {}".format(
                                  self.epoch, iteration, len(self.now_loader), self.epe_losses.avg * 0.26, int(self.count),
                                  getattr(self.args, "count", 0), self.pck_losses.avg * 100, self.best_loss, tt))

    def _fused(self, pred, gt_2d_joints, gt_heatmaps, train):
        n, j, h, w = pred.shape
        joints = gt_2d_joints.to(self.device, torch.float32, non_blocking=True)
        use_joints = self.regenerate_targets or gt_heatmaps is None
        target = None if use_joints else gt_heatmaps.to(self.device, torch.float32, non_blocking=True).contiguous()
        key = (n, j, h, w, train)
        buf = self._buffers.get(key)
        if buf is None:   # outputs are allocated once per batch shape and reused
            buf = {"coords": torch.empty((n, j, 2), device=self.device), "maxvals": torch.empty((n, j, 1), device=self.device),
                   "argidx": torch.empty((n, j), dtype=torch.int32, device=self.device),
                   "loss": torch.empty((), device=self.device), "weight": torch.empty((n, j, 1), device=self.device)}
            if train:
                buf["grad"] = torch.empty((n, j, h, w), device=self.device)
            self._buffers[key] = buf
        stride = float(self.multiply)
        return ops.heatmap_fused(pred.detach().float().contiguous(), joints.contiguous(), target_in=target, sigma=self.sigma,
                                 feat_stride=(stride, stride), want_grad=train, want_decode=True, out_scale=stride,
                                 loss_meter=self.meter.loss_meter,
                                 metrics=None if train else (self.meter.counts, self.meter.sums), pck_T=0.2,
                                 buffers=buf)

    # ------------------------------------------------------------------------------------------
    def run(self, end):
        if self.phase == "TRAIN":
            self.model.train()
            self.meter.reset()
            for iteration, (images, gt_2d_joints, gt_heatmaps) in enumerate(self.train_loader):
                images = images.to(self.device, non_blocking=True)
                pred = self.model(images)
                out = self._fused(pred, gt_2d_joints, gt_heatmaps, train=True)      # loss, d loss / d pred, joints * 4
                self.pred_joint = out["coords"]
                self.optimizer.zero_grad()
                pred.backward(out["grad"])            # == loss.backward() of the reference (src/utils/method.py:182)
                self.optimizer.step()
                self.batch_time.update(time.time() - end)
                end = time.time()
                eta_seconds = self.batch_time.avg * ((self.len_total - iteration)
                                                     + (getattr(self.args, "epoch", 1) - self.epoch - 1) * self.len_total)
                self._log(iteration, eta_seconds, end)
            self._sync()
            if self.writer is not None:
                self.writer.add_scalar("Loss/train", self.log_losses.avg, self.epoch)
            return self.model, self.optimizer, self.batch_time

        self.model.eval()
        self.meter.reset()
        with torch.no_grad():
            for iteration, (images, gt_2d_joints, gt_heatmaps) in enumerate(self.now_loader):
                images = images.to(self.device, non_blocking=True)
                pred = self.model(images)
                out = self._fused(pred, gt_2d_joints, gt_heatmaps, train=False)     # + PCK@0.2 / EPE counts
                self.pred_joint = out["coords"]
                self.batch_time.update(time.time() - end)
                end = time.time()
                eta_seconds = self.batch_time.avg * ((len(self.now_loader) - iteration)
                                                     + (getattr(self.args, "epoch", 1) - self.epoch - 1) * self.len_total)
                self._log(iteration, eta_seconds, end)
        self._sync()
        if self.writer is not None:
            self.writer.add_scalar("Loss/valid", self.log_losses.avg, self.epoch)
        return self.log_losses.avg, self.count, self.pck_losses.avg * 100, self.batch_time


class Runner_v(Runner_t):
    def __init__(self, train_runner, model, test_dataloader, phase, best_loss):
        self.__dict__ = train_runner.__dict__.copy()
        self.now_loader = test_dataloader
        self.best_loss = best_loss
        self.model = model
        self.phase = phase
        self.meter = DeviceMeter(self.device)
        self.log_losses = _MeterView()
        self.pck_losses = _MeterView()
        self.epe_losses = _MeterView()
        self._buffers = {}
