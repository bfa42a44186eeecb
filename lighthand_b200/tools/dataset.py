"""Target generation and batch collation with the reference's call surface (src/tools/dataset.py).

``generate_target`` replaces CustomDataset.generate_target (src/tools/dataset.py:165-212 and its three
verbatim copies); ``GpuCollate`` replaces ``CustomDataset.__getitem__`` + default collate
(src/tools/dataset.py:125-163): DataLoader workers ship uint8 HWC images + keypoints, the
ToTensor -> Resize -> Normalize and the Gaussian targets run batched on the device.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from ..utils.loss import _to_cuda


def generate_target(joints, num_joints=21, heatmap_size=(64, 64), feat_stride=(4, 4), sigma=2,
                    return_weight=False):
    """One sample [J,2] (or a batch [N,J,2]) of input-pixel joints -> float32 target [J,H,W].
    Defaults are the reference's hard-coded geometry.  CPU input -> CPU tensor (like the
    reference's ``torch.tensor(target)``), CUDA input -> CUDA tensor."""
    on_dev = isinstance(joints, torch.Tensor) and joints.is_cuda
    t = _to_cuda(joints)
    single = t.dim() == 2
    if single:
        t = t[None]
    if t.shape[1] != num_joints:
        raise ops.LighthandError("generate_target: wrong number of joints")
    target, weight = ops.gen_targets(t, heatmap_size, (float(feat_stride[0]), float(feat_stride[1])), sigma,
                                     style=ops.LH_TARGET_HRNET, want_weight=True)
    if single:
        target, weight = target[0], weight[0]
    if not on_dev:
        target, weight = target.cpu(), weight.cpu()
    return (target, weight) if return_weight else target


def color_jitter_params(apply_flags, brightness=0.5, contrast=0.5, saturation=0.5, hue=0.5):
    """Host half of ``transforms.ColorJitter(brightness, contrast, saturation, hue)`` (src/tools/dataset.py:139-141)
    for a batch: for every sample with a true flag, the draw torchvision's ``ColorJitter.get_params`` makes --
    ``torch.randperm(4)``, then one ``uniform_`` each for brightness, contrast, saturation, hue, from torch's default
    generator -- so a seeded run reproduces the reference's stream sample by sample.  Returns a ctypes array of
    ``lh_jitter_params``."""
    from .. import _lib
    n = len(apply_flags)
    recs = (_lib.JitterParams * n)()
    ranges = [(max(0.0, 1 - brightness), 1 + brightness), (max(0.0, 1 - contrast), 1 + contrast),
              (max(0.0, 1 - saturation), 1 + saturation), (-hue, hue)]
    for i, flag in enumerate(apply_flags):
        r = recs[i]
        r.apply = int(bool(flag))
        for k in range(4):
            r.order[k] = -1
        if not flag:
            continue
        fn_idx = torch.randperm(4)
        fac = [float(torch.empty(1).uniform_(lo, hi)) for lo, hi in ranges]
        for k in range(4):
            r.order[k] = int(fn_idx[k])
            r.factor[k] = fac[k]
            r.one_minus[k] = 1.0 - fac[k]
    return recs


class GpuCollate:
    """collate_fn for a DataLoader whose dataset yields ``(image_u8_hwc, joint_2d)``:
    returns ``(images f32[N,3,S,S], joint_2d f32[N,J,2], heatmaps f32[N,J,H,W])`` on ``device`` --
    the triple the reference's loops consume (src/utils/method.py:160).  Keypoints are scaled by
    S / S_src (src/tools/dataset.py:132).  A dataset that yields a third element per sample,
    ``(image, joint_2d, augment)``, gets ``ColorJitter(0.5, 0.5, 0.5, 0.5)`` on the samples whose flag is true
    (the reference's ``idx < len(meta) * ratio_of_aug``, src/tools/dataset.py:134-146), drawn per sample from
    torch's generator in torchvision's order.  ``phase="eval"``: the Armo evaluation set's item shape, see ``_eval_batch``."""

    def __init__(self, image_size=256, heatmap_size=64, sigma=2, device="cuda",
                 mean=ops.IMAGENET_MEAN, std=ops.IMAGENET_STD, jitter=(0.5, 0.5, 0.5, 0.5), phase="train"):
        self.image_size, self.heatmap_size, self.sigma = image_size, heatmap_size, sigma
        self.device, self.mean, self.std = torch.device(device), mean, std
        self.jitter = jitter
        self.phase = phase

    def _eval_batch(self, samples):
        """``eval_set.__getitem__`` with ``phase == "eval"`` (src/tools/dataset.py:257-300) + default collate: samples are
        ``(image_u8_hwc, joint [21, >=2] normalised to 0..1, visible [21], pose_type, idx)``; returns what ``pred_store``
        iterates over (src/utils/argparser.py:255-257): ``(images f32[N,3,S,S], joint_2d_v f32[N,21,3] = x*S, y*S, vis,
        [pose_types, idx])``.  ToTensor -> Resize -> Normalize runs batched on the device."""
        imgs = np.stack([np.asarray(s[0], dtype=np.uint8) for s in samples])
        joint = torch.stack([torch.as_tensor(s[1], dtype=torch.float32)[:, :2] for s in samples])
        vis = torch.stack([torch.as_tensor(s[2], dtype=torch.float32) for s in samples])
        joint_2d_v = torch.cat([joint * self.image_size, vis[:, :, None]], dim=2)      # :286-289
        d_imgs = torch.from_numpy(imgs).to(self.device, non_blocking=True)
        images = ops.warp_normalize(d_imgs, None, (self.image_size, self.image_size), mean=self.mean, std=self.std,
                                    interp=ops.LH_INTERP_TORCH)
        pose = [s[3] for s in samples]
        idx = [s[4] for s in samples]
        return images, joint_2d_v.to(self.device, non_blocking=True), [pose, idx]

    def __call__(self, samples):
        if self.phase == "eval":
            return self._eval_batch(samples)
        imgs = np.stack([np.asarray(s[0], dtype=np.uint8) for s in samples])
        joints = torch.stack([torch.as_tensor(s[1], dtype=torch.float32)[:, :2] for s in samples])
        src = imgs.shape[1]
        joints = joints * (self.image_size / src)
        d_imgs = torch.from_numpy(imgs).to(self.device, non_blocking=True)
        d_joints = joints.to(self.device, non_blocking=True)
        flags = [bool(s[2]) for s in samples] if samples and len(samples[0]) > 2 else None
        if flags is not None and any(flags):
            # ToTensor -> Resize, then ColorJitter -> Normalize in one more pass
            images = ops.warp_normalize(d_imgs, None, (self.image_size, self.image_size), mean=None, std=None,
                                        interp=ops.LH_INTERP_TORCH)
            images = ops.color_jitter_normalize(images, color_jitter_params(flags, *self.jitter), self.mean, self.std,
                                                out=images)
        else:
            images = ops.warp_normalize(d_imgs, None, (self.image_size, self.image_size), mean=self.mean,
                                        std=self.std, interp=ops.LH_INTERP_TORCH)
        stride = self.image_size / self.heatmap_size
        heat, _ = ops.gen_targets(d_joints, (self.heatmap_size, self.heatmap_size), (stride, stride),
                                  self.sigma, want_weight=False)
        return images, d_joints, heat
