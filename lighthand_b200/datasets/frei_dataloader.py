"""Device-side tail of HandMeshTSVDataset.__getitem__ (src/datasets/frei_dataloader.py:105-160, 255-338).

The TSV / JSON annotation store and the SMPL / 3-D branches of the reference dataset are outside the heat-map
path; what is mirrored here is what produces the triple the training loop consumes,

    transfromed_img[(2,1,0)], joint_2d, heatmap          (frei_dataloader.py:338)

for a whole batch: ``rgb_processing`` (crop, noise, /255), ``Resize((256,256))``, ``Normalize``, the channel
reversal, ``j2d_processing`` and ``generate_target``.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from ..utils import image_ops

IMG_RES = 224        # frei_dataloader.py:75
OUT_SIZE = 256       # frei_dataloader.py:279
NOISE_FACTOR, ROT_FACTOR, SCALE_FACTOR = 0.4, 90, 0.25     # frei_dataloader.py:72-74


def augm_params(is_train: bool, rng=np.random):
    """src/datasets/frei_dataloader.py:105-132 (same draw order, so a seeded ``np.random`` reproduces it)."""
    flip, pn, rot, sc = 0, np.ones(3), 0, 1.0
    if is_train:
        pn = rng.uniform(1 - NOISE_FACTOR, 1 + NOISE_FACTOR, 3)
        rot = min(2 * ROT_FACTOR, max(-2 * ROT_FACTOR, rng.randn() * ROT_FACTOR))
        sc = min(1 + SCALE_FACTOR, max(1 - SCALE_FACTOR, rng.randn() * SCALE_FACTOR + 1))
        if rng.uniform() <= 0.6:
            rot = 0
    return flip, pn, rot, sc


def rgb_processing(images_u8: torch.Tensor, centers, scales, rots, pns) -> torch.Tensor:
    """frei_dataloader.py:134-147 for a batch (flip is always 0, :107): [N,H,W,3] uint8 CUDA -> [N,3,224,224]."""
    return image_ops.crop_batch(images_u8, centers, scales, [IMG_RES, IMG_RES], rots, pns)


def j2d_processing(kp, center, scale, r, f=0):
    """frei_dataloader.py:148-160 (host; 21 points through ``transform`` with its integer truncation)."""
    kp = np.array(kp, dtype=np.float64)
    for i in range(kp.shape[0]):
        kp[i, 0:2] = image_ops.transform(kp[i, 0:2] + 1, center, scale, [IMG_RES, IMG_RES], rot=r)
    kp[:, :-1] = 2. * kp[:, :-1] / IMG_RES - 1.
    if f:
        raise NotImplementedError("flip is always 0 in the reference (frei_dataloader.py:107)")
    return kp.astype("float32")


def frei_batch(images_u8: torch.Tensor, joints_2d, centers, scales, augm=None, want_heatmap=True):
    """The (image, joint_2d, heatmap) triple of ``__getitem__`` for a batch.

    images_u8 [N,H,W,3] uint8 CUDA (as decoded, channel order untouched); joints_2d [N,21,3] annotation
    keypoints; centers [N,2]; scales [N]; augm = list of (flip, pn, rot, sc) per sample (``augm_params``),
    default: the eval parameters.  Returns CUDA tensors ([N,3,256,256], [N,21,2], [N,21,64,64])."""
    n = images_u8.shape[0]
    if augm is None:
        augm = [augm_params(False) for _ in range(n)]
    pns = [a[1] for a in augm]
    rots = [a[2] for a in augm]
    sc_scales = [a[3] * float(scales[i]) for i, a in enumerate(augm)]
    # rgb_processing -> Resize((256, 256)) -> Normalize -> [(2, 1, 0)] in one pass (lh_frei_fused: the 224 x 224 float image
    # stays in shared memory); bit-identical to rgb_processing(...) followed by ops.resize_normalize_f32(...)
    recs, scratch_floats, max_wh = image_ops.crop_descriptors(centers, sc_scales, [IMG_RES, IMG_RES], rots, pns)
    desc = image_ops.descriptors_to_device(recs, images_u8.device)
    rot_idx = torch.tensor([i for i, r in enumerate(recs) if r.rotated], dtype=torch.int32).to(images_u8.device)
    img = ops.frei_fused(images_u8, desc, IMG_RES, scratch_floats, max_wh, (OUT_SIZE, OUT_SIZE), chan_order=ops.LH_CHAN_BGR,
                         rot_index=rot_idx)
    j2d = np.stack([j2d_processing(np.asarray(joints_2d[i]).copy(), centers[i], sc_scales[i], rots[i], augm[i][0])
                    for i in range(n)]) if n else np.zeros((0, 21, 3), np.float32)
    joint_2d = (torch.from_numpy(j2d).float()[:, :, :-1] * 100 + 112) * (OUT_SIZE / IMG_RES)      # :335
    joint_2d = joint_2d.to(images_u8.device)
    heatmap = ops.gen_targets(joint_2d.contiguous(), (64, 64), (4.0, 4.0), 2)[0] if want_heatmap else None
    return img, joint_2d, heatmap
