"""FreiHAND crop with the reference's call surface (src/utils/image_ops.py:59-131).

``get_transform`` / ``transform`` are the reference's host arithmetic (float64, ``torch.inverse`` for the
inverse, truncation toward zero).  ``crop`` keeps its signature; the window bookkeeping of the reference
(``ul``, ``br``, rotation padding, :99-106) stays on the host -- a few integers per sample -- and is handed to the
device as ``lh_frei_desc`` records; the resampling (window copy, rotation about the canvas centre on cv2's
fixed-point grid, ``cv2.resize``) runs batched in ``lh_frei_rgb_processing``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib, ops
from ..ops import LighthandError


def get_transform(center, scale, res, rot=0):
    """src/utils/image_ops.py:59-83: the 3x3 map from image pixels to the res x res crop of a box of side
    200 * scale around ``center``, optionally rotated about the crop centre.  Same float64 operations in the same
    order as the reference (the integer window corners of ``crop`` depend on the last bit)."""
    side = 200 * scale
    t = np.array([[float(res[1]) / side, 0.0, res[1] * (-float(center[0]) / side + .5)],
                  [0.0, float(res[0]) / side, res[0] * (-float(center[1]) / side + .5)],
                  [0.0, 0.0, 1.0]])
    if rot == 0:
        return t
    angle = -rot * np.pi / 180            # the crop rotates against the image
    sn, cs = np.sin(angle), np.cos(angle)
    spin = np.array([[cs, -sn, 0.0], [sn, cs, 0.0], [0.0, 0.0, 1.0]])
    to_centre, back = np.eye(3), np.eye(3)
    to_centre[0, 2], to_centre[1, 2] = -res[1] / 2, -res[0] / 2
    back[0, 2], back[1, 2] = -to_centre[0, 2], -to_centre[1, 2]
    return np.dot(back, np.dot(spin, np.dot(to_centre, t)))


def transform(pt, center, scale, res, invert=0, rot=0):
    """src/utils/image_ops.py:85-95."""
    t = get_transform(center, scale, res, rot=rot)
    if invert:
        t = torch.inverse(torch.from_numpy(t)).numpy()
    new_pt = np.array([pt[0] - 1, pt[1] - 1, 1.]).T
    new_pt = np.dot(t, new_pt)
    return new_pt[:2].astype(int) + 1


def rotation_matrix(center, angle, scale=1.0):
    """cv2.getRotationMatrix2D (called by myimrotate, src/utils/image_ops.py:33): float32 centre, double
    trigonometry."""
    cx, cy = float(np.float32(center[0])), float(np.float32(center[1]))
    a = angle * np.pi / 180
    alpha, beta = np.cos(a) * scale, np.sin(a) * scale
    return np.array([alpha, beta, (1 - alpha) * cx - beta * cy, -beta, alpha, beta * cx + (1 - alpha) * cy])


def invert_affine_cv2(m):
    """The double-precision inversion cv2.warpAffine applies to a forward 2x3 matrix (OpenCV imgwarp.cpp;
    SURVEY.md Appendix A step 1), operation for operation."""
    m0, m1, m2, m3, m4, m5 = [float(v) for v in m]
    D = m0 * m4 - m1 * m3
    D = 1.0 / D if D != 0 else 0.0
    a11, a22 = m4 * D, m0 * D
    m0 = a11
    m1 = m1 * (-D)
    m3 = m3 * (-D)
    m4 = a22
    b1 = -m0 * m2 - m1 * m5
    b2 = -m3 * m2 - m4 * m5
    return [m0, m1, b1, m3, m4, b2]


def crop_descriptors(centers, scales, res, rots, pns=None):
    """Host half of ``crop`` for a batch (src/utils/image_ops.py:99-106): returns (records, scratch_floats,
    (max_rot_w, max_rot_h)) where ``records`` is a ctypes array of ``lh_frei_desc``."""
    n = len(scales)
    recs = (_lib.FreiDesc * n)()
    off, max_w, max_h = 0, 0, 0
    for i in range(n):
        center, scale, rot = centers[i], float(scales[i]), float(rots[i])
        ul = np.array(transform([1, 1], center, scale, res, invert=1)) - 1
        br = np.array(transform([res[0] + 1, res[1] + 1], center, scale, res, invert=1)) - 1
        pad = int(np.linalg.norm(br - ul) / 2 - float(br[1] - ul[1]) / 2)
        r = recs[i]
        if not rot == 0:
            ul = ul - pad
            br = br + pad
            if pad <= 0:
                raise LighthandError("crop: rotation with a non-positive padding (the reference slices [pad:-pad])")
        r.ulx, r.uly = int(ul[0]), int(ul[1])
        r.cw, r.ch = int(br[0] - ul[0]), int(br[1] - ul[1])
        r.pad, r.rotated = int(pad), int(not rot == 0)
        if r.cw <= 0 or r.ch <= 0:
            raise LighthandError("crop: empty window")
        iw, ih = r.cw, r.ch
        if r.rotated:
            im = invert_affine_cv2(rotation_matrix(((r.cw - 1) * 0.5, (r.ch - 1) * 0.5), rot))
            for k in range(6):
                r.iM[k] = im[k]
            iw, ih = r.cw - 2 * pad, r.ch - 2 * pad
            if iw <= 0 or ih <= 0:
                raise LighthandError("crop: nothing left after removing the rotation padding")
            r.scratch_off = off
            off += 3 * iw * ih
            max_w, max_h = max(max_w, iw), max(max_h, ih)
        r.scale_x = 1.0 / (float(res[0]) / float(iw))      # cv2.resize: scale = 1 / inv_scale, inv_scale = dsize / ssize
        r.scale_y = 1.0 / (float(res[1]) / float(ih))
        pn = (1.0, 1.0, 1.0) if pns is None else pns[i]
        r.pn[0], r.pn[1], r.pn[2] = float(pn[0]), float(pn[1]), float(pn[2])
    return recs, off, (max_w, max_h)


def descriptors_to_device(recs, device) -> torch.Tensor:
    raw = np.frombuffer(bytes(recs), dtype=np.uint8).copy() if len(recs) else np.zeros(0, np.uint8)
    return torch.from_numpy(raw).to(device)


def crop_batch(images_u8: torch.Tensor, centers, scales, res, rots, pns=None) -> torch.Tensor:
    """``crop`` + pixel noise + /255 for a batch of equally sized uint8 images on the device:
    [N,H,W,3] -> float32 [N,3,res,res] in 0..1 (rgb_processing's result, frei_dataloader.py:134-147)."""
    if res[0] != res[1]:
        raise LighthandError("crop: square output only (the reference passes [img_res, img_res])")
    recs, scratch_floats, max_wh = crop_descriptors(centers, scales, res, rots, pns)
    desc = descriptors_to_device(recs, images_u8.device)
    rot = torch.tensor([i for i, r in enumerate(recs) if r.rotated], dtype=torch.int32).to(images_u8.device)
    return ops.frei_rgb_processing(images_u8, desc, int(res[0]), scratch_floats, max_wh, rot_index=rot)


def crop(img, center, scale, res, rot=0):
    """src/utils/image_ops.py:97-131 for one image: HWC uint8 (numpy or CUDA tensor) -> CUDA float32
    [res, res, 3] on the 0..255 scale (the reference returns float64 numpy)."""
    t = img if isinstance(img, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(img))
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
        raise LighthandError("crop: expected an HWC uint8 image")
    ops.require_cuda()
    out = crop_batch(t.cuda()[None], [center], [scale], res, [rot])
    return (out[0] * 255.0).permute(1, 2, 0).contiguous()
