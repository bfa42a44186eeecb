"""Single-warp crop with the reference's call surface (src/utils/preprocessing.py:144-210).

The 2x3 matrices are built on the host, vectorised over the batch, mirroring the reference's
float32/float64 steps; the warp itself (cv2.warpAffine's fixed-point bilinear grid) runs in K1 on
the device.
"""
from __future__ import annotations

import random

import numpy as np
import torch

from .. import ops
from ..ops import LighthandError


def rotate_2d(pt_2d, rot_rad):
    """src/utils/preprocessing.py:164-170."""
    x, y = pt_2d[0], pt_2d[1]
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    return np.array([x * cs - y * sn, x * sn + y * cs], dtype=np.float32)


def _affine_from_3pts(src, dst):
    """The affine map through three point pairs, batched: src,dst [N,3,2] -> float64 [N,2,3].

    Same arithmetic as cv2.getAffineTransform, which the reference calls
    (src/utils/preprocessing.py:200-202): the 6x6 system [x y 1 0 0 0; 0 0 0 x y 1] solved in double by
    LU with partial pivoting (first largest pivot), elimination with alpha = a_ji * (-1/a_ii), back
    substitution dividing by the pivot.  Vectorised over the batch; bit-identical to cv2 4.13 on
    the tested inputs (tests/test_host_logic.py)."""
    src = np.asarray(src).astype(np.float64)
    dst = np.asarray(dst).astype(np.float64)
    n = src.shape[0]
    A = np.zeros((n, 6, 6))
    b = np.zeros((n, 6))
    for i in range(3):
        A[:, 2 * i, 0] = src[:, i, 0]; A[:, 2 * i, 1] = src[:, i, 1]; A[:, 2 * i, 2] = 1
        A[:, 2 * i + 1, 3] = src[:, i, 0]; A[:, 2 * i + 1, 4] = src[:, i, 1]; A[:, 2 * i + 1, 5] = 1
        b[:, 2 * i] = dst[:, i, 0]; b[:, 2 * i + 1] = dst[:, i, 1]
    ar = np.arange(n)
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(6):
            k = i + np.argmax(np.abs(A[:, i:, i]), axis=1)
            Ai, Ak = A[ar, i, :].copy(), A[ar, k, :].copy()
            A[ar, i, :], A[ar, k, :] = Ak, Ai
            bi, bk = b[ar, i].copy(), b[ar, k].copy()
            b[ar, i], b[ar, k] = bk, bi
            d = -1.0 / A[:, i, i]
            for j in range(i + 1, 6):
                alpha = A[:, j, i] * d
                A[:, j, i + 1:] += alpha[:, None] * A[:, i, i + 1:]
                b[:, j] += alpha * b[:, i]
        for i in range(5, -1, -1):
            s_ = b[:, i].copy()
            for k in range(i + 1, 6):
                s_ -= A[:, i, k] * b[:, k]
            b[:, i] = s_ / A[:, i, i]
    return b.reshape(n, 2, 3)


def gen_trans_from_patch_batch(c_x, c_y, src_width, src_height, dst_width, dst_height, scale, rot, inv=False):
    """Batched gen_trans_from_patch_cv (src/utils/preprocessing.py:172-205): every argument may be a
    scalar or an [N] array; returns float32 [N,2,3]."""
    c_x, c_y, src_width, src_height, scale, rot = np.broadcast_arrays(
        *[np.asarray(v, dtype=np.float64) for v in (c_x, c_y, src_width, src_height, scale, rot)])
    n = c_x.shape[0] if c_x.ndim else 1
    c_x, c_y, src_width, src_height, scale, rot = [np.reshape(v, (n,)) for v in (c_x, c_y, src_width, src_height, scale, rot)]
    f32 = np.float32
    src_w, src_h = src_width * scale, src_height * scale
    centre = np.stack([c_x, c_y], -1).astype(f32)
    rad = np.pi * rot / 180
    sn, cs = np.sin(rad), np.cos(rad)
    zero = np.zeros(n, f32)
    hy = (src_h * 0.5).astype(f32).astype(np.float64)   # np.array([0, src_h*0.5], float32)
    hx = (src_w * 0.5).astype(f32).astype(np.float64)
    down = np.stack([zero * cs - hy * sn, zero * sn + hy * cs], -1).astype(f32)
    right = np.stack([hx * cs - zero * sn, hx * sn + zero * cs], -1).astype(f32)
    dcentre = np.array([dst_width * 0.5, dst_height * 0.5], dtype=f32)
    ddown = np.array([0, dst_height * 0.5], dtype=f32)
    dright = np.array([dst_width * 0.5, 0], dtype=f32)
    src = np.stack([centre, centre + down, centre + right], 1).astype(f32)
    dst = np.broadcast_to(np.stack([dcentre, dcentre + ddown, dcentre + dright])[None], (n, 3, 2)).astype(f32)
    m = _affine_from_3pts(dst, src) if inv else _affine_from_3pts(src, dst)
    return m.astype(f32)


def gen_trans_from_patch_cv(c_x, c_y, src_width, src_height, dst_width, dst_height, scale, rot, inv=False):
    """src/utils/preprocessing.py:172-205 -> float32 [2,3]."""
    return gen_trans_from_patch_batch(c_x, c_y, src_width, src_height, dst_width, dst_height, scale, rot, inv)[0]


def trans_point2d(pt_2d, trans):
    """src/utils/preprocessing.py:207-210."""
    src_pt = np.array([pt_2d[0], pt_2d[1], 1.]).T
    return np.dot(trans, src_pt)[0:2]


def get_aug_config():
    """src/utils/preprocessing.py:48-63 (same RNG streams: numpy then python ``random``)."""
    trans_factor, scale_factor, rot_factor, color_factor = 0.15, 0.25, 45, 0.2
    trans = [np.random.uniform(-trans_factor, trans_factor), np.random.uniform(-trans_factor, trans_factor)]
    scale = np.clip(np.random.randn(), -1.0, 1.0) * scale_factor + 1.0
    rot = np.clip(np.random.randn(), -2.0, 2.0) * rot_factor if random.random() <= 0.6 else 0
    do_flip = random.random() <= 0.5
    c_up, c_low = 1.0 + color_factor, 1.0 - color_factor
    color_scale = np.array([random.uniform(c_low, c_up), random.uniform(c_low, c_up), random.uniform(c_low, c_up)])
    return trans, scale, rot, do_flip, color_scale


def _as_u8_source(cvimg):
    """uint8 HWC passes through; float images must hold exact 0..255 integers (what load_img yields,
    src/utils/preprocessing.py:13-24) and then select cv2's float-weight path."""
    a = np.asarray(cvimg)
    if a.ndim != 3 or a.shape[2] != 3:
        raise LighthandError("generate_patch_image: expected an HxWx3 image")
    if a.dtype == np.uint8:
        return np.ascontiguousarray(a), ops.LH_INTERP_CV2_U8
    u8 = a.astype(np.uint8)
    if not np.array_equal(u8.astype(a.dtype), a):
        raise LighthandError("generate_patch_image: float images must contain integer values 0..255 "
                             "(the device path samples uint8 sources)")
    return np.ascontiguousarray(u8), ops.LH_INTERP_CV2_F32


def generate_patch_image(cvimg, bbox, do_flip, scale, rot, out_shape):
    """src/utils/preprocessing.py:144-162 -> (img_patch float32 HWC 0..255, trans, inv_trans)."""
    src, interp = _as_u8_source(cvimg)
    ih, iw = src.shape[:2]
    cx = float(bbox[0] + 0.5 * bbox[2])
    cy = float(bbox[1] + 0.5 * bbox[3])
    bw, bh = float(bbox[2]), float(bbox[3])
    if do_flip:
        cx = iw - cx - 1
    trans = gen_trans_from_patch_cv(cx, cy, bw, bh, out_shape[1], out_shape[0], scale, rot)
    inv_trans = gen_trans_from_patch_cv(cx, cy, bw, bh, out_shape[1], out_shape[0], scale, rot, inv=True)
    ops.require_cuda()
    d_img = torch.from_numpy(src)[None].cuda()
    d_m = torch.from_numpy(trans.astype(np.float64))[None].cuda()
    d_flip = torch.tensor([1 if do_flip else 0], dtype=torch.uint8, device="cuda")
    patch = ops.warp_normalize(d_img, d_m, (int(out_shape[0]), int(out_shape[1])), flip=d_flip, mean=None,
                               std=None, interp=interp, layout=ops.LH_LAYOUT_HWC)
    return patch[0].cpu().numpy(), trans, inv_trans


def batch_patch_matrices(bbox, do_flip, scale, rot, src_hw, out_hw):
    """Batched front half of generate_patch_image: bbox [N,4] (x,y,w,h), flip/scale/rot [N] ->
    (trans f32[N,2,3], inv_trans f32[N,2,3]).  Feed ``trans.astype(float64)`` to ops.warp_normalize."""
    bbox = np.asarray(bbox, dtype=np.float64)
    do_flip = np.asarray(do_flip).astype(bool)
    cx = bbox[:, 0] + 0.5 * bbox[:, 2]
    cy = bbox[:, 1] + 0.5 * bbox[:, 3]
    cx = np.where(do_flip, src_hw[1] - cx - 1, cx)
    args = (cx, cy, bbox[:, 2], bbox[:, 3], out_hw[1], out_hw[0], scale, rot)
    return gen_trans_from_patch_batch(*args), gen_trans_from_patch_batch(*args, inv=True)
