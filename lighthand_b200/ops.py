"""Batched device ops: thin torch-tensor wrappers over the C ABI (one call = one kernel launch,
asynchronous on torch's current stream).  torch is used for device memory and streams only.

No CPU path exists: every function raises unless its tensors live on a CUDA device and the
sm_100a library is built.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import (LH_CHAN_BGR, LH_CHAN_RGB, LH_INTERP_CV2_F32, LH_INTERP_CV2_U8, LH_INTERP_TORCH,
                   LH_LAYOUT_CHW, LH_LAYOUT_HWC, LH_PCK_MM, LH_PCK_PROPORTION, LH_TARGET_HOURGLASS,
                   LH_TARGET_HRNET, LighthandError, check)

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise LighthandError("lighthand_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")


def _dev(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise LighthandError(f"{name}: expected a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise LighthandError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _mean_std(mean, std):
    if mean is None:
        return None
    arr = (C.c_float * 6)(*[float(np.float32(v)) for v in (*mean, *std)])
    return arr


# ---------------------------------------------------------------------------------------------
# Gaussian tables, computed on the host exactly as the reference computes them (bit-exact targets)
# ---------------------------------------------------------------------------------------------
_table_cache = {}


def hrnet_table(sigma: float, device) -> Tuple[torch.Tensor, int]:
    """float32 numpy exp table of src/tools/dataset.py:189-195 for tmp_size = 3*sigma."""
    key = ("hrnet", float(sigma), str(device))
    if key not in _table_cache:
        tmp = int(sigma * 3)
        size = 2 * tmp + 1
        x = np.arange(0, size, 1, np.float32)
        y = x[:, None]
        c = size // 2
        g = np.exp(-((x - c) ** 2 + (y - c) ** 2) / (2 * sigma ** 2)).astype(np.float32)
        _table_cache[key] = (torch.from_numpy(g).to(device).contiguous(), size)
    return _table_cache[key]


def hourglass_table(output_res: int, device) -> Tuple[torch.Tensor, int]:
    """float64 table of src/utils/dataset_loader.py:23-32 rounded to float32 (what the paste stores)."""
    key = ("hg", int(output_res), str(device))
    if key not in _table_cache:
        if output_res % 64 != 0:
            raise LighthandError("GenerateHeatmap: output_res must be a multiple of 64 (integer sigma)")
        sigma = output_res / 64
        size = 6 * sigma + 3
        x = np.arange(0, size, 1, float)
        y = x[:, None]
        c = 3 * sigma + 1
        g = np.exp(-((x - c) ** 2 + (y - c) ** 2) / (2 * sigma ** 2)).astype(np.float32)
        _table_cache[key] = (torch.from_numpy(g).to(device).contiguous(), int(size))
    return _table_cache[key]


# ---------------------------------------------------------------------------------------------
# K1
# ---------------------------------------------------------------------------------------------

def build_norm_lut(mean=IMAGENET_MEAN, std=IMAGENET_STD, chan_gain: Optional[torch.Tensor] = None,
                   device=None) -> torch.Tensor:
    """[T,3,256] float32 table: ((clip(v*gain,0,255))/255 - mean)/std; T = N with gain else 1.
    mean=None gives the identity table (0..255 floats)."""
    require_cuda()
    lib = _lib.load()
    if chan_gain is not None:
        chan_gain = _dev(chan_gain, torch.float64, "chan_gain")
        n = chan_gain.shape[0]
        device = chan_gain.device
    else:
        n = 1
        device = torch.device(device if device is not None else "cuda")
    lut = torch.empty((n, 3, 256), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        check(lib.lh_norm_lut_build(_ptr(chan_gain), n, _mean_std(mean, std), lut.data_ptr(), _stream()),
              "lh_norm_lut_build")
    return lut


def warp_normalize(images_u8: torch.Tensor, fwd_affine: Optional[torch.Tensor], out_hw: Sequence[int],
                   flip: Optional[torch.Tensor] = None, chan_gain: Optional[torch.Tensor] = None,
                   mean=IMAGENET_MEAN, std=IMAGENET_STD, lut: Optional[torch.Tensor] = None,
                   interp: int = LH_INTERP_CV2_U8, chan_order: int = LH_CHAN_RGB,
                   layout: int = LH_LAYOUT_CHW, out: Optional[torch.Tensor] = None,
                   exact_norm: bool = True) -> torch.Tensor:
    """K1.  images_u8 [N,Hs,Ws,3] uint8 CUDA; fwd_affine [N,2,3] float64 (src->dst, as cv2.warpAffine
    takes it); returns float32 [N,3,oh,ow] (CHW) or [N,oh,ow,3] (HWC).

    The warped uint8 pixel is always cv2's, bit for bit.  ``exact_norm=True`` maps it through a table
    holding torch's float32 ``((v/255) - mean)/std`` (three roundings, bit-identical to
    ToTensor+Normalize); ``exact_norm=False`` uses one fused multiply-add ``v*a + b`` (within 2 ulp of the
    former, i.e. ~1e-7 relative against the path's 1e-5 tolerance) and needs no table."""
    require_cuda()
    lib = _lib.load()
    images_u8 = _dev(images_u8, torch.uint8, "images")
    if images_u8.dim() != 4 or images_u8.shape[3] != 3:
        raise LighthandError("images: expected [N,H,W,3] uint8")
    n, sh, sw, _ = images_u8.shape
    oh, ow = int(out_hw[0]), int(out_hw[1])
    if fwd_affine is not None:
        fwd_affine = _dev(fwd_affine, torch.float64, "fwd_affine")
        if fwd_affine.numel() != n * 6:
            raise LighthandError("fwd_affine: expected [N,2,3]")
    if flip is not None:
        flip = _dev(flip, torch.uint8, "flip")
    if chan_gain is not None:
        chan_gain = _dev(chan_gain, torch.float64, "chan_gain")
    if interp == LH_INTERP_CV2_U8 and lut is None and (exact_norm or chan_gain is not None) and mean is not None:
        lut = build_norm_lut(mean, std, chan_gain, images_u8.device)
    if interp == LH_INTERP_CV2_U8 and lut is None and chan_gain is not None:
        lut = build_norm_lut(None, None, chan_gain, images_u8.device)
    if lut is not None:
        lut = _dev(lut, torch.float32, "lut")
    shape = (n, 3, oh, ow) if layout == LH_LAYOUT_CHW else (n, oh, ow, 3)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=images_u8.device)
    else:
        out = _dev(out, torch.float32, "out")
        if tuple(out.shape) != shape:
            raise LighthandError("out: wrong shape")
    if n == 0:
        return out
    with torch.cuda.device(images_u8.device):
        check(lib.lh_warp_normalize(images_u8.data_ptr(), sh * sw * 3, sh, sw, sw * 3, _ptr(fwd_affine), _ptr(flip),
                                    _ptr(chan_gain), _mean_std(mean, std), _ptr(lut), interp, chan_order, layout,
                                    out.data_ptr(), n, oh, ow, _stream()), "lh_warp_normalize")
    return out


# ---------------------------------------------------------------------------------------------
# FreiHAND image path (a-3)
# ---------------------------------------------------------------------------------------------

def frei_rgb_processing(images_u8: torch.Tensor, desc: torch.Tensor, res: int, scratch_floats: int,
                        max_rot_wh: Sequence[int], out: Optional[torch.Tensor] = None,
                        rot_index: Optional[torch.Tensor] = None) -> torch.Tensor:
    """rgb_processing (src/datasets/frei_dataloader.py:134-147) of a batch.  images_u8 [N,H,W,3] uint8 CUDA;
    ``desc`` = uint8 CUDA tensor holding N packed ``lh_frei_desc`` records (utils.image_ops.crop_descriptors);
    returns float32 [N,3,res,res] in 0..1."""
    require_cuda()
    lib = _lib.load()
    images_u8 = _dev(images_u8, torch.uint8, "images")
    if images_u8.dim() != 4 or images_u8.shape[3] != 3:
        raise LighthandError("images: expected [N,H,W,3] uint8")
    n, sh, sw, _ = images_u8.shape
    desc = _dev(desc, torch.uint8, "desc")
    if desc.numel() != n * C.sizeof(_lib.FreiDesc):
        raise LighthandError("desc: expected N packed lh_frei_desc records")
    if out is None:
        out = torch.empty((n, 3, res, res), dtype=torch.float32, device=images_u8.device)
    else:
        out = _dev(out, torch.float32, "out")
        if tuple(out.shape) != (n, 3, res, res):
            raise LighthandError("out: wrong shape")
    if n == 0:
        return out
    max_w, max_h = int(max_rot_wh[0]), int(max_rot_wh[1])
    scratch = torch.empty(int(scratch_floats), dtype=torch.float32, device=images_u8.device) if max_w > 0 else None
    if rot_index is not None:
        rot_index = _dev(rot_index, torch.int32, "rot_index")
    with torch.cuda.device(images_u8.device):
        check(lib.lh_frei_rgb_processing(images_u8.data_ptr(), sh * sw * 3, sh, sw, sw * 3, desc.data_ptr(), n, int(res),
                                         _ptr(rot_index), 0 if rot_index is None else rot_index.numel(),
                                         _ptr(scratch), max_w, max_h, out.data_ptr(), _stream()),
              "lh_frei_rgb_processing")
    return out


def frei_fused(images_u8: torch.Tensor, desc: torch.Tensor, res: int, scratch_floats: int, max_rot_wh: Sequence[int],
               out_hw: Sequence[int], mean=IMAGENET_MEAN, std=IMAGENET_STD, chan_order: int = LH_CHAN_RGB,
               rot_index: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """rgb_processing + Resize(out_hw) + Normalize (+ channel order) in one pass (``lh_frei_fused``): the res x res float image
    never goes to HBM.  Bit-identical to ``frei_rgb_processing`` followed by ``resize_normalize_f32``; falls back to those two
    when the scale res / out is too large for the fused kernel's shared-memory tile."""
    require_cuda()
    lib = _lib.load()
    images_u8 = _dev(images_u8, torch.uint8, "images")
    n, sh, sw, _ = images_u8.shape
    oh, ow = int(out_hw[0]), int(out_hw[1])
    if 32.0 * res / ow + 2.0 > 32 or 32.0 * res / oh + 2.0 > 32:
        img = frei_rgb_processing(images_u8, desc, res, scratch_floats, max_rot_wh, rot_index=rot_index)
        return resize_normalize_f32(img, out_hw, mean, std, chan_order, out)
    desc = _dev(desc, torch.uint8, "desc")
    if desc.numel() != n * C.sizeof(_lib.FreiDesc):
        raise LighthandError("desc: expected N packed lh_frei_desc records")
    if out is None:
        out = torch.empty((n, 3, oh, ow), dtype=torch.float32, device=images_u8.device)
    if n == 0:
        return out
    max_w, max_h = int(max_rot_wh[0]), int(max_rot_wh[1])
    scratch = torch.empty(int(scratch_floats), dtype=torch.float32, device=images_u8.device) if max_w > 0 else None
    if rot_index is not None:
        rot_index = _dev(rot_index, torch.int32, "rot_index")
    with torch.cuda.device(images_u8.device):
        check(lib.lh_frei_fused(images_u8.data_ptr(), sh * sw * 3, sh, sw, sw * 3, desc.data_ptr(), n, int(res),
                                _ptr(rot_index), 0 if rot_index is None else rot_index.numel(), _ptr(scratch), max_w, max_h,
                                _mean_std(mean, std), chan_order, out.data_ptr(), oh, ow, _stream()), "lh_frei_fused")
    return out


def resize_normalize_f32(src: torch.Tensor, out_hw: Sequence[int], mean=IMAGENET_MEAN, std=IMAGENET_STD,
                         chan_order: int = LH_CHAN_RGB, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """transforms.Resize(out_hw) on a float CHW batch + Normalize (+ channel reversal): src [N,3,h,w] float32 CUDA
    -> [N,3,oh,ow] (src/datasets/frei_dataloader.py:279-283,338).  mean=None skips the normalisation."""
    require_cuda()
    lib = _lib.load()
    src = _dev(src, torch.float32, "src")
    if src.dim() != 4 or src.shape[1] != 3:
        raise LighthandError("src: expected [N,3,H,W] float32")
    n, _, sh, sw = src.shape
    oh, ow = int(out_hw[0]), int(out_hw[1])
    if out is None:
        out = torch.empty((n, 3, oh, ow), dtype=torch.float32, device=src.device)
    else:
        out = _dev(out, torch.float32, "out")
        if tuple(out.shape) != (n, 3, oh, ow):
            raise LighthandError("out: wrong shape")
    if n == 0:
        return out
    with torch.cuda.device(src.device):
        check(lib.lh_resize_normalize_f32(src.data_ptr(), n, sh, sw, _mean_std(mean, std), chan_order, out.data_ptr(),
                                          oh, ow, _stream()), "lh_resize_normalize_f32")
    return out


def color_jitter_normalize(images: torch.Tensor, params, mean=IMAGENET_MEAN, std=IMAGENET_STD,
                           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """transforms.ColorJitter -> transforms.Normalize (src/tools/dataset.py:139-145) on float CHW images in 0..1.
    images [N,3,H,W] float32 CUDA; ``params`` = ctypes array of N ``lh_jitter_params``
    (tools.dataset.color_jitter_params); ``out`` may be ``images`` itself.  mean=None skips the normalisation."""
    require_cuda()
    lib = _lib.load()
    images = _dev(images, torch.float32, "images")
    if images.dim() != 4 or images.shape[1] != 3:
        raise LighthandError("images: expected [N,3,H,W] float32")
    n, _, h, w = images.shape
    if len(params) != n:
        raise LighthandError("params: expected one lh_jitter_params per image")
    if out is None:
        out = torch.empty_like(images)
    else:
        out = _dev(out, torch.float32, "out")
        if tuple(out.shape) != tuple(images.shape):
            raise LighthandError("out: wrong shape")
    if n == 0:
        return out
    raw = np.frombuffer(bytes(params), dtype=np.uint8).copy()
    d_params = torch.from_numpy(raw).to(images.device)
    ws = torch.empty(n * _lib.LH_JITTER_PARTS, dtype=torch.float64, device=images.device)
    with torch.cuda.device(images.device):
        check(lib.lh_color_jitter_normalize(images.data_ptr(), d_params.data_ptr(), n, h, w, _mean_std(mean, std),
                                            out.data_ptr(), ws.data_ptr(), _stream()), "lh_color_jitter_normalize")
    return out


# ---------------------------------------------------------------------------------------------
# K2
# ---------------------------------------------------------------------------------------------

def gen_targets(joints: torch.Tensor, heatmap_hw=(64, 64), feat_stride=(4.0, 4.0), sigma: float = 2,
                style: int = LH_TARGET_HRNET, want_weight: bool = True,
                out: Optional[torch.Tensor] = None, use_vis: bool = False):
    """K2.  joints [N,J,2|3] float32 CUDA (network-input pixels) -> (target [N,J,H,W], weight [N,J,1]).
    ``use_vis``: target_weight = visibility (third column) * in-bounds and joints with weight <= 0.5 get an empty map
    (upstream HRNet; the reference ignores visibility -- the default)."""
    require_cuda()
    lib = _lib.load()
    joints = _dev(joints, torch.float32, "joints")
    if joints.dim() != 3 or joints.shape[2] < 2:
        raise LighthandError("joints: expected [N,J,>=2]")
    n, j, js = joints.shape
    h, w = int(heatmap_hw[0]), int(heatmap_hw[1])
    if style == LH_TARGET_HRNET:
        table, side = hrnet_table(sigma, joints.device)
    else:
        if h != w:
            raise LighthandError("GenerateHeatmap style needs a square map")
        table, side = hourglass_table(h, joints.device)
    if out is None:
        out = torch.empty((n, j, h, w), dtype=torch.float32, device=joints.device)
    weight = torch.empty((n, j, 1), dtype=torch.float32, device=joints.device) if want_weight else None
    if n == 0:
        return out, weight
    with torch.cuda.device(joints.device):
        check(lib.lh_gen_targets(joints.data_ptr(), js, n, j, h, w, float(feat_stride[0]), float(feat_stride[1]),
                                 style, int(bool(use_vis)), table.data_ptr(), side, out.data_ptr(), _ptr(weight), _stream()),
              "lh_gen_targets")
    return out, weight


# ---------------------------------------------------------------------------------------------
# K3 / K4 / fused
# ---------------------------------------------------------------------------------------------
_ws_cache = {}


def heatmap_workspace(n: int, j: int, device) -> torch.Tensor:
    """Zero-initialised workspace (ticket + per-map partials); the kernels leave it clean, so one
    buffer per (device, stream, size) is reused."""
    lib = _lib.load()
    key = (str(device), int(torch.cuda.current_stream(device).cuda_stream), n, j)
    ws = _ws_cache.get(key)
    if ws is None:
        nbytes = lib.lh_heatmap_ws_bytes(n, j)
        ws = torch.zeros((nbytes + 7) // 8, dtype=torch.int64, device=device)
        _ws_cache[key] = ws
    return ws


def transform_joints(joints: torch.Tensor, trans: torch.Tensor, out_hw, flip: Optional[torch.Tensor] = None,
                     orig_w: int = 0, swap: Optional[torch.Tensor] = None, valid: Optional[torch.Tensor] = None):
    """Keypoint half of ``augmentation`` (src/utils/preprocessing.py:83-90), batched on the device: mirror + left/right
    swap for flipped samples, ``trans_point2d`` with the crop matrix, ``joint_valid`` in-image test.
    joints [N,J,2|3] float32; trans [N,2,3] (the float32 matrices ``generate_patch_image`` returns, any float dtype);
    returns (joints_out [N,J,3] = x, y, valid; valid [N,J])."""
    require_cuda()
    lib = _lib.load()
    joints = _dev(joints, torch.float32, "joints")
    n, j, js = joints.shape
    trans = _dev(trans.to(torch.float64) if trans.dtype != torch.float64 else trans, torch.float64, "trans")
    flip = _dev(flip, torch.uint8, "flip") if flip is not None else None
    swap = _dev(swap, torch.int32, "swap") if swap is not None else None
    valid = _dev(valid, torch.float32, "valid").reshape(n, j) if valid is not None else None
    out = torch.empty((n, j, 3), dtype=torch.float32, device=joints.device)
    vout = torch.empty((n, j), dtype=torch.float32, device=joints.device)
    with torch.cuda.device(joints.device):
        check(lib.lh_transform_joints(joints.data_ptr(), js, _ptr(valid), trans.data_ptr(), _ptr(flip), int(orig_w),
                                      _ptr(swap), n, j, int(out_hw[1]), int(out_hw[0]), out.data_ptr(), 3,
                                      vout.data_ptr(), _stream()), "lh_transform_joints")
    return out, vout


def bbox_diag(gt: torch.Tensor) -> torch.Tensor:
    """float32 bounding-box diagonal of each sample's joints [N,J,>=2] -> [N] (src/utils/argparser.py:268-272)."""
    require_cuda()
    lib = _lib.load()
    gt = _dev(gt, torch.float32, "gt")
    n, j, gs = gt.shape
    out = torch.empty((n,), dtype=torch.float32, device=gt.device)
    with torch.cuda.device(gt.device):
        check(lib.lh_bbox_diag(gt.data_ptr(), gs, n, j, out.data_ptr(), _stream()), "lh_bbox_diag")
    return out


def jmse_loss(pred: torch.Tensor, target: torch.Tensor, weight: Optional[torch.Tensor] = None,
              want_grad: bool = True):
    """K3.  Returns (loss float32 0-dim CUDA tensor, grad [N,J,H,W] or None)."""
    require_cuda()
    lib = _lib.load()
    pred = _dev(pred, torch.float32, "pred")
    target = _dev(target, torch.float32, "target")
    if pred.dim() != 4 or pred.shape != target.shape:
        raise LighthandError("pred/target: expected equal [N,J,H,W] shapes")
    n, j, h, w = pred.shape
    if weight is not None:
        weight = _dev(weight, torch.float32, "target_weight").reshape(n, j)
    loss = torch.empty((), dtype=torch.float32, device=pred.device)
    grad = torch.empty_like(pred) if want_grad else None
    with torch.cuda.device(pred.device):
        ws = heatmap_workspace(n, j, pred.device)
        check(lib.lh_jmse_loss(pred.data_ptr(), target.data_ptr(), _ptr(weight), n, j, h, w, None,
                               loss.data_ptr(), _ptr(grad), ws.data_ptr(), _stream()), "lh_jmse_loss")
    return loss, grad


def decode(heat: torch.Tensor, out_scale: float = 1.0, subpixel: bool = False,
           inv_affine: Optional[torch.Tensor] = None):
    """K4.  Returns (coords [N,J,2] f32, maxvals [N,J,1] f32, argidx [N,J] i32), all CUDA."""
    require_cuda()
    lib = _lib.load()
    heat = _dev(heat, torch.float32, "heatmaps")
    if heat.dim() != 4:
        raise LighthandError("batch_images should be 4-ndim")
    n, j, h, w = heat.shape
    if inv_affine is not None:
        inv_affine = _dev(inv_affine, torch.float64, "inv_affine")
    coords = torch.empty((n, j, 2), dtype=torch.float32, device=heat.device)
    maxvals = torch.empty((n, j, 1), dtype=torch.float32, device=heat.device)
    argidx = torch.empty((n, j), dtype=torch.int32, device=heat.device)
    if n == 0:
        return coords, maxvals, argidx
    with torch.cuda.device(heat.device):
        check(lib.lh_decode(heat.data_ptr(), n, j, h, w, float(out_scale), int(bool(subpixel)), _ptr(inv_affine),
                            coords.data_ptr(), maxvals.data_ptr(), argidx.data_ptr(), _stream()), "lh_decode")
    return coords, maxvals, argidx


def heatmap_fused(pred: torch.Tensor, joints: Optional[torch.Tensor] = None, *,
                  target_in: Optional[torch.Tensor] = None, sigma: float = 2, feat_stride=(4.0, 4.0),
                  style: int = LH_TARGET_HRNET, want_target: bool = False, want_grad: bool = True,
                  want_decode: bool = True, out_scale: float = 4.0, subpixel: bool = False,
                  inv_affine: Optional[torch.Tensor] = None, use_target_weight: bool = False,
                  loss_meter: Optional[torch.Tensor] = None, buffers: Optional[dict] = None,
                  metrics: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, pck_T: float = 0.2,
                  pck_threshold: str = "proportion", use_vis: bool = False) -> dict:
    """Fused K2+K3+K4: one read of ``pred`` gives loss, grad, target (optional) and decode.
    ``buffers`` may carry preallocated outputs (keys: target, grad, loss, coords, maxvals, argidx, weight)
    so the call is allocation-free (CUDA-graph capturable).  ``metrics=(counts, sums)`` (see
    ``new_metric_state``) additionally folds PCK_2d_loss(T, threshold) + EPE_train of the decoded joints
    against ``joints`` into the accumulators inside the same launch (K5 fused)."""
    require_cuda()
    lib = _lib.load()
    pred = _dev(pred, torch.float32, "pred")
    n, j, h, w = pred.shape
    dev = pred.device
    b = dict(buffers or {})
    a = _lib.HeatmapArgs()
    a.pred = pred.data_ptr()
    if target_in is not None:
        target_in = _dev(target_in, torch.float32, "target")
        a.target_in = target_in.data_ptr()
        if joints is not None:
            joints = _dev(joints, torch.float32, "joints")
            a.joints = joints.data_ptr()
            a.joint_stride = joints.shape[2]
    else:
        if joints is None:
            raise LighthandError("heatmap_fused: joints or target_in required")
        joints = _dev(joints, torch.float32, "joints")
        a.joints = joints.data_ptr()
        a.joint_stride = joints.shape[2]
        if style == LH_TARGET_HRNET:
            table, side = hrnet_table(sigma, dev)
        else:
            table, side = hourglass_table(h, dev)
        a.gauss_table = table.data_ptr()
        a.table_side = side
        if want_target:
            b.setdefault("target", torch.empty_like(pred))
            a.target_out = b["target"].data_ptr()
        b.setdefault("weight", torch.empty((n, j, 1), dtype=torch.float32, device=dev))
        a.target_weight_out = b["weight"].data_ptr()
    if want_grad:
        b.setdefault("grad", torch.empty_like(pred))
        a.grad = b["grad"].data_ptr()
    b.setdefault("loss", torch.empty((), dtype=torch.float32, device=dev))
    a.loss_f32 = b["loss"].data_ptr()
    if loss_meter is not None:
        a.loss_meter = _dev(loss_meter, torch.float64, "loss_meter").data_ptr()
    if want_decode:
        b.setdefault("coords", torch.empty((n, j, 2), dtype=torch.float32, device=dev))
        b.setdefault("maxvals", torch.empty((n, j, 1), dtype=torch.float32, device=dev))
        b.setdefault("argidx", torch.empty((n, j), dtype=torch.int32, device=dev))
        a.coords = b["coords"].data_ptr()
        a.maxvals = b["maxvals"].data_ptr()
        a.argidx = b["argidx"].data_ptr()
    if inv_affine is not None:
        a.inv_affine = _dev(inv_affine, torch.float64, "inv_affine").data_ptr()
    a.N, a.J, a.H, a.W = n, j, h, w
    a.style = style
    a.use_target_weight = int(bool(use_target_weight))
    a.use_vis = int(bool(use_vis))
    a.subpixel = int(bool(subpixel))
    a.stride_x, a.stride_y = float(feat_stride[0]), float(feat_stride[1])
    a.out_scale = float(out_scale)
    if metrics is not None:
        if joints is None or not want_decode:
            raise LighthandError("heatmap_fused: fused metrics need joints and the decode")
        a.metric_counts = _dev(metrics[0], torch.int64, "counts").data_ptr()
        a.metric_sums = _dev(metrics[1], torch.float64, "sums").data_ptr()
        a.pck_mode, a.pck_thr = _pck_mode(pck_T, pck_threshold)
    with torch.cuda.device(dev):
        ws = b.get("ws")
        if ws is None:
            ws = heatmap_workspace(n, j, dev)
        a.ws = ws.data_ptr()
        check(lib.lh_heatmap_fused(C.byref(a), _stream()), "lh_heatmap_fused")
    return b


# ---------------------------------------------------------------------------------------------
# K5
# ---------------------------------------------------------------------------------------------

def _pck_mode(T: float, threshold: str):
    """(mode, float32 threshold) as torch compares it: T for 'proportion', T*3.78 for 'mm' (loss.py:138,144)."""
    if threshold == "proportion":
        return LH_PCK_PROPORTION, float(np.float32(T))
    if threshold == "mm":
        return LH_PCK_MM, float(np.float32(T * 3.78))
    raise AssertionError("Please check variable threshold is right")


def new_metric_state(device) -> Tuple[torch.Tensor, torch.Tensor]:
    """(counts int64[4] = n_wrong, n_total, epe_count, n_images ; sums float64[4] = epe_sum, loss*n, n, -)."""
    return (torch.zeros(4, dtype=torch.int64, device=device), torch.zeros(4, dtype=torch.float64, device=device))


def pck_epe_batch(pred_xy: torch.Tensor, gt: torch.Tensor, T: float, threshold: str, visible: bool,
                  counts: torch.Tensor, sums: torch.Tensor) -> None:
    """K5 in-loop metrics, accumulating into (counts, sums)."""
    require_cuda()
    lib = _lib.load()
    pred_xy = _dev(pred_xy, torch.float32, "pred_2d")
    gt = _dev(gt, torch.float32, "gt_2d")
    n, j, gs = gt.shape
    if pred_xy.shape[0] != n or pred_xy.shape[1] != j or pred_xy.shape[2] != 2:
        raise LighthandError("pred_2d: expected [N,J,2] matching gt")
    mode, thr = _pck_mode(T, threshold)
    counts = _dev(counts, torch.int64, "counts")
    sums = _dev(sums, torch.float64, "sums")
    with torch.cuda.device(gt.device):
        check(lib.lh_pck_epe_batch(pred_xy.data_ptr(), gt.data_ptr(), gs, n, j, thr, mode, int(bool(visible)),
                                   counts.data_ptr(), sums.data_ptr(), _stream()), "lh_pck_epe_batch")


def pck_curve_accum(pred_xy: torch.Tensor, gt: torch.Tensor, thr_sorted: torch.Tensor, method: str,
                    use_vis: bool, hist: torch.Tensor, n_valid: torch.Tensor, n_all: torch.Tensor,
                    sum_d: torch.Tensor, bbox: Optional[torch.Tensor] = None,
                    category: Optional[torch.Tensor] = None) -> None:
    """K5 curve histograms (pred_eval / pred_test counting core), accumulating."""
    require_cuda()
    lib = _lib.load()
    pred_xy = _dev(pred_xy, torch.float32, "pred")
    gt = _dev(gt, torch.float32, "gt")
    thr_sorted = _dev(thr_sorted, torch.float64, "thresholds")
    n, j, gs = gt.shape
    n_cat, kp1 = hist.shape
    if kp1 != thr_sorted.numel() + 1:
        raise LighthandError("hist: expected [n_cat, K+1]")
    if bbox is not None:
        bbox = _dev(bbox, torch.float32, "bbox")
    if category is not None:
        category = _dev(category, torch.int32, "category")
    if method not in ("mm", "pckb"):
        raise AssertionError("this method is the wrong")
    with torch.cuda.device(gt.device):
        check(lib.lh_pck_curve_accum(pred_xy.data_ptr(), gt.data_ptr(), gs, _ptr(bbox), _ptr(category), n_cat, n, j,
                                     thr_sorted.data_ptr(), thr_sorted.numel(), 1 if method == "pckb" else 0,
                                     int(bool(use_vis)), _dev(hist, torch.int64, "hist").data_ptr(),
                                     _dev(n_valid, torch.int64, "n_valid").data_ptr(),
                                     _dev(n_all, torch.int64, "n_all").data_ptr(),
                                     _dev(sum_d, torch.float64, "sum_d").data_ptr(), _stream()),
              "lh_pck_curve_accum")
