"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the LightHand heatmap pipeline.

Every function cites the reference lines it restates (paths relative to the
reference repo root).  The per-sample / per-joint Python loop structure of the
reference is kept on purpose: besides being the parity checker, this module is
what ``bench.py`` times for ``cpu_baseline`` / ``--impl reference`` (kind
"port"), so it must cost what the reference costs.

Pinning: ``tests/test_oracle_golden.py`` checks every function below against
fixtures produced by running the reference's own functions
(``tests/golden/make_golden.py``); ``tests/test_oracle_vs_reference.py`` does
the same live when /root/reference is mounted.  Functions marked
"PARITY UNPINNED" have no counterpart in the reference.

Third-party arithmetic the reference depends on and that is *not* in its tree:
OpenCV (requirements.yaml:87 pins opencv-python 4.6.0.66; 4.13.0 is installed
here) -- ``cv2.warpAffine`` / ``cv2.getAffineTransform``.  ``warp_affine_model``
restates warpAffine's published fixed-point algorithm
and is itself checked against the installed cv2.
"""
from __future__ import annotations

import sys

import numpy as np
import torch

try:  # cv2 is a dependency of the reference, present in the image
    import cv2
except Exception:  # pragma: no cover
    cv2 = None

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)

# ----------------------------------------------------------------------------
# a-1  single-warp crop  (src/utils/preprocessing.py:144-210)
# ----------------------------------------------------------------------------

def rotate_2d(pt, rot_rad):
    """src/utils/preprocessing.py:164-170 (float32 result)."""
    s, c = np.sin(rot_rad), np.cos(rot_rad)
    return np.array([pt[0] * c - pt[1] * s, pt[0] * s + pt[1] * c], dtype=np.float32)


def gen_trans_from_patch_cv(c_x, c_y, src_width, src_height, dst_width, dst_height,
                            scale, rot, inv=False):
    """src/utils/preprocessing.py:172-205: 3 float32 point pairs -> 2x3 float32.

    Uses the installed cv2.getAffineTransform exactly like the reference does.
    """
    sw, sh = src_width * scale, src_height * scale
    centre = np.array([c_x, c_y], dtype=np.float32)
    rad = np.pi * rot / 180
    down = rotate_2d(np.array([0, sh * 0.5], dtype=np.float32), rad)
    right = rotate_2d(np.array([sw * 0.5, 0], dtype=np.float32), rad)
    dcentre = np.array([dst_width * 0.5, dst_height * 0.5], dtype=np.float32)
    ddown = np.array([0, dst_height * 0.5], dtype=np.float32)
    dright = np.array([dst_width * 0.5, 0], dtype=np.float32)
    src = np.stack([centre, centre + down, centre + right]).astype(np.float32)
    dst = np.stack([dcentre, dcentre + ddown, dcentre + dright]).astype(np.float32)
    m = cv2.getAffineTransform(dst, src) if inv else cv2.getAffineTransform(src, dst)
    return m.astype(np.float32)


def trans_point2d(pt_2d, trans):
    """src/utils/preprocessing.py:207-210."""
    return np.dot(trans, np.array([pt_2d[0], pt_2d[1], 1.0]).T)[0:2]


def cv2_invert_affine(M):
    """The double-precision inversion cv2.warpAffine applies to a forward 2x3
    matrix (OpenCV imgwarp.cpp, restated in SURVEY.md Appendix A step 1)."""
    M = np.array(M, dtype=np.float64).reshape(2, 3).copy()
    m0, m1, m2, m3, m4, m5 = M.ravel()
    D = m0 * m4 - m1 * m3
    D = 1.0 / D if D != 0 else 0.0
    a11, a22 = m4 * D, m0 * D
    m0 = a11
    m1 = m1 * (-D)
    m3 = m3 * (-D)
    m4 = a22
    b1 = -m0 * m2 - m1 * m5
    b2 = -m3 * m2 - m4 * m5
    return np.array([[m0, m1, b1], [m3, m4, b2]], dtype=np.float64)


def warp_affine_model(src, M, out_h, out_w, flip=False):
    """Numpy model of cv2.warpAffine(src, M, (out_w,out_h), INTER_LINEAR,
    BORDER_CONSTANT=0) -- SURVEY.md Appendix A; call sites
    src/utils/preprocessing.py:158, src/utils/image_ops.py:43.

    src: HWC uint8 (fixed-point path, uint8 result) or float32 (float path).
    flip=True samples the horizontally mirrored source (preprocessing.py:153).
    """
    src = np.asarray(src)
    if flip:
        src = src[:, ::-1, :]
    sh, sw = src.shape[:2]
    iM = cv2_invert_affine(M)
    xs = np.arange(out_w, dtype=np.float64)
    ys = np.arange(out_h, dtype=np.float64)
    adelta = np.rint(iM[0, 0] * xs * 1024).astype(np.int64)
    bdelta = np.rint(iM[1, 0] * xs * 1024).astype(np.int64)
    X0 = np.rint((iM[0, 1] * ys + iM[0, 2]) * 1024).astype(np.int64) + 16
    Y0 = np.rint((iM[1, 1] * ys + iM[1, 2]) * 1024).astype(np.int64) + 16
    X = (X0[:, None] + adelta[None, :]) >> 5
    Y = (Y0[:, None] + bdelta[None, :]) >> 5
    ix = np.clip(X >> 5, -32768, 32767)
    iy = np.clip(Y >> 5, -32768, 32767)
    ax = (X & 31)
    ay = (Y & 31)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < sh) & (xx >= 0) & (xx < sw)
        v = src[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)]
        return np.where(ok[..., None], v, 0)

    p00, p01 = tap(iy, ix), tap(iy, ix + 1)
    p10, p11 = tap(iy + 1, ix), tap(iy + 1, ix + 1)
    if src.dtype == np.uint8:
        w00 = (32 * (32 - ax) * (32 - ay))[..., None]
        w01 = (32 * ax * (32 - ay))[..., None]
        w10 = (32 * (32 - ax) * ay)[..., None]
        w11 = (32 * ax * ay)[..., None]
        acc = (p00.astype(np.int64) * w00 + p01.astype(np.int64) * w01 +
               p10.astype(np.int64) * w10 + p11.astype(np.int64) * w11 + (1 << 14)) >> 15
        return acc.astype(np.uint8)
    fx = (ax.astype(np.float32) * np.float32(1 / 32))
    fy = (ay.astype(np.float32) * np.float32(1 / 32))
    one = np.float32(1)
    w00 = ((one - fy) * (one - fx))[..., None]
    w01 = ((one - fy) * fx)[..., None]
    w10 = (fy * (one - fx))[..., None]
    w11 = (fy * fx)[..., None]
    f = np.float32
    return (p00.astype(f) * w00 + p01.astype(f) * w01 + p10.astype(f) * w10 + p11.astype(f) * w11).astype(f)


def generate_patch_image(cvimg, bbox, do_flip, scale, rot, out_shape):
    """src/utils/preprocessing.py:144-162 (calls the installed cv2 like the reference)."""
    img = cvimg.copy()
    ih, iw = img.shape[:2]
    cx = float(bbox[0] + 0.5 * bbox[2])
    cy = float(bbox[1] + 0.5 * bbox[3])
    bw, bh = float(bbox[2]), float(bbox[3])
    if do_flip:
        img = img[:, ::-1, :]
        cx = iw - cx - 1
    trans = gen_trans_from_patch_cv(cx, cy, bw, bh, out_shape[1], out_shape[0], scale, rot)
    patch = cv2.warpAffine(img, trans, (int(out_shape[1]), int(out_shape[0])), flags=cv2.INTER_LINEAR)
    patch = patch.astype(np.float32)
    inv = gen_trans_from_patch_cv(cx, cy, bw, bh, out_shape[1], out_shape[0], scale, rot, inv=True)
    return patch, trans, inv


def colour_scale_clip(patch_f32, color_scale):
    """src/utils/preprocessing.py:81 -- np.clip(img * color_scale, 0, 255) (float64 product)."""
    return np.clip(patch_f32 * np.asarray(color_scale)[None, None, :], 0, 255)


def to_tensor_normalize(patch_hwc, mean=IMAGENET_MEAN, std=IMAGENET_STD, divide_255=True):
    """ToTensor (u8 or float HWC -> CHW float32 [/255]) + Normalize, the tail every dataset
    applies after the crop (src/tools/dataset.py:147-157; torchvision semantics:
    x.to(float32).div(255), then (x - mean) / std in float32)."""
    t = torch.from_numpy(np.ascontiguousarray(patch_hwc)).permute(2, 0, 1).contiguous()
    t = t.to(torch.float32)
    if divide_255:
        t = t.div(255)
    m = torch.tensor(mean, dtype=torch.float32).view(-1, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(-1, 1, 1)
    return ((t - m) / s).numpy()


def warp_normalize_sample(img_u8, bbox, do_flip, scale, rot, out_hw, color_scale=None,
                          mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """K1 mode A for one image: generate_patch_image on a uint8 source (north_star feeds
    uint8) -> [colour gain + clip] -> /255 -> Normalize -> CHW float32.
    Returns (chw_f32, trans, inv_trans)."""
    patch, trans, inv = generate_patch_image(img_u8, bbox, do_flip, scale, rot, out_hw)
    if color_scale is not None:
        patch = colour_scale_clip(patch, color_scale).astype(np.float32)
    return to_tensor_normalize(patch, mean, std), trans, inv


# ----------------------------------------------------------------------------
# a-2  ToTensor -> Resize -> Normalize  (src/tools/dataset.py:134-159)
# ----------------------------------------------------------------------------

def resize_bilinear_model(chw_f32, out_h, out_w):
    """Numpy model of torch bilinear resize, align_corners=False, no antialias
    (torchvision 0.12 ``Resize`` on a tensor, src/tools/dataset.py:150) -- SURVEY.md
    Appendix A "Mode B".  float32 arithmetic in ATen's order:
    w_y0*(w_x0*p00 + w_x1*p01) + w_y1*(w_x0*p10 + w_x1*p11)."""
    f = np.float32
    c, ih, iw = chw_f32.shape

    def axis(insz, outsz):
        scale = f(insz) / f(outsz)
        d = np.arange(outsz, dtype=f)
        # ATen evaluates scale * (dst + 0.5) - 0.5 as ONE fused multiply-add (float32); emulated through float64,
        # where the product of two float32 values is exact
        s = (np.float64(scale) * (d + f(0.5)).astype(np.float64) - 0.5).astype(f)
        s = np.maximum(s, f(0)).astype(f)
        i0 = np.minimum(s.astype(np.int64), insz - 1)
        i1 = np.minimum(i0 + 1, insz - 1)
        l1 = np.clip(s - i0.astype(f), f(0), f(1)).astype(f)
        l0 = (f(1) - l1).astype(f)
        return i0, i1, l0, l1

    y0, y1, wy0, wy1 = axis(ih, out_h)
    x0, x1, wx0, wx1 = axis(iw, out_w)
    src = chw_f32.astype(f)
    top = wx0[None, None, :] * src[:, y0][:, :, x0] + wx1[None, None, :] * src[:, y0][:, :, x1]
    bot = wx0[None, None, :] * src[:, y1][:, :, x0] + wx1[None, None, :] * src[:, y1][:, :, x1]
    return (wy0[None, :, None] * top + wy1[None, :, None] * bot).astype(f)


def resize_normalize_sample(img_u8, out_hw, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """K1 mode B for one image, through torch itself (the reference's dependency):
    ToTensor -> Resize (bilinear, antialias off as in torchvision 0.12) -> Normalize."""
    t = torch.from_numpy(np.ascontiguousarray(img_u8)).permute(2, 0, 1).contiguous().to(torch.float32).div(255)
    t = torch.nn.functional.interpolate(t[None], size=tuple(out_hw), mode="bilinear",
                                        align_corners=False, antialias=False)[0]
    m = torch.tensor(mean, dtype=torch.float32).view(-1, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(-1, 1, 1)
    return ((t - m) / s).numpy()


# ----------------------------------------------------------------------------
# a-3  FreiHAND multi-stage crop  (src/utils/image_ops.py:25-131,
#      src/datasets/frei_dataloader.py:134-160,276-338)
# ----------------------------------------------------------------------------
FREI_RES = 224      # HandMeshTSVDataset.img_res (frei_dataloader.py:75)
FREI_OUT = 256      # the Resize after rgb_processing (frei_dataloader.py:279-280)


def frei_get_transform(center, scale, res, rot=0):
    """src/utils/image_ops.py:59-83."""
    h = 200 * scale
    t = np.zeros((3, 3))
    t[0, 0] = float(res[1]) / h
    t[1, 1] = float(res[0]) / h
    t[0, 2] = res[1] * (-float(center[0]) / h + .5)
    t[1, 2] = res[0] * (-float(center[1]) / h + .5)
    t[2, 2] = 1
    if not rot == 0:
        rot = -rot
        rot_mat = np.zeros((3, 3))
        rot_rad = rot * np.pi / 180
        sn, cs = np.sin(rot_rad), np.cos(rot_rad)
        rot_mat[0, :2] = [cs, -sn]
        rot_mat[1, :2] = [sn, cs]
        rot_mat[2, 2] = 1
        t_mat = np.eye(3)
        t_mat[0, 2] = -res[1] / 2
        t_mat[1, 2] = -res[0] / 2
        t_inv = t_mat.copy()
        t_inv[:2, 2] *= -1
        t = np.dot(t_inv, np.dot(rot_mat, np.dot(t_mat, t)))
    return t


def frei_transform(pt, center, scale, res, invert=0, rot=0):
    """src/utils/image_ops.py:85-95 -- the inverse goes through torch.inverse (float64) and the
    result is truncated toward zero."""
    t = frei_get_transform(center, scale, res, rot=rot)
    if invert:
        t = torch.inverse(torch.from_numpy(t)).numpy()
    new_pt = np.array([pt[0] - 1, pt[1] - 1, 1.]).T
    new_pt = np.dot(t, new_pt)
    return new_pt[:2].astype(int) + 1


def frei_crop_geometry(center, scale, res, rot, img_hw):
    """The integer bookkeeping of ``crop`` (src/utils/image_ops.py:99-121): window corners in the
    source (after the rotation padding), the padding, and the copy ranges."""
    ul = np.array(frei_transform([1, 1], center, scale, res, invert=1)) - 1
    br = np.array(frei_transform([res[0] + 1, res[1] + 1], center, scale, res, invert=1)) - 1
    pad = int(np.linalg.norm(br - ul) / 2 - float(br[1] - ul[1]) / 2)
    if not rot == 0:
        ul = ul - pad
        br = br + pad
    ih, iw = img_hw
    return dict(ul=ul, br=br, pad=pad, shape=(int(br[1] - ul[1]), int(br[0] - ul[0])),
                new_x=(max(0, -ul[0]), min(br[0], iw) - ul[0]), new_y=(max(0, -ul[1]), min(br[1], ih) - ul[1]),
                old_x=(max(0, ul[0]), min(iw, br[0])), old_y=(max(0, ul[1]), min(ih, br[1])))


def cv2_rotation_matrix(center, angle, scale=1.0):
    """Model of cv2.getRotationMatrix2D (OpenCV imgwarp.cpp; call site src/utils/image_ops.py:33):
    the centre is a Point2f (float32), the trigonometry is double."""
    cx, cy = float(np.float32(center[0])), float(np.float32(center[1]))
    a = angle * np.pi / 180
    alpha, beta = np.cos(a) * scale, np.sin(a) * scale
    return np.array([[alpha, beta, (1 - alpha) * cx - beta * cy],
                     [-beta, alpha, beta * cx + (1 - alpha) * cy]], dtype=np.float64)


def cv2_resize_linear_model(img, out_w, out_h):
    """Numpy model of cv2.resize(img_float64, (out_w, out_h), INTER_LINEAR) as the installed OpenCV (4.13)
    computes it for CV_64F (call site src/utils/image_ops.py:49): half-pixel source coordinates, edge clamp,
    weights and accumulation in double -- rows first, then columns.  (Agrees with cv2 to 2e-11 on 0..255
    data; a float32-weight variant, which older OpenCV builds use, differs from it by <= 0.005, far inside
    the 1-LSB bar of this path.)"""
    img = np.asarray(img, dtype=np.float64)
    sh, sw = img.shape[:2]

    def axis(ssize, dsize):
        scale = 1.0 / (float(dsize) / float(ssize))
        f = (np.arange(dsize, dtype=np.float64) + 0.5) * scale - 0.5
        s = np.floor(f).astype(np.int64)
        f = f - s
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= ssize - 1
        f[hi] = 0
        s[hi] = ssize - 1
        return s, np.minimum(s + 1, ssize - 1), 1.0 - f, f

    x0, x1, a0, a1 = axis(sw, out_w)
    y0, y1, b0, b1 = axis(sh, out_h)
    rows = img[:, x0] * a0[None, :, None] + img[:, x1] * a1[None, :, None]
    return rows[y0] * b0[:, None, None] + rows[y1] * b1[:, None, None]


def frei_crop(img, center, scale, res, rot=0, model=False):
    """src/utils/image_ops.py:97-131: window copy into a zero float64 canvas, optional rotation about
    the canvas centre (``myimrotate`` :25-44) with the rotation padding removed, cv2.resize to ``res``.
    model=False calls the installed cv2 exactly as the reference does; model=True uses the numpy models
    above (what the CUDA kernels implement)."""
    g = frei_crop_geometry(center, scale, res, rot, img.shape[:2])
    shape = list(g["shape"]) + ([img.shape[2]] if img.ndim > 2 else [])
    new_img = np.zeros(shape)
    (nx0, nx1), (ny0, ny1), (ox0, ox1), (oy0, oy1) = g["new_x"], g["new_y"], g["old_x"], g["old_y"]
    new_img[ny0:ny1, nx0:nx1] = img[oy0:oy1, ox0:ox1]
    if not rot == 0:
        h, w = new_img.shape[:2]
        centre = ((w - 1) * 0.5, (h - 1) * 0.5)
        pad = g["pad"]
        if model:
            new_img = warp_affine_model(new_img.astype(np.float32), cv2_rotation_matrix(centre, rot), h, w).astype(np.float64)
        else:
            new_img = cv2.warpAffine(new_img, cv2.getRotationMatrix2D(centre, rot, 1.0), (w, h), borderValue=0)
        new_img = new_img[pad:-pad, pad:-pad]
    if model:
        return cv2_resize_linear_model(new_img, res[0], res[1])
    return cv2.resize(new_img, (res[0], res[1]), interpolation=cv2.INTER_LINEAR)


def frei_rgb_processing(img, center, scale, rot, pn, res=FREI_RES, model=False):
    """src/datasets/frei_dataloader.py:134-147 with flip = 0 (:107): crop, per-channel pixel noise with
    clip to 0..255 (float64), CHW float32 / 255."""
    rgb = frei_crop(img, center, scale, [res, res], rot=rot, model=model)
    for c in range(3):
        rgb[:, :, c] = np.minimum(255.0, np.maximum(0.0, rgb[:, :, c] * pn[c]))
    return np.transpose(rgb.astype("float32"), (2, 0, 1)) / 255.0


def frei_image(img, center, scale, rot, pn, model=False, mean=IMAGENET_MEAN, std=IMAGENET_STD, res=FREI_RES,
               out=FREI_OUT):
    """The image HandMeshTSVDataset.__getitem__ returns (src/datasets/frei_dataloader.py:276-283,338):
    rgb_processing -> Resize((256,256)) on the float tensor (torchvision 0.12: no antialias) -> Normalize ->
    channels (2,1,0)."""
    t = torch.from_numpy(frei_rgb_processing(img, center, scale, rot, pn, res=res, model=model)).float()
    t = torch.nn.functional.interpolate(t[None], size=(out, out), mode="bilinear",
                                        align_corners=False, antialias=False)[0]
    m = torch.tensor(mean, dtype=torch.float32).view(-1, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(-1, 1, 1)
    return ((t - m) / s)[(2, 1, 0), :, :].numpy()


def frei_j2d_processing(kp, center, scale, r, res=FREI_RES):
    """src/datasets/frei_dataloader.py:148-160 with f = 0: every keypoint through ``transform`` (integer
    truncation), then normalised to [-1, 1]."""
    kp = np.array(kp, dtype=np.float64)
    for i in range(kp.shape[0]):
        kp[i, 0:2] = frei_transform(kp[i, 0:2] + 1, center, scale, [res, res], rot=r)
    kp[:, :-1] = 2. * kp[:, :-1] / res - 1.
    return kp.astype("float32")


def frei_joint_2d(kp, center, scale, r):
    """src/datasets/frei_dataloader.py:334-335: the 2-D joints the heat-maps are drawn from."""
    j = torch.from_numpy(frei_j2d_processing(kp, center, scale, r)).float()
    return ((j[:, :-1] * 100 + 112) * (FREI_OUT / FREI_RES)).numpy()


def color_jitter_sample(chw_f32, fn_idx, b, c, s, h, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """transforms.ColorJitter.forward with GIVEN parameters followed by transforms.Normalize
    (src/tools/dataset.py:139-145), through torchvision's own functional ops like the reference (torchvision is
    its dependency, requirements.yaml:105).  ``fn_idx`` = application order of (0 brightness, 1 contrast,
    2 saturation, 3 hue); the random draw itself is ``ColorJitter.get_params``."""
    from torchvision.transforms import functional as F
    img = torch.from_numpy(np.ascontiguousarray(chw_f32)).float()
    for fn_id in fn_idx:
        fn_id = int(fn_id)
        if fn_id == 0 and b is not None:
            img = F.adjust_brightness(img, b)
        elif fn_id == 1 and c is not None:
            img = F.adjust_contrast(img, c)
        elif fn_id == 2 and s is not None:
            img = F.adjust_saturation(img, s)
        elif fn_id == 3 and h is not None:
            img = F.adjust_hue(img, h)
    if mean is not None:
        m = torch.tensor(mean, dtype=torch.float32).view(-1, 1, 1)
        sd = torch.tensor(std, dtype=torch.float32).view(-1, 1, 1)
        img = (img - m) / sd
    return img.numpy()


# ----------------------------------------------------------------------------
# a-4 / a-5  Gaussian targets
# ----------------------------------------------------------------------------

def gaussian_table(sigma):
    """The constant patch src/tools/dataset.py:189-195 recomputes per joint: numpy float32
    exp of -((x-c)^2+(y-c)^2)/(2 sigma^2), side 2*(3 sigma)+1."""
    tmp = int(sigma * 3)
    size = 2 * tmp + 1
    x = np.arange(0, size, 1, np.float32)
    y = x[:, None]
    c = size // 2
    return np.exp(-((x - c) ** 2 + (y - c) ** 2) / (2 * sigma ** 2)).astype(np.float32)


def generate_target(joints, num_joints=21, heatmap_hw=(64, 64), feat_stride=(4, 4), sigma=2, use_vis=False):
    """src/tools/dataset.py:165-212 (verbatim copies at src/utils/dataset_loader.py:236-286,
    :370-420, src/datasets/frei_dataloader.py:340-390), parametrised.  At the defaults it is
    the reference; other geometries follow upstream HRNet (tmp_size = 3 sigma) -- PARITY
    UNPINNED away from 64x64 / sigma 2 / stride 4.

    Returns (target f32[J,H,W], target_weight f32[J,1]); the reference returns only the
    first (its target_weight is computed and dropped, dataset.py:171,185,212).

    ``use_vis`` -- PARITY UNPINNED, not in the reference (which never looks at visibility): upstream HRNet's
    JointsDataset.generate_target, the semantics BASELINE.json's north star names: ``target_weight[:, 0] =
    joints_vis[:, 0]`` (third joint column), zeroed when the window leaves the map, and the Gaussian is pasted
    only ``if target_weight > 0.5``.  With ``use_vis=False`` (or all-visible joints) it is the function above.
    """
    H, W = heatmap_hw
    joints = torch.as_tensor(joints, dtype=torch.float32)
    weight = np.ones((num_joints, 1), dtype=np.float32)
    if use_vis:
        weight[:, 0] = joints[:, 2].numpy()
    target = np.zeros((num_joints, H, W), dtype=np.float32)
    tmp = int(sigma * 3)
    for j in range(num_joints):
        mu_x = int(joints[j][0] / feat_stride[0] + 0.5)
        mu_y = int(joints[j][1] / feat_stride[1] + 0.5)
        ul = (mu_x - tmp, mu_y - tmp)
        br = (mu_x + tmp + 1, mu_y + tmp + 1)
        if ul[0] >= W or ul[1] >= H or br[0] < 0 or br[1] < 0:
            weight[j] = 0
            continue
        if use_vis and not weight[j, 0] > 0.5:
            continue
        size = 2 * tmp + 1
        x = np.arange(0, size, 1, np.float32)
        y = x[:, None]
        c = size // 2
        g = np.exp(-((x - c) ** 2 + (y - c) ** 2) / (2 * sigma ** 2))
        gx = max(0, -ul[0]), min(br[0], W) - ul[0]
        gy = max(0, -ul[1]), min(br[1], H) - ul[1]
        ix = max(0, ul[0]), min(br[0], W)
        iy = max(0, ul[1]), min(br[1], H)
        target[j][iy[0]:iy[1], ix[0]:ix[1]] = g[gy[0]:gy[1], gx[0]:gx[1]]
    return target, weight


def hourglass_table(output_res):
    """src/utils/dataset_loader.py:23-32 -- float64 (6 sigma + 3)^2 table, sigma = res/64."""
    sigma = output_res / 64
    size = 6 * sigma + 3
    x = np.arange(0, size, 1, float)
    y = x[:, None]
    c = 3 * sigma + 1
    return np.exp(-((x - c) ** 2 + (y - c) ** 2) / (2 * sigma ** 2))


def generate_heatmap_hourglass(p, output_res=64, num_parts=21):
    """src/utils/dataset_loader.py:22-53 (dup src/datasets/frei_dataloader.py:17-46).
    p is joint_2d / 4 (heat-map space).  Gate on x > 0 only, int() truncation, np.maximum paste."""
    g = hourglass_table(output_res)
    sigma = output_res / 64
    hms = np.zeros((num_parts, output_res, output_res), dtype=np.float32)
    for idx in range(num_parts):
        pt = p[idx]
        if not (pt[0] > 0):
            continue
        x, y = int(pt[0]), int(pt[1])
        if x < 0 or y < 0 or x >= output_res or y >= output_res:
            continue
        ul = int(x - 3 * sigma - 1), int(y - 3 * sigma - 1)
        br = int(x + 3 * sigma + 2), int(y + 3 * sigma + 2)
        c, d = max(0, -ul[0]), min(br[0], output_res) - ul[0]
        a, b = max(0, -ul[1]), min(br[1], output_res) - ul[1]
        cc, dd = max(0, ul[0]), min(br[0], output_res)
        aa, bb = max(0, ul[1]), min(br[1], output_res)
        hms[idx, aa:bb, cc:dd] = np.maximum(hms[idx, aa:bb, cc:dd], g[a:b, c:d])
    return hms


# ----------------------------------------------------------------------------
# a-6  JointsMSELoss  (src/utils/loss.py:306-325)
# ----------------------------------------------------------------------------

def joints_mse_loss(output, target, target_weight=None, use_target_weight=False):
    """src/utils/loss.py:312-325: sum_j 0.5*MSE(pred_j, gt_j) / J, torch float32.
    The reference ignores target_weight.  With use_target_weight=True this follows
    upstream HRNet (pred*w, gt*w per joint) -- PARITY UNPINNED; w == 1 reduces to the reference."""
    output = torch.as_tensor(output)
    target = torch.as_tensor(target)
    n, j = output.shape[0], output.shape[1]
    pred = output.reshape(n, j, -1)
    gt = target.reshape(n, j, -1)
    mse = torch.nn.MSELoss(reduction="mean")
    loss = 0
    for k in range(j):
        p, g = pred[:, k], gt[:, k]
        if use_target_weight and target_weight is not None:
            w = torch.as_tensor(target_weight).reshape(n, j, 1)[:, k]
            p, g = p * w, g * w
        loss = loss + 0.5 * mse(p, g)
    return loss / j


def joints_mse_loss_and_grad(output, target, target_weight=None, use_target_weight=False):
    """Loss and d loss / d output through torch autograd (what loss.backward() gives the
    reference at src/utils/method.py:182).  Returns (float loss, grad f32 ndarray)."""
    o = torch.as_tensor(np.asarray(output)).clone().requires_grad_(True)
    loss = joints_mse_loss(o, torch.as_tensor(np.asarray(target)), target_weight, use_target_weight)
    loss.backward()
    return float(loss.item()), o.grad.numpy()


# ----------------------------------------------------------------------------
# a-7  decode  (src/utils/loss.py:327-355)
# ----------------------------------------------------------------------------

def get_max_preds(batch_heatmaps):
    """src/utils/loss.py:327-355: first-index argmax (NaN wins), float32 x = idx % W,
    y = floor(idx / W), zeroed where maxval <= 0."""
    assert isinstance(batch_heatmaps, np.ndarray)
    assert batch_heatmaps.ndim == 4
    n, j, _h, w = batch_heatmaps.shape
    flat = batch_heatmaps.reshape(n, j, -1)
    idx = np.argmax(flat, 2).reshape(n, j, 1)
    maxvals = np.amax(flat, 2).reshape(n, j, 1)
    preds = np.tile(idx, (1, 1, 2)).astype(np.float32)
    preds[:, :, 0] = preds[:, :, 0] % w
    preds[:, :, 1] = np.floor(preds[:, :, 1] / w)
    preds *= np.tile(np.greater(maxvals, 0.0), (1, 1, 2)).astype(np.float32)
    return preds, maxvals


def decode_final(batch_heatmaps, out_scale=4.0, subpixel=False, inv_affine=None):
    """Decode as the callers use it (src/utils/method.py:173-178: get_max_preds then * 4),
    plus two optional sub-steps the reference does NOT implement -- PARITY UNPINNED:

    * subpixel: upstream HRNet get_final_preds quarter-pixel shift -- with px=floor(x+0.5),
      py=floor(y+0.5), if 1 < px < W-1 and 1 < py < H-1 then coords += 0.25*sign(
      [hm[py,px+1]-hm[py,px-1], hm[py+1,px]-hm[py-1,px]]).
    * inv_affine [N,2,3]: map (coords*out_scale) back to image space with the inverse crop
      matrix, as trans_point2d does (src/utils/preprocessing.py:160,207-210), in float64.

    Returns (coords f32[N,J,2], maxvals f32[N,J,1], idx int64[N,J]).
    """
    hm = np.asarray(batch_heatmaps)
    n, j, h, w = hm.shape
    coords, maxvals = get_max_preds(hm)
    idx = np.argmax(hm.reshape(n, j, -1), 2)
    coords = coords.copy()
    if subpixel:
        for a in range(n):
            for b in range(j):
                px = int(np.floor(coords[a, b, 0] + 0.5))
                py = int(np.floor(coords[a, b, 1] + 0.5))
                if 1 < px < w - 1 and 1 < py < h - 1:
                    m = hm[a, b]
                    diff = np.array([m[py][px + 1] - m[py][px - 1], m[py + 1][px] - m[py - 1][px]],
                                    dtype=np.float32)
                    coords[a, b] += np.sign(diff) * np.float32(0.25)
    coords = coords * np.float32(out_scale)
    if inv_affine is not None:
        A = np.asarray(inv_affine, dtype=np.float64).reshape(n, 2, 3)
        c64 = coords.astype(np.float64)
        x = A[:, None, 0, 0] * c64[..., 0] + A[:, None, 0, 1] * c64[..., 1] + A[:, None, 0, 2]
        y = A[:, None, 1, 0] * c64[..., 0] + A[:, None, 1, 1] * c64[..., 1] + A[:, None, 1, 2]
        coords = np.stack([x, y], -1).astype(np.float32)
    return coords.astype(np.float32), maxvals, idx


# ----------------------------------------------------------------------------
# a-8 / a-9  per-batch PCK and EPE  (src/utils/loss.py:28-202)
# ----------------------------------------------------------------------------

def _bbox_diag_list(gt_2d):
    """The per-sample Python loop at src/utils/loss.py:122-127 (float32 tensors, builtin
    max/min over 0-dim tensors)."""
    sizes = []
    for j in gt_2d:
        width = max(j[:, 0]) - min(j[:, 0])
        height = max(j[:, 1]) - min(j[:, 1])
        sizes.append(np.sqrt(width ** 2 + height ** 2))
    return sizes


def pck_2d_counts(pred_2d, gt_2d, T=0.1, threshold="proportion"):
    """src/utils/loss.py:116-148 -> (num_correct, num_total).  All joints, strict '>'."""
    pred_2d = torch.as_tensor(pred_2d).detach().cpu()
    gt_2d = torch.as_tensor(gt_2d).detach().cpu()
    bbox = _bbox_diag_list(gt_2d)
    diff = gt_2d - pred_2d
    distance = diff.square().sum(2).sqrt()
    total = len(distance.flatten())
    if threshold == "proportion":
        norm = distance.permute(1, 0) / torch.tensor(bbox)
        correct = total - len(norm[norm > T])
    elif threshold == "mm":
        correct = total - len(distance[distance > (T * 3.78)])
    else:
        raise AssertionError("Please check variable threshold is right")
    return correct, total


def pck_2d(pred_2d, gt_2d, T=0.1, threshold="proportion"):
    c, t = pck_2d_counts(pred_2d, gt_2d, T, threshold)
    return float(c / t)


def pck_2d_visible_counts(pred_2d, gt_2d, T=0.1, threshold="proportion"):
    """src/utils/loss.py:83-114 -> (num_correct, num_vis): joints 1.., gt[...,2]==1 mask
    (masked distances become 0, never > T)."""
    pred_2d = torch.as_tensor(pred_2d).detach().cpu()
    gt_2d = torch.as_tensor(gt_2d).detach().cpu()
    bbox = _bbox_diag_list(gt_2d)
    gv, vis, pv = gt_2d[:, 1:, :-1], gt_2d[:, 1:, -1], pred_2d[:, 1:]
    vmask = vis == 1
    distance = (gv - pv).square().sum(2).sqrt() * vmask
    num_vis = len(vmask[vmask == True])  # noqa: E712
    if threshold == "proportion":
        norm = distance.permute(1, 0) / torch.tensor(bbox)
        correct = num_vis - len(norm[norm > T])
    elif threshold == "mm":
        correct = num_vis - len(distance[distance > (T * 3.78)])
    else:
        raise AssertionError("Please check variable threshold is right")
    return correct, num_vis


def pck_2d_visible(pred_2d, gt_2d, T=0.1, threshold="proportion"):
    c, v = pck_2d_visible_counts(pred_2d, gt_2d, T, threshold)
    return c / (v + sys.float_info.epsilon)


def pck_2d_list(pred_2d, gt_2d, T_list, threshold="proportion", auc_t=()):
    """src/utils/loss.py:150-202 -> ([[T, pck%], ...], [pck% for auc_t])."""
    out, auc = [], []
    for T in T_list:
        out.append([T, pck_2d(pred_2d, gt_2d, T, threshold) * 100])
    for T in auc_t:
        auc.append(pck_2d(pred_2d, gt_2d, T, threshold) * 100)
    return out, auc


def epe_train(pred_2d, gt_2d):
    """src/utils/loss.py:50-67 -> (sum, count): per joint i>=1 mean distance * n, summed over
    ``range(1, len(distance))`` -- which drops joint J-1 as well as the wrist (Appendix B.7)."""
    pred_2d = torch.as_tensor(pred_2d).detach().cpu()
    gt_2d = torch.as_tensor(gt_2d).detach().cpu()
    n, J = pred_2d.shape[0], pred_2d.shape[1]
    dist = {}
    for i in range(1, J):
        err = []
        for b in range(n):
            err.append(torch.sqrt(torch.sum((pred_2d[b, i] - gt_2d[b, i, :2]) ** 2)))
        m = np.mean(np.array(err))
        dist[str(i)] = [m if not np.isnan(m) else 0, len(err)]
    rows = [[dist[str(i)][0] * dist[str(i)][1], dist[str(i)][1]] for i in range(1, len(dist))]
    rows = np.array(rows)
    return float(np.sum(rows[:, 0])), float(np.sum(rows[:, 1]))


def epe_visible(pred_2d, gt_2d):
    """src/utils/loss.py:28-47: like epe_train but samples with gt[...,2]==0 are skipped."""
    pred_2d = torch.as_tensor(pred_2d).detach().cpu()
    gt_2d = torch.as_tensor(gt_2d).detach().cpu()
    n, J = pred_2d.shape[0], pred_2d.shape[1]
    dist = {}
    for i in range(1, J):
        err = []
        for b in range(n):
            if gt_2d[b, i, 2] == 0:
                continue
            err.append(torch.sqrt(torch.sum((pred_2d[b, i] - gt_2d[b, i, :2]) ** 2)))
        with np.errstate(all="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                m = np.mean(np.array(err))
        dist[str(i)] = [m if not np.isnan(m) else 0, len(err)]
    rows = np.array([[dist[str(i)][0] * dist[str(i)][1], dist[str(i)][1]] for i in range(1, len(dist))])
    return float(np.sum(rows[:, 0])), float(np.sum(rows[:, 1]))


class AverageMeter:
    """src/utils/metric_logger.py:7-28."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

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


# ----------------------------------------------------------------------------
# a-10  offline sweep  (src/utils/argparser.py:246-438)
# ----------------------------------------------------------------------------

MM_PRED_EVAL = 2.83464567
MM_PRED_TEST = 3.7795275591


def eval_thresholds(T_list, method, mm_factor=MM_PRED_EVAL):
    """src/utils/argparser.py:334-342 (pred_eval) / :397-403 (pred_test uses 3.7795275591)."""
    if method == "mm":
        return np.array(np.linspace(T_list[0], T_list[-1], 101)[1:] * mm_factor)
    if method == "pckb":
        return np.array(np.linspace(T_list[0], T_list[-1], 100))
    raise AssertionError("this method is the wrong")


def bbox_diag_store(gt_2d):
    """src/utils/argparser.py:268-272: float32 diag, ``.item()`` -> python float."""
    out = []
    for j in torch.as_tensor(gt_2d):
        width = max(j[:, 0]) - min(j[:, 0])
        height = max(j[:, 1]) - min(j[:, 1])
        out.append(np.sqrt(width ** 2 + height ** 2).item())
    return out


def _trapz(y, x):
    f = getattr(np, "trapz", None) or np.trapezoid
    return f(y, x)


def pred_eval_meta(meta, T_list, method):
    """src/utils/argparser.py:326-388 on an in-memory ``meta`` dict
    {category: {"bb": [n], "pred": [n,21,2], "gt": [n,21,3]}} (what pred_store dumps,
    :247-281).  Returns {category: [auc, epe_mm, pck_t[100]], "mean_auc": [...]}, including
    the 971x21 zero rows seeded into the pooled EPE (:345) and the strict '<' (:371)."""
    thr = eval_thresholds(T_list, method, MM_PRED_EVAL)
    norm_factor = _trapz(np.ones_like(thr), thr)
    pooled = np.empty(0)
    total_epe = np.zeros([971, 21])
    out = {}
    for cat in meta:
        bbox = np.array(meta[cat]["bb"])
        pred = np.array(meta[cat]["pred"])
        gt = np.array(meta[cat]["gt"])
        diff = np.sqrt(np.sum(np.square(gt[:, :, :2] - pred[:, :, :2]), axis=-1))
        nd = diff / bbox[:, None].repeat(21, axis=1) if method == "pckb" else diff
        nd = nd[gt[:, :, -1] == 1]
        total_epe = np.concatenate([total_epe, diff], axis=0)
        pooled = np.concatenate([nd, pooled])
        total = len(nd)
        pck_t = np.array([(np.count_nonzero(nd < T) / total) * 100 for T in thr])
        auc = _trapz(pck_t, thr) / (norm_factor + sys.float_info.epsilon)
        out[str(cat)] = [auc, diff.mean().item() / MM_PRED_TEST, pck_t]
    total = len(pooled)
    pck_t = np.array([(np.count_nonzero(pooled < T) / total) * 100 for T in thr])
    auc = _trapz(pck_t, thr) / (norm_factor + sys.float_info.epsilon)
    out["mean_auc"] = [auc, total_epe.mean() / MM_PRED_TEST, pck_t]
    return out


def pred_test_meta(meta, T_list, method):
    """src/utils/argparser.py:391-438 on an in-memory meta {"pred": [B][n,21,2], "gt": ..., "bb": [B][n]}:
    no visibility mask, thresholds * 3.7795275591.  Returns (auc, mean_diff_px)."""
    thr = eval_thresholds(T_list, method, MM_PRED_TEST)
    norm_factor = _trapz(np.ones_like(thr), thr)
    bbox = np.array([np.array(b) for batch in meta["bb"] for b in batch])
    gt = np.array([np.array(g) for batch in meta["gt"] for g in batch])
    pred = np.array([np.array(p) for batch in meta["pred"] for p in batch])
    diff = np.sqrt(np.sum(np.square(gt - pred), axis=-1))
    nd = (diff / bbox[:, None].repeat(21, axis=1) if method == "pckb" else diff).flatten()
    total = len(nd)
    pck_t = np.array([(np.count_nonzero(nd < T) / total) * 100 for T in thr])
    auc = _trapz(pck_t, thr) / (norm_factor + sys.float_info.epsilon)
    return auc, diff.mean()


# ----------------------------------------------------------------------------
# a-1 (keypoint half)  augmentation()  src/utils/preprocessing.py:83-90
# ----------------------------------------------------------------------------

def augment_keypoints(joint_coord, joint_valid, trans, do_flip, orig_w, input_hw, right=None, left=None):
    """The keypoint half of ``augmentation`` (src/utils/preprocessing.py:83-90), same statements in the same order:
    mirror x about the original image width and swap the right / left joint blocks (coordinates and validity) when
    flipped, ``trans_point2d`` every joint with the float32 crop matrix (float64 dot, stored into the float32 array),
    then ``joint_valid *= (0 <= x < W_in) * (0 <= y < H_in)``.  ``right`` / ``left``: index arrays of the two hands
    (``joint_type`` of the InterHand lineage); None for a single-hand skeleton (nothing to swap).
    Returns (joint_coord f32[J,2+], joint_valid f32[J])."""
    joint_coord = np.array(joint_coord, dtype=np.float32, copy=True)
    joint_valid = np.array(joint_valid, dtype=np.float32, copy=True)
    if do_flip:
        joint_coord[:, 0] = orig_w - joint_coord[:, 0] - 1
        if right is not None:
            joint_coord[right], joint_coord[left] = joint_coord[left].copy(), joint_coord[right].copy()
            joint_valid[right], joint_valid[left] = joint_valid[left].copy(), joint_valid[right].copy()
    for i in range(len(joint_coord)):
        joint_coord[i, :2] = trans_point2d(joint_coord[i, :2], trans)
        joint_valid[i] = joint_valid[i] * (joint_coord[i, 0] >= 0) * (joint_coord[i, 0] < input_hw[1]) * \
            (joint_coord[i, 1] >= 0) * (joint_coord[i, 1] < input_hw[0])
    return joint_coord, joint_valid
