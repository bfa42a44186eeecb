"""K1 parity: cv2.warpAffine's fixed-point grid (uint8 and float32 weight paths), flip, colour gain,
normalisation, layouts; torch-style bilinear resize."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import restate as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from lighthand_b200 import ops as _ops
    _ops.require_cuda()
    return _ops


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def test_patch_golden_u8_and_f32_paths_exact(ops):
    g = golden("warp.npz")
    oh, ow = (int(v) for v in g["out_hw"])
    imgs, M, flip = cu(g["imgs"]), cu(g["trans"].astype(np.float64)), cu(g["flip"].astype(np.uint8))
    for interp, want in ((ops.LH_INTERP_CV2_U8, g["patch_u8src"]), (ops.LH_INTERP_CV2_F32, g["patch_f32src"])):
        out = ops.warp_normalize(imgs, M, (oh, ow), flip=flip, mean=None, std=None, interp=interp,
                                 layout=ops.LH_LAYOUT_HWC)
        got = out.cpu().numpy()
        assert np.abs(got - want).max() <= (0.0 if interp == ops.LH_INTERP_CV2_U8 else 1e-4), interp
        chw = ops.warp_normalize(imgs, M, (oh, ow), flip=flip, mean=None, std=None, interp=interp)
        assert torch.equal(chw, out.permute(0, 3, 1, 2).contiguous())


@pytest.mark.parametrize("src_hw,out_hw", [((256, 256), (256, 256)), ((224, 224), (224, 224)), ((97, 131), (64, 80)), ((300, 200), (96, 100))])
def test_warp_normalize_random_augmentations_vs_cv2(ops, src_hw, out_hw):
    """Whole K1 (warp + /255 + mean/std, CHW) against cv2 + torchvision semantics: identical uint8 pixel
    => identical float (LUT holds torch's float32 result), so the bar 'within 1 LSB' is met with 0."""
    rng = np.random.default_rng(src_hw[0] + out_hw[1])
    n = 6
    imgs = rng.integers(0, 256, (n, *src_hw, 3), dtype=np.uint8)
    Ms, flips, wants = [], [], []
    for i in range(n):
        side = 0.8 * min(src_hw)
        bbox = [(src_hw[1] - side) / 2 + rng.uniform(-8, 8), (src_hw[0] - side) / 2 + rng.uniform(-8, 8), side, side]
        scale = float(np.clip(rng.standard_normal(), -1, 1) * 0.25 + 1)
        rot = float(np.clip(rng.standard_normal(), -2, 2) * 45) if rng.uniform() <= 0.6 else 0.0
        flip = bool(rng.uniform() <= 0.5)
        chw, trans, _ = R.warp_normalize_sample(imgs[i], bbox, flip, scale, rot, out_hw)
        Ms.append(trans.astype(np.float64)); flips.append(flip); wants.append(chw)
    d_imgs, d_M, d_f = cu(imgs), cu(np.stack(Ms)), cu(np.array(flips, np.uint8))
    want = np.stack(wants)
    # exact table normalisation: bit-identical to torch's ToTensor + Normalize of cv2's uint8 pixel
    got = ops.warp_normalize(d_imgs, d_M, out_hw, flip=d_f, exact_norm=True).cpu().numpy()
    assert np.array_equal(got, want)
    # one-FMA normalisation: same uint8 pixel, float result within 2 ulp (values are O(1))
    fast = ops.warp_normalize(d_imgs, d_M, out_hw, flip=d_f, exact_norm=False).cpu().numpy()
    assert np.abs(fast - want).max() <= 5e-7
    # and the raw pixel itself is cv2's, bit for bit, through both code paths
    raw = ops.warp_normalize(d_imgs, d_M, out_hw, flip=d_f, mean=None, std=None).cpu().numpy()
    back = np.rint((want * np.array(R.IMAGENET_STD, np.float32)[None, :, None, None]
                    + np.array(R.IMAGENET_MEAN, np.float32)[None, :, None, None]) * 255)
    assert np.array_equal(raw, back.astype(np.float32))


def test_warp_strong_downscale_falls_back_to_global_taps(ops):
    """Footprint of a 32x32 tile larger than the staging buffer (4x down-scaling): same pixels."""
    rng = np.random.default_rng(15)
    img = rng.integers(0, 256, (1, 512, 512, 3), dtype=np.uint8)
    M = R.gen_trans_from_patch_cv(256, 256, 480, 480, 64, 64, 1.0, 33.0).astype(np.float64)[None]
    out = ops.warp_normalize(cu(img), cu(M), (64, 64), mean=None, std=None, layout=ops.LH_LAYOUT_HWC)
    assert np.array_equal(out[0].cpu().numpy(), R.warp_affine_model(img[0], M[0], 64, 64).astype(np.float32))
    fl = cu(np.ones(1, np.uint8))
    out = ops.warp_normalize(cu(img), cu(M), (64, 64), flip=fl, mean=None, std=None, layout=ops.LH_LAYOUT_HWC)
    assert np.array_equal(out[0].cpu().numpy(), R.warp_affine_model(img[0], M[0], 64, 64, flip=True).astype(np.float32))


def test_warp_crop_far_outside_image_is_constant_border(ops):
    rng = np.random.default_rng(16)
    img = rng.integers(1, 256, (2, 64, 64, 3), dtype=np.uint8)
    Ms = np.stack([np.array([[1.0, 0, 500.0], [0, 1.0, -300.0]]), np.array([[1.0, 0, -20.5], [0, 1.0, 40.25]])])
    out = ops.warp_normalize(cu(img), cu(Ms), (64, 64), mean=None, std=None, layout=ops.LH_LAYOUT_HWC).cpu().numpy()
    assert float(np.abs(out[0]).max()) == 0.0
    for i in range(2):
        assert np.array_equal(out[i], R.warp_affine_model(img[i], Ms[i], 64, 64).astype(np.float32))


def test_warp_unaligned_source_uses_unstaged_path(ops):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (1, 50, 37, 3), dtype=np.uint8)       # pitch 111: not 16-byte aligned
    M = R.gen_trans_from_patch_cv(18, 25, 30, 30, 32, 32, 1.0, 20.0).astype(np.float64)[None]
    out = ops.warp_normalize(cu(img), cu(M), (32, 32), mean=None, std=None, layout=ops.LH_LAYOUT_HWC)
    want = R.warp_affine_model(img[0], M[0], 32, 32).astype(np.float32)
    assert np.array_equal(out[0].cpu().numpy(), want)


def test_colour_gain_clip_matches_reference_semantics(ops):
    rng = np.random.default_rng(6)
    imgs = rng.integers(0, 256, (3, 64, 64, 3), dtype=np.uint8)
    gain = rng.uniform(0.8, 1.2, (3, 3))
    Ms = np.stack([R.gen_trans_from_patch_cv(32, 32, 50, 50, 48, 48, 1.0, 15.0 * i) for i in range(3)]).astype(np.float64)
    out = ops.warp_normalize(cu(imgs), cu(Ms), (48, 48), chan_gain=cu(gain))
    for i in range(3):
        patch = R.warp_affine_model(imgs[i], Ms[i], 48, 48).astype(np.float32)
        want = R.to_tensor_normalize(R.colour_scale_clip(patch, gain[i]).astype(np.float32))
        np.testing.assert_allclose(out[i].cpu().numpy(), want, rtol=0, atol=1e-6)


def test_bgr_channel_order(ops):
    rng = np.random.default_rng(8)
    imgs = cu(rng.integers(0, 256, (2, 40, 40, 3), dtype=np.uint8))
    M = cu(np.stack([np.array([[1.0, 0, 0], [0, 1.0, 0]])] * 2))
    a = ops.warp_normalize(imgs, M, (40, 40), mean=None, std=None)
    b = ops.warp_normalize(imgs, M, (40, 40), mean=None, std=None, chan_order=ops.LH_CHAN_BGR)
    assert torch.equal(a[:, [2, 1, 0]], b)
    assert torch.equal(a, imgs.permute(0, 3, 1, 2).float())          # identity warp reproduces the source


def test_resize_mode_golden(ops):
    g = golden("resize.npz")
    out = ops.warp_normalize(cu(g["imgs"]), None, (64, 64), interp=ops.LH_INTERP_TORCH)
    np.testing.assert_allclose(out.cpu().numpy(), g["out"], rtol=1e-5, atol=1e-5)
    out = ops.warp_normalize(cu(g["imgs"]), None, (80, 72), interp=ops.LH_INTERP_TORCH)
    np.testing.assert_allclose(out.cpu().numpy(), g["out_80x72"], rtol=1e-5, atol=1e-5)


def test_resize_224_to_256_vs_oracle(ops):
    rng = np.random.default_rng(9)
    imgs = rng.integers(0, 256, (2, 224, 224, 3), dtype=np.uint8)
    out = ops.warp_normalize(cu(imgs), None, (256, 256), interp=ops.LH_INTERP_TORCH).cpu().numpy()
    for i in range(2):
        np.testing.assert_allclose(out[i], R.resize_normalize_sample(imgs[i], (256, 256)), rtol=1e-5, atol=1e-5)


def test_tma_kernel_and_record_kernel_agree_bit_for_bit(tmp_path):
    """The TMA-staged persistent kernel (default for 16-byte aligned rows) and the record-staging kernel it replaced
    (LH_K1=records, also the path for unaligned pitches) must produce identical floats: tiles sticking out of the source,
    tiles missing it altogether, flips, the three normalisation modes, a 2-image tail batch."""
    import json, os, subprocess, sys
    from conftest import ROOT
    code = r'''
import sys, json, hashlib
import numpy as np, torch
sys.path.insert(0, %r)
from lighthand_b200 import ops
from oracle import restate as R
rng = np.random.default_rng(5)
res = {}
for S, n in ((256, 7), (224, 2), (512, 3)):
    imgs = rng.integers(0, 256, (n, S, S, 3), dtype=np.uint8)
    Ms, fl = [], []
    for i in range(n):
        side = S * rng.uniform(0.5, 1.6)                      # crops larger than the image: borders and all-outside tiles
        cx, cy = S / 2 + rng.uniform(-S / 3, S / 3), S / 2 + rng.uniform(-S / 3, S / 3)
        Ms.append(R.gen_trans_from_patch_cv(cx, cy, side, side, S, S, rng.uniform(0.75, 1.25), rng.uniform(-90, 90)).astype(np.float64))
        fl.append(rng.uniform() < 0.5)
    d = torch.from_numpy(imgs).cuda(); M = torch.from_numpy(np.stack(Ms)).cuda(); f = torch.from_numpy(np.array(fl, np.uint8)).cuda()
    for tag, kw in (("lut", dict(exact_norm=True)), ("fma", dict(exact_norm=False)), ("raw", dict(mean=None, std=None))):
        out = ops.warp_normalize(d, M, (S, S), flip=f, **kw)
        res["%%d_%%s" %% (S, tag)] = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()
    for out in (256, 224, 200):            # LH_INTERP_TORCH: the ring kernel (16-byte aligned rows) vs resize_pairs_kernel
        r = ops.warp_normalize(d, None, (out, out), interp=ops.LH_INTERP_TORCH)
        res["%%d_resize_%%d" %% (S, out)] = hashlib.sha256(r.cpu().numpy().tobytes()).hexdigest()
    want = np.stack([R.warp_affine_model(imgs[i], Ms[i], S, S, flip=bool(fl[i])) for i in range(n)]).astype(np.float32)
    raw = ops.warp_normalize(d, M, (S, S), flip=f, mean=None, std=None, layout=ops.LH_LAYOUT_HWC).cpu().numpy()
    res["%%d_model" %% S] = bool(np.array_equal(raw, want))
print(json.dumps(res))
''' % ROOT
    outs = []
    for variant in ("", "records"):
        env = dict(os.environ, LH_K1=variant)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert outs[0] == outs[1]
    assert all(v is True for k, v in outs[0].items() if k.endswith("_model"))


def test_random_affine_matrices_vs_cv2_model(ops):
    """Arbitrary 2x3 matrices through the TMA-staged kernel (and its non-staged fallback inside the same launch): shear,
    anisotropic scale from 0.3x to 3x, reflections, translations that push the crop partly or wholly off the source, a
    singular matrix (cv2 inverts it with D = 0) -- every pixel equal to the cv2 fixed-point model."""
    rng = np.random.default_rng(77)
    S = 128
    n = 24
    imgs = rng.integers(0, 256, (n, S, S, 3), dtype=np.uint8)
    Ms = []
    for i in range(n):
        ang, shear = rng.uniform(-np.pi, np.pi), rng.uniform(-0.5, 0.5)
        sx, sy = np.exp(rng.uniform(np.log(0.3), np.log(3.0), 2))
        A = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]) @ np.array([[1, shear], [0, 1]]) @ np.diag([sx, sy])
        if i % 7 == 0:
            A[:, 0] *= -1                                        # reflection
        t = rng.uniform(-S, S, 2) if i % 3 else rng.uniform(0, S / 4, 2)
        Ms.append(np.concatenate([A, t[:, None]], 1).astype(np.float32).astype(np.float64))
    Ms[5] = np.array([[1.0, 2.0, 3.0], [2.0, 4.0, 5.0]])          # singular
    Ms[6] = np.array([[1.0, 0.0, 5000.0], [0.0, 1.0, -5000.0]])   # nothing of the source left
    M = np.stack(Ms)
    for flip in (False, True):
        fl = np.full(n, flip, np.uint8)
        got = ops.warp_normalize(cu(imgs), cu(M), (S, S), flip=cu(fl), mean=None, std=None, layout=ops.LH_LAYOUT_HWC).cpu().numpy()
        for i in range(n):
            want = R.warp_affine_model(imgs[i], M[i], S, S, flip=flip).astype(np.float32)
            assert np.array_equal(got[i], want), (i, flip)
