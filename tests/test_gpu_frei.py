"""a-3 / f-2: the FreiHAND image path on the device against the CPU oracle (which calls cv2 / torch exactly like the
reference) and against the committed fixture the reference itself produced.  Tolerances: the rotated canvas is
bit-identical to cv2 by construction; cv2.resize is blended in float32 instead of float64 -> 1e-4 on the 0..255
scale (bar for this path: 1 uint8 LSB), 2e-6 on the 0..1 scale, 2e-5 after Normalize."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import restate as R

pytestmark = pytest.mark.gpu


def _inputs(n, h, w, seed, res=224):
    rng = np.random.default_rng(seed)
    imgs = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    centers = np.stack([w / 2 + rng.uniform(-50, 50, n), h / 2 + rng.uniform(-50, 50, n)], axis=1)
    scales = rng.uniform(0.5, 1.5, n) * (res / 224.0)
    rots = np.where(np.arange(n) % 3 == 0, 0.0, np.clip(rng.standard_normal(n) * 90, -180, 180))
    pns = rng.uniform(0.6, 1.4, (n, 3))
    return imgs, centers, scales, rots, pns


@pytest.mark.parametrize("h,w", [(224, 224), (180, 260)])
def test_rgb_processing_vs_oracle(h, w):
    from lighthand_b200.datasets import frei_dataloader as F
    imgs, centers, scales, rots, pns = _inputs(9, h, w, 31)
    got = F.rgb_processing(torch.from_numpy(imgs).cuda(), centers, scales, rots, pns).cpu().numpy()
    for i in range(len(imgs)):
        want = R.frei_rgb_processing(imgs[i], centers[i], float(scales[i]), float(rots[i]), pns[i])
        np.testing.assert_allclose(got[i], want, rtol=0, atol=2e-6, err_msg=f"sample {i} rot {rots[i]}")


def test_crop_single_image_call_surface():
    from lighthand_b200.utils import image_ops
    imgs, centers, scales, rots, _ = _inputs(3, 224, 224, 32)
    for i in range(3):
        got = image_ops.crop(imgs[i], centers[i], float(scales[i]), [224, 224], rot=float(rots[i])).cpu().numpy()
        want = R.frei_crop(imgs[i], centers[i], float(scales[i]), [224, 224], rot=float(rots[i]))
        assert got.shape == (224, 224, 3)
        np.testing.assert_allclose(got, want, rtol=0, atol=5e-4)


def test_window_outside_the_image_and_identity_size():
    """Zero padding on every side (window larger than / shifted off the image), and a window of exactly res
    pixels (cv2.resize degenerates to a copy)."""
    from lighthand_b200.datasets import frei_dataloader as F
    rng = np.random.default_rng(33)
    imgs = rng.integers(0, 256, (4, 224, 224, 3), dtype=np.uint8)
    centers = np.array([[10.0, 20.0], [230.0, 215.0], [112.0, 112.0], [112.0, 112.0]])
    scales = np.array([1.2, 0.8, 2.5, 1.12])        # 200*1.12 = 224 -> identity-size window
    rots = np.array([0.0, 33.0, -120.0, 0.0])
    pns = np.ones((4, 3))
    got = F.rgb_processing(torch.from_numpy(imgs).cuda(), centers, scales, rots, pns).cpu().numpy()
    for i in range(4):
        want = R.frei_rgb_processing(imgs[i], centers[i], float(scales[i]), float(rots[i]), pns[i])
        np.testing.assert_allclose(got[i], want, rtol=0, atol=2e-6, err_msg=f"sample {i}")
    g = R.frei_crop_geometry(centers[3], 1.12, [224, 224], 0, (224, 224))
    if g["shape"] == (224, 224):
        # a pure copy: only the final * (1/255) (instead of / 255) can differ, by one ulp
        np.testing.assert_allclose(got[3], R.frei_rgb_processing(imgs[3], centers[3], 1.12, 0.0, pns[3]), rtol=0, atol=6e-8)


def test_resize_normalize_f32_vs_torch():
    from lighthand_b200 import ops
    x = torch.rand((3, 3, 224, 224), device="cuda")
    got = ops.resize_normalize_f32(x, (256, 256), mean=None)
    want = torch.nn.functional.interpolate(x.cpu(), size=(256, 256), mode="bilinear", align_corners=False, antialias=False)
    torch.testing.assert_close(got.cpu(), want, rtol=1e-6, atol=1e-6)     # ATen's CPU kernel contracts differently: ~1 ulp
    got = ops.resize_normalize_f32(x, (200, 300), chan_order=ops.LH_CHAN_BGR)
    w2 = torch.nn.functional.interpolate(x.cpu(), size=(200, 300), mode="bilinear", align_corners=False, antialias=False)
    m = torch.tensor(R.IMAGENET_MEAN).view(1, 3, 1, 1); s = torch.tensor(R.IMAGENET_STD).view(1, 3, 1, 1)
    torch.testing.assert_close(got.cpu(), ((w2 - m) / s)[:, (2, 1, 0)], rtol=1e-5, atol=1e-5)
    assert ops.resize_normalize_f32(x[:0], (8, 8)).shape == (0, 3, 8, 8)


def test_frei_batch_triple_vs_oracle():
    """The (image, joint_2d, heatmap) triple of HandMeshTSVDataset.__getitem__ for a batch."""
    from lighthand_b200.datasets import frei_dataloader as F
    n = 6
    imgs, centers, scales, rots, pns = _inputs(n, 224, 224, 34)
    rng = np.random.default_rng(35)
    kps = np.concatenate([rng.uniform(20, 200, (n, 21, 2)), np.ones((n, 21, 1))], axis=2)
    augm = [(0, pns[i], float(rots[i]), 1.0 + 0.1 * (i % 3 - 1)) for i in range(n)]
    img, j2d, hm = F.frei_batch(torch.from_numpy(imgs).cuda(), kps, centers, scales, augm)
    assert img.shape == (n, 3, 256, 256) and j2d.shape == (n, 21, 2) and hm.shape == (n, 21, 64, 64)
    for i in range(n):
        sc = augm[i][3] * float(scales[i])
        want = R.frei_image(imgs[i], centers[i], sc, float(rots[i]), pns[i])
        np.testing.assert_allclose(img[i].cpu().numpy(), want, rtol=0, atol=2e-5)
        wj = R.frei_joint_2d(kps[i], centers[i], sc, float(rots[i]))
        assert np.array_equal(j2d[i].cpu().numpy(), wj)
        assert np.array_equal(hm[i].cpu().numpy(), R.generate_target(wj, 21, (64, 64), (4, 4), 2)[0])


def test_against_the_reference_fixture():
    from lighthand_b200 import ops
    from lighthand_b200.utils import image_ops
    g = golden("frei.npz")
    x = torch.from_numpy(g["imgs"]).cuda()
    got = image_ops.crop_batch(x, g["centers"], g["scales"], [56, 56], g["rots"], g["pns"])
    np.testing.assert_allclose(got.cpu().numpy(), g["rgb"], rtol=0, atol=2e-6)
    fin = ops.resize_normalize_f32(got, (64, 64), chan_order=ops.LH_CHAN_BGR)
    np.testing.assert_allclose(fin.cpu().numpy(), g["final"], rtol=1e-5, atol=2e-5)
    par = g["big_par"][0]
    c = image_ops.crop(g["big"][0], par[:2], float(par[2]), [224, 224], rot=float(par[3]))
    np.testing.assert_allclose(c.cpu().numpy(), g["crops"][0], rtol=0, atol=5e-4)


def test_empty_batch_and_unrotated_only_batch():
    """N = 0 is a no-op; a batch without any rotated sample needs no scratch canvas and no rotate launch."""
    from lighthand_b200 import ops
    from lighthand_b200.datasets import frei_dataloader as F
    from lighthand_b200.utils import image_ops
    empty = torch.zeros((0, 224, 224, 3), dtype=torch.uint8, device="cuda")
    out = image_ops.crop_batch(empty, [], [], [224, 224], [])
    assert out.shape == (0, 3, 224, 224)
    img, j2d, hm = F.frei_batch(empty, np.zeros((0, 21, 3)), [], [])
    assert img.shape == (0, 3, 256, 256) and j2d.shape == (0, 21, 2) and hm.shape == (0, 21, 64, 64)
    assert ops.color_jitter_normalize(torch.zeros((0, 3, 8, 8), device="cuda"), []).shape == (0, 3, 8, 8)
    imgs, centers, scales, _, pns = _inputs(4, 224, 224, 36)
    rots = np.zeros(4)
    recs, scratch, max_wh = image_ops.crop_descriptors(centers, scales, [224, 224], rots, pns)
    assert scratch == 0 and max_wh == (0, 0)
    got = F.rgb_processing(torch.from_numpy(imgs).cuda(), centers, scales, rots, pns).cpu().numpy()
    for i in range(4):
        want = R.frei_rgb_processing(imgs[i], centers[i], float(scales[i]), 0.0, pns[i])
        np.testing.assert_allclose(got[i], want, rtol=0, atol=2e-6)


def test_fused_rgb_resize_equals_the_two_kernel_path_bitwise():
    """lh_frei_fused (rgb_processing's 224 x 224 image kept in shared memory) == lh_frei_rgb_processing followed by
    lh_resize_normalize_f32, bit for bit: rotated and un-rotated samples, windows sticking out of the image, both channel orders."""
    import torch
    from lighthand_b200 import ops
    from lighthand_b200.datasets import frei_dataloader as F
    from lighthand_b200.utils import image_ops
    rng = np.random.RandomState(11)
    n = 24
    imgs = torch.from_numpy(np.random.default_rng(3).integers(0, 256, (n, 224, 224, 3), dtype=np.uint8)).cuda()
    augm = [F.augm_params(True, rng) for _ in range(n)]
    centers = 112 + rng.uniform(-40, 40, (n, 2))
    scales = [a[3] * rng.uniform(0.7, 1.3) for a in augm]
    recs, sfl, mwh = image_ops.crop_descriptors(centers, scales, [224, 224], [a[2] for a in augm], [a[1] for a in augm])
    assert any(r.rotated for r in recs) and not all(r.rotated for r in recs)
    desc = image_ops.descriptors_to_device(recs, imgs.device)
    rot = torch.tensor([i for i, r in enumerate(recs) if r.rotated], dtype=torch.int32).cuda()
    for order in (ops.LH_CHAN_BGR, ops.LH_CHAN_RGB):
        two = ops.resize_normalize_f32(ops.frei_rgb_processing(imgs, desc, 224, sfl, mwh, rot_index=rot), (256, 256), chan_order=order)
        one = ops.frei_fused(imgs, desc, 224, sfl, mwh, (256, 256), chan_order=order, rot_index=rot)
        assert torch.equal(one, two)
    # a ragged output size and the fallback for scales the fused tile cannot hold
    two = ops.resize_normalize_f32(ops.frei_rgb_processing(imgs, desc, 224, sfl, mwh, rot_index=rot), (250, 270))
    assert torch.equal(ops.frei_fused(imgs, desc, 224, sfl, mwh, (250, 270), rot_index=rot), two)
    two = ops.resize_normalize_f32(ops.frei_rgb_processing(imgs, desc, 224, sfl, mwh, rot_index=rot), (224, 224))
    assert torch.equal(ops.frei_fused(imgs, desc, 224, sfl, mwh, (224, 224), rot_index=rot), two)
