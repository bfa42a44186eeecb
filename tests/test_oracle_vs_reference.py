"""Live pin of oracle/restate.py against the reference imported from /root/reference (build container
only; skipped on the GPU box where the tree does not exist).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import reference_bridge as rb
from oracle import restate as R

pytestmark = pytest.mark.skipif(not rb.available(), reason="/root/reference not mounted")


@pytest.fixture(scope="module")
def ref():
    return rb.load()


def test_targets_random_and_offimage(ref):
    rng = np.random.default_rng(101)
    for _ in range(40):
        j = rng.uniform(-60, 320, (21, 2)).astype(np.float32)
        assert np.array_equal(R.generate_target(j)[0], ref.generate_target(torch.tensor(j)).numpy())
        assert np.array_equal(R.generate_heatmap_hourglass(j / 4), ref.GenerateHeatmap(64, 21)(j / 4))


def test_loss_random(ref):
    rng = np.random.default_rng(102)
    for n in (1, 3):
        p = rng.standard_normal((n, 21, 64, 64)).astype(np.float32)
        t = rng.uniform(size=(n, 21, 64, 64)).astype(np.float32)
        want = ref.JointsMSELoss(False)(torch.tensor(p), torch.tensor(t), None).item()
        assert float(R.joints_mse_loss(p, t)) == pytest.approx(want, rel=1e-6)


def test_decode_random_with_ties(ref):
    rng = np.random.default_rng(103)
    hm = rng.integers(-2, 3, (4, 21, 64, 64)).astype(np.float32)   # many ties
    a, b = R.get_max_preds(hm)
    c, d = ref.get_max_preds(hm)
    assert np.array_equal(a, c) and np.array_equal(b, d)


def test_metrics_random(ref):
    rng = np.random.default_rng(104)
    gt = rng.uniform(0, 256, (16, 21, 3)).astype(np.float32)
    gt[..., 2] = rng.uniform(size=(16, 21)) > 0.3
    pred = (gt[..., :2] + 10 * rng.standard_normal((16, 21, 2))).astype(np.float32)
    tp, tg = torch.tensor(pred), torch.tensor(gt)
    for T in (0.05, 0.2, 0.5):
        assert R.pck_2d(pred, gt[..., :2], T, "proportion") == ref.PCK_2d_loss(tp, tg[..., :2], T, "proportion")
        assert R.pck_2d(pred, gt[..., :2], T * 20, "mm") == ref.PCK_2d_loss(tp, tg[..., :2], T * 20, "mm")
        assert R.pck_2d_visible(pred, gt, T, "proportion") == ref.PCK_2d_loss_visible(tp, tg, T, "proportion")
    (s, c), _ = ref.EPE_train(tp, tg)
    assert R.epe_train(pred, gt) == (pytest.approx(s, rel=1e-12), c)
    (s, c), _ = ref.EPE(tp, tg)
    assert R.epe_visible(pred, gt) == (pytest.approx(s, rel=1e-12), c)


def test_patch_random(ref):
    rng = np.random.default_rng(105)
    img = rng.integers(0, 256, (120, 100, 3), dtype=np.uint8)
    for _ in range(10):
        bbox = [rng.uniform(-10, 30), rng.uniform(-10, 30), rng.uniform(40, 90), rng.uniform(40, 90)]
        flip = bool(rng.integers(2)); scale = float(rng.uniform(0.75, 1.25)); rot = float(rng.uniform(-90, 90))
        a, t, it = R.generate_patch_image(img, bbox, flip, scale, rot, (64, 48))
        b, t2, it2 = ref.generate_patch_image(img, bbox, flip, scale, rot, (64, 48))
        assert np.array_equal(t, t2) and np.array_equal(it, it2) and np.array_equal(a, b)
        assert np.array_equal(R.warp_affine_model(img, t, 64, 48, flip=flip).astype(np.float32), b)


def test_frei_crop_and_dataset_tail_random(ref):
    """a-3 live: crop / rgb_processing / j2d_processing of the reference on random windows (inside, partly and
    mostly outside the image; rotated or not)."""
    import types
    stub = types.SimpleNamespace(img_res=224)
    ds = ref.HandMeshTSVDataset
    rng = np.random.default_rng(106)
    for it in range(14):
        h, w = (224, 224) if it % 2 else (int(rng.integers(150, 300)), int(rng.integers(150, 300)))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        center = [w / 2 + rng.uniform(-60, 60), h / 2 + rng.uniform(-60, 60)]
        scale = float(rng.uniform(0.5, 1.6))
        rot = 0 if it % 3 == 0 else float(np.clip(rng.standard_normal() * 90, -180, 180))
        pn = rng.uniform(0.6, 1.4, 3)
        a = ref.crop(img, center, scale, [224, 224], rot=rot)
        assert np.array_equal(R.frei_crop(img, center, scale, [224, 224], rot=rot), a)
        np.testing.assert_allclose(R.frei_crop(img, center, scale, [224, 224], rot=rot, model=True), a, rtol=0, atol=1e-9)
        assert np.array_equal(R.frei_rgb_processing(img, center, scale, rot, pn),
                              ds.rgb_processing(stub, img, center, scale, rot, 0, pn))
        kp = np.concatenate([rng.uniform(-20, 260, (21, 2)), np.ones((21, 1))], axis=1)
        assert np.array_equal(R.frei_j2d_processing(kp, center, scale, rot),
                              ds.j2d_processing(stub, kp.copy(), center, scale, rot, 0))
        for pt in ([1, 1], [225, 225], [17.5, 90.25]):
            for inv in (0, 1):
                assert np.array_equal(R.frei_transform(pt, center, scale, [224, 224], invert=inv, rot=rot),
                                      ref.transform(pt, center, scale, [224, 224], invert=inv, rot=rot))
