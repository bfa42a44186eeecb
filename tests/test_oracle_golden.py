"""Pins oracle/restate.py against the committed fixtures that the REFERENCE's own functions produced
(tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden
from oracle import restate as R


def test_generate_target_matches_reference_bitexact():
    g = golden("targets.npz")
    for j, want in zip(g["joints"], g["target"]):
        got, _w = R.generate_target(j)
        assert got.dtype == np.float32
        assert np.array_equal(got, want)


def test_generate_target_weight_semantics():
    g = golden("targets.npz")
    tgt, w = R.generate_target(g["joints"][0])
    # weight 0 exactly where the whole window is off the map (src/tools/dataset.py:183-186)
    mx = tgt.reshape(21, -1).max(1)
    assert np.all(mx[w[:, 0] == 0] == 0)
    assert list(np.nonzero(w[:, 0] == 0)[0]) == [9, 10, 11, 20]
    # joint 3 (-30,100): mu_x = int(-7.0) = -7, br = 0 is not < 0 -> weight stays 1 although nothing is pasted
    assert w[3, 0] == 1 and mx[3] == 0


def test_gaussian_table_is_the_pasted_patch():
    t = R.gaussian_table(2)
    assert t.shape == (13, 13) and t.dtype == np.float32 and t[6, 6] == 1.0
    tgt, _ = R.generate_target(np.full((21, 2), 128.0, np.float32))
    assert np.array_equal(tgt[0, 32 - 6:32 + 7, 32 - 6:32 + 7], t)


def test_hourglass_matches_reference_bitexact():
    g = golden("targets.npz")
    for j, want, want_np in zip(g["joints"], g["hourglass"], g["hourglass_np"]):
        assert np.array_equal(R.generate_heatmap_hourglass(torch.tensor(j) / 4), want)
        assert np.array_equal(R.generate_heatmap_hourglass(j / 4), want_np)


def test_joints_mse_loss_and_grad():
    g = golden("loss.npz")
    target = np.stack([R.generate_target(j)[0] for j in g["joints"]])
    loss, grad = R.joints_mse_loss_and_grad(g["pred"], target)
    assert loss == pytest.approx(float(g["loss"]), rel=1e-6)
    np.testing.assert_allclose(grad, g["grad"], rtol=1e-6, atol=1e-12)
    # the reference ignores target_weight (Appendix B.1): same value with the weight argument
    assert float(g["loss_with_weight_arg"]) == pytest.approx(float(g["loss"]), rel=1e-7)
    assert float(R.joints_mse_loss(g["pred"], target, g["weight"], use_target_weight=False)) == pytest.approx(float(g["loss"]), rel=1e-6)
    # closed form (SURVEY a-6)
    closed = 0.5 * np.mean((g["pred"].astype(np.float64) - target) ** 2)
    assert loss == pytest.approx(closed, rel=1e-5)
    np.testing.assert_allclose(grad, (g["pred"] - target) / g["pred"].size, rtol=1e-5, atol=1e-12)


def test_weighted_loss_reduces_to_reference_for_unit_weights():
    g = golden("loss.npz")
    target = np.stack([R.generate_target(j)[0] for j in g["joints"]])
    ones = np.ones((2, 21, 1), np.float32)
    a = float(R.joints_mse_loss(g["pred"], target, ones, use_target_weight=True))
    assert a == pytest.approx(float(g["loss"]), rel=1e-6)


def test_get_max_preds_edge_cases_bitexact():
    g = golden("decode.npz")
    preds, maxvals = R.get_max_preds(g["hm"])
    assert np.array_equal(preds, g["preds"])
    assert np.array_equal(maxvals, g["maxvals"], equal_nan=True)
    p64, m64 = R.get_max_preds(g["hm64"])
    assert np.array_equal(p64, g["preds64"]) and np.array_equal(m64, g["maxvals64"])
    c, _m, idx = R.decode_final(g["hm64"], 4.0)
    assert np.array_equal(idx, g["idx64"])
    assert np.array_equal(c, g["preds64"] * 4)


def test_pck_and_epe_match_reference():
    g = golden("metrics.npz")
    pred, gt, gt3 = g["pred"], g["gt"], g["gt3"]
    for T, want in zip(g["Ts"], g["pck_prop"]):
        assert R.pck_2d(pred, gt, float(T), "proportion") == want
    for T, want in zip([1.0, 3.0, 5.0, 10.0, 20.0], g["pck_mm"]):
        assert R.pck_2d(pred, gt, T, "mm") == want
    for T, want in zip(g["Ts"], g["pck_deg"]):
        assert R.pck_2d(g["pred_deg"], g["gt_deg"], float(T), "proportion") == want
    for T, want in zip(g["Ts"], g["pck_vis_prop"]):
        assert R.pck_2d_visible(pred, gt3, float(T), "proportion") == want
    for T, want in zip([1.0, 3.0, 5.0, 10.0, 20.0], g["pck_vis_mm"]):
        assert R.pck_2d_visible(pred, gt3, T, "mm") == want
    pl, al = R.pck_2d_list(pred, gt, [0.1, 0.2], "proportion", [0.05, 0.3])
    assert np.array_equal(np.array(pl), g["pck_list"]) and np.array_equal(np.array(al), g["pck_list_auc"])
    s, c = R.epe_train(pred, gt)
    assert c == g["epe_train"][1] == 19 * pred.shape[0]            # joints 1..19 (Appendix B.7)
    assert s == pytest.approx(g["epe_train"][0], rel=1e-12)
    s, c = R.epe_visible(pred, gt3)
    assert c == g["epe_vis"][1] and s == pytest.approx(g["epe_vis"][0], rel=1e-12)


def test_average_meter():
    g = golden("metrics.npz")["meter"]
    m = R.AverageMeter(); m.update(0.5, 4); m.update(0.25, 12)
    mp = R.AverageMeter(); mp.update_p(10.0, 4); mp.update_p(30.0, 16)
    assert [m.sum, m.count, m.avg, mp.sum, mp.count, mp.avg] == list(g)


def test_pred_eval_and_pred_test_match_reference():
    g = golden("eval.npz")
    with open(os.path.join(GOLDEN, "eval_meta.json")) as f:
        metas = json.load(f)
    for tag, T_list, method in (("mm30", [0, 30], "mm"), ("mm50", [0, 50], "mm"), ("pckb", [0.1, 0.3], "pckb")):
        out = R.pred_eval_meta(metas["eval"], T_list, method)
        for k, v in out.items():
            assert np.array_equal(np.asarray(v[2]), g[f"{tag}/{k}/pck_t"])          # counts -> bit-exact
            assert v[0] == pytest.approx(float(g[f"{tag}/{k}/auc"]), rel=1e-12)
            assert v[1] == pytest.approx(float(g[f"{tag}/{k}/epe"]), rel=1e-12)
    for tag, T_list, method in (("mm30", [0, 30], "mm"), ("pckb", [0.1, 0.3], "pckb")):
        auc, mean_px = R.pred_test_meta(metas["test"], T_list, method)
        assert auc == pytest.approx(float(g[f"test/{tag}/auc"]), rel=1e-12)
        assert mean_px == pytest.approx(float(g[f"test/{tag}/mean_px"]), rel=1e-12)


def test_generate_patch_image_matches_reference():
    g = golden("warp.npz")
    out_hw = tuple(int(v) for v in g["out_hw"])
    for i in range(len(g["imgs"])):
        args = (g["bbox"][i], bool(g["flip"][i]), float(g["scale"][i]), float(g["rot"][i]), out_hw)
        patch, trans, inv = R.generate_patch_image(g["imgs"][i], *args)
        assert np.array_equal(trans, g["trans"][i]) and np.array_equal(inv, g["inv_trans"][i])
        assert np.array_equal(patch, g["patch_u8src"][i])
        patch_f, _, _ = R.generate_patch_image(g["imgs"][i].astype(np.float32), *args)
        assert np.array_equal(patch_f, g["patch_f32src"][i])
        for p, want in zip(g["pts"], g["pts_moved"][i]):
            assert np.array_equal(R.trans_point2d(p, trans), want)


def test_warp_affine_model_is_cv2_bitexact():
    """The fixed-point model the CUDA kernel is written to == the reference's cv2 output."""
    g = golden("warp.npz")
    oh, ow = (int(v) for v in g["out_hw"])
    for i in range(len(g["imgs"])):
        flip = bool(g["flip"][i])
        u8 = R.warp_affine_model(g["imgs"][i], g["trans"][i], oh, ow, flip=flip)
        assert np.array_equal(u8.astype(np.float32), g["patch_u8src"][i])
        f32 = R.warp_affine_model(g["imgs"][i].astype(np.float32), g["trans"][i], oh, ow, flip=flip)
        assert np.array_equal(f32, g["patch_f32src"][i])


def test_resize_normalize_matches_torchvision_compose():
    g = golden("resize.npz")
    for img, want, want2 in zip(g["imgs"], g["out"], g["out_80x72"]):
        got = R.resize_normalize_sample(img, (64, 64))
        # float tolerance of the path: 1e-5 (north_star); values are O(1)
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(R.resize_normalize_sample(img, (80, 72)), want2, rtol=1e-5, atol=1e-5)
        # the numpy model of the resize (what the kernel implements) against the same fixture
        model = R.resize_bilinear_model(np.transpose(img, (2, 0, 1)).astype(np.float32) / np.float32(255), 64, 64)
        m = np.array(R.IMAGENET_MEAN, np.float32)[:, None, None]
        s = np.array(R.IMAGENET_STD, np.float32)[:, None, None]
        np.testing.assert_allclose((model - m) / s, want, rtol=1e-5, atol=1e-5)


def test_frei_path_matches_reference():
    """a-3: rgb_processing / j2d_processing / the __getitem__ tail, through cv2 like the reference (tolerance 0
    for the float64 cv2 stages) and through the numpy models the CUDA kernels implement."""
    g = golden("frei.npz")
    for i in range(len(g["imgs"])):
        args = (g["imgs"][i], g["centers"][i], float(g["scales"][i]), float(g["rots"][i]), g["pns"][i])
        assert np.array_equal(R.frei_rgb_processing(*args, res=56), g["rgb"][i])
        np.testing.assert_allclose(R.frei_rgb_processing(*args, res=56, model=True), g["rgb"][i], rtol=0, atol=1e-7)
        # Resize: torchvision 0.26 routes through its antialias kernel (same weights when up-sampling, ~1 ulp apart)
        np.testing.assert_allclose(R.frei_image(*args, res=56, out=64), g["final"][i], rtol=1e-5, atol=1e-5)
        kp = R.frei_j2d_processing(g["kps"][i], g["centers"][i], float(g["scales"][i]), float(g["rots"][i]), res=56)
        assert np.array_equal(kp, g["j2d"][i])
    par = g["big_par"][0]
    for model in (False, True):
        c = R.frei_crop(g["big"][0], par[:2], float(par[2]), [224, 224], rot=float(par[3]), model=model)
        np.testing.assert_allclose(c.astype(np.float32), g["crops"][0], rtol=0, atol=0 if not model else 1e-4)


def test_cv2_models_of_the_frei_stages():
    """The numpy models of cv2.getRotationMatrix2D / warpAffine(float64) / resize(float64) against the installed cv2."""
    import cv2
    rng = np.random.default_rng(5)
    for _ in range(6):
        sh, sw = int(rng.integers(40, 120)), int(rng.integers(40, 120))
        img = rng.integers(0, 256, (sh, sw, 3)).astype(np.float64)
        ang = float(rng.uniform(-180, 180))
        c = ((sw - 1) * 0.5, (sh - 1) * 0.5)
        M = R.cv2_rotation_matrix(c, ang)
        np.testing.assert_allclose(M, cv2.getRotationMatrix2D(c, ang, 1.0), rtol=0, atol=1e-12)
        a = cv2.warpAffine(img, cv2.getRotationMatrix2D(c, ang, 1.0), (sw, sh), borderValue=0)
        assert np.array_equal(R.warp_affine_model(img.astype(np.float32), M, sh, sw).astype(np.float64), a)
        np.testing.assert_allclose(R.cv2_resize_linear_model(img, 56, 56),
                                   cv2.resize(img, (56, 56), interpolation=cv2.INTER_LINEAR), rtol=0, atol=1e-9)


def test_augment_keypoints_matches_reference_bitexact():
    """Keypoint half of augmentation() (src/utils/preprocessing.py:83-90): fixtures made by running the reference with
    get_aug_config pinned (tests/golden/make_golden.py: make_augment)."""
    for n_j in (42, 21):
        g = golden(f"augment{n_j}.npz")
        right, left = (np.arange(0, 21), np.arange(21, 42)) if n_j == 42 else (None, None)
        assert g["flip"].any() and not g["flip"].all()
        for i in range(len(g["joints"])):
            jc, jv = R.augment_keypoints(g["joints"][i], g["valid"][i], g["trans"][i], bool(g["flip"][i]), int(g["orig_w"]),
                                         (256, 256), right, left)
            assert np.array_equal(jc, g["joints_out"][i]) and np.array_equal(jv, g["valid_out"][i])


def test_generate_target_use_vis_reduces_to_reference_when_all_visible():
    g = golden("targets.npz")
    j3 = np.concatenate([g["joints"], np.ones((*g["joints"].shape[:2], 1), np.float32)], -1)
    for i in range(4):
        t0, w0 = R.generate_target(g["joints"][i])
        t1, w1 = R.generate_target(j3[i], use_vis=True)
        assert np.array_equal(t0, t1) and np.array_equal(w0, w1)
    j3[0, 3, 2] = 0.0
    t, w = R.generate_target(j3[0], use_vis=True)
    assert not t[3].any() and w[3, 0] == 0.0 and np.array_equal(t[4], R.generate_target(g["joints"][0])[0][4])
