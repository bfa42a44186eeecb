"""Parity of every CUDA kernel (through the C ABI, via lighthand_b200.ops) against the committed
reference fixtures and the CPU oracle on the same seeded inputs.  Bars: bit-exact for targets (uploaded
numpy table), argmax indices, decoded coordinates and PCK counts; 1e-5 relative for loss / gradients /
float sums; warped pixels equal to cv2's in the uint8 domain (bar: 1 LSB)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden
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


# ------------------------------------------------------------------------------------------ K2
def test_gen_targets_golden_bitexact(ops):
    g = golden("targets.npz")
    tgt, w = ops.gen_targets(cu(g["joints"]))
    assert np.array_equal(tgt.cpu().numpy(), g["target"])
    want_w = np.stack([R.generate_target(j)[1] for j in g["joints"]])
    assert np.array_equal(w.cpu().numpy(), want_w)


def test_gen_targets_hourglass_golden_bitexact(ops):
    g = golden("targets.npz")
    hm, _ = ops.gen_targets(cu(g["joints"]) / 4, (64, 64), (1.0, 1.0), style=ops.LH_TARGET_HOURGLASS, want_weight=False)
    assert np.array_equal(hm.cpu().numpy(), g["hourglass"])


@pytest.mark.parametrize("S,H,sigma", [(224, 56, 2), (256, 64, 2), (512, 128, 3), (100, 25, 1)])
def test_gen_targets_other_geometries_vs_oracle(ops, S, H, sigma):
    rng = np.random.default_rng(S)
    joints = rng.uniform(-S / 8, 9 * S / 8, (5, 21, 2)).astype(np.float32)
    stride = S / H
    tgt, w = ops.gen_targets(cu(joints), (H, H), (stride, stride), sigma)
    for i in range(5):
        want, ww = R.generate_target(joints[i], 21, (H, H), (stride, stride), sigma)
        assert np.array_equal(tgt[i].cpu().numpy(), want)
        assert np.array_equal(w[i].cpu().numpy(), ww)


def test_gen_targets_joint_stride3_and_empty(ops):
    g = golden("targets.npz")
    j3 = np.concatenate([g["joints"], np.ones((6, 21, 1), np.float32)], -1)
    tgt, _ = ops.gen_targets(cu(j3))
    assert np.array_equal(tgt.cpu().numpy(), g["target"])
    tgt, _ = ops.gen_targets(torch.empty((0, 21, 2), device="cuda"))
    assert tgt.shape == (0, 21, 64, 64)


# ------------------------------------------------------------------------------------------ K3
def test_jmse_loss_grad_golden(ops):
    g = golden("loss.npz")
    target = np.stack([R.generate_target(j)[0] for j in g["joints"]])
    loss, grad = ops.jmse_loss(cu(g["pred"]), cu(target))
    assert loss.item() == pytest.approx(float(g["loss"]), rel=1e-5)
    np.testing.assert_allclose(grad.cpu().numpy(), g["grad"], rtol=1e-5, atol=1e-12)
    # twice on the same workspace: the kernel leaves the ticket clean, result is deterministic
    loss2, _ = ops.jmse_loss(cu(g["pred"]), cu(target))
    assert loss2.item() == loss.item()


def test_jmse_weighted_vs_oracle(ops):
    g = golden("loss.npz")
    target = np.stack([R.generate_target(j)[0] for j in g["joints"]])
    want, want_grad = R.joints_mse_loss_and_grad(g["pred"], target, g["weight"], use_target_weight=True)
    loss, grad = ops.jmse_loss(cu(g["pred"]), cu(target), cu(g["weight"]))
    assert loss.item() == pytest.approx(want, rel=1e-5)
    np.testing.assert_allclose(grad.cpu().numpy(), want_grad, rtol=1e-5, atol=1e-12)


@pytest.mark.parametrize("shape", [(3, 21, 56, 56), (2, 21, 128, 128), (1, 5, 7, 9), (4, 21, 64, 64)])
def test_jmse_shapes_vs_oracle(ops, shape):
    rng = np.random.default_rng(sum(shape))
    p = rng.standard_normal(shape).astype(np.float32)
    t = rng.uniform(size=shape).astype(np.float32)
    want, want_grad = R.joints_mse_loss_and_grad(p, t)
    loss, grad = ops.jmse_loss(cu(p), cu(t))
    assert loss.item() == pytest.approx(want, rel=1e-5)
    np.testing.assert_allclose(grad.cpu().numpy(), want_grad, rtol=1e-5, atol=1e-12)


# ------------------------------------------------------------------------------------------ K4
def test_decode_edge_cases_bitexact(ops):
    g = golden("decode.npz")
    coords, maxvals, idx = ops.decode(cu(g["hm"]), 1.0)
    assert np.array_equal(coords.cpu().numpy(), g["preds"])
    assert np.array_equal(maxvals.cpu().numpy(), g["maxvals"], equal_nan=True)
    hm = g["hm"]
    want_idx = np.argmax(hm.reshape(3, 21, -1), 2)
    assert np.array_equal(idx.cpu().numpy(), want_idx)


def test_decode_64_times4_bitexact(ops):
    g = golden("decode.npz")
    coords, maxvals, idx = ops.decode(cu(g["hm64"]), 4.0)
    assert np.array_equal(coords.cpu().numpy(), g["preds64"] * 4)
    assert np.array_equal(maxvals.cpu().numpy(), g["maxvals64"])
    assert np.array_equal(idx.cpu().numpy(), g["idx64"])


@pytest.mark.parametrize("shape", [(3, 21, 56, 56), (2, 21, 128, 128), (2, 3, 5, 7), (1, 21, 64, 48)])
def test_decode_ties_shapes_vs_oracle(ops, shape):
    rng = np.random.default_rng(sum(shape) + 1)
    hm = rng.integers(-3, 3, shape).astype(np.float32)        # heavy ties, negative maxima
    hm[0, 0] = -2.0
    coords, maxvals, idx = ops.decode(cu(hm), 1.0)
    p, m = R.get_max_preds(hm)
    assert np.array_equal(coords.cpu().numpy(), p) and np.array_equal(maxvals.cpu().numpy(), m)
    assert np.array_equal(idx.cpu().numpy(), np.argmax(hm.reshape(shape[0], shape[1], -1), 2))


def test_decode_subpixel_and_inverse_affine_vs_oracle(ops):
    """PARITY UNPINNED sub-steps (not in the reference): upstream-HRNet quarter-pixel shift + inverse crop matrix."""
    rng = np.random.default_rng(7)
    joints = rng.uniform(0, 256, (4, 21, 2)).astype(np.float32)
    tgt = np.stack([R.generate_target(j)[0] for j in joints])
    hm = (tgt + 0.05 * rng.standard_normal(tgt.shape)).astype(np.float32)
    inv = np.stack([R.gen_trans_from_patch_cv(120 + i, 130, 200, 200, 256, 256, 1.1, 10.0 * i, inv=True) for i in range(4)]).astype(np.float64)
    for sub in (False, True):
        for A in (None, inv):
            want, wm, widx = R.decode_final(hm, 4.0, sub, A)
            coords, maxvals, idx = ops.decode(cu(hm), 4.0, sub, None if A is None else cu(A))
            assert np.array_equal(coords.cpu().numpy(), want)
            assert np.array_equal(idx.cpu().numpy(), widx)


# ------------------------------------------------------------------------------------------ fused
def test_fused_equals_separate_kernels(ops):
    rng = np.random.default_rng(21)
    joints = rng.uniform(-20, 276, (6, 21, 2)).astype(np.float32)
    d_j = cu(joints)
    tgt, w = ops.gen_targets(d_j)
    pred = (tgt + 0.1 * torch.randn_like(tgt)).contiguous()
    loss, grad = ops.jmse_loss(pred, tgt)
    coords, maxvals, idx = ops.decode(pred, 4.0)
    out = ops.heatmap_fused(pred, d_j, want_target=True)
    assert torch.equal(out["target"], tgt) and torch.equal(out["weight"], w)
    assert torch.equal(out["grad"], grad) and out["loss"].item() == loss.item()
    assert torch.equal(out["coords"], coords) and torch.equal(out["maxvals"], maxvals) and torch.equal(out["argidx"], idx)
    # target_in variant
    out2 = ops.heatmap_fused(pred, target_in=tgt)
    assert torch.equal(out2["grad"], grad) and out2["loss"].item() == loss.item() and torch.equal(out2["coords"], coords)
    # use_target_weight from the generated weights == explicit-weight loss
    lw, gw = ops.jmse_loss(pred, tgt, w)
    out3 = ops.heatmap_fused(pred, d_j, use_target_weight=True)
    assert out3["loss"].item() == lw.item() and torch.equal(out3["grad"], gw)


def test_fused_metrics_equal_separate_k5(ops):
    """K5 folded into the fused launch == lh_pck_epe_batch on the decoded joints (counts bit-exact)."""
    rng = np.random.default_rng(23)
    for n, mode, T in ((7, "proportion", 0.2), (33, "mm", 3.0), (256, "proportion", 0.05)):
        joints = rng.uniform(-10, 266, (n, 21, 2)).astype(np.float32)
        d_j = cu(joints)
        tgt, _ = ops.gen_targets(d_j)
        pred = (tgt + 0.3 * torch.randn_like(tgt)).contiguous()
        fused = ops.new_metric_state("cuda")
        for _ in range(2):                                    # accumulates over launches
            out = ops.heatmap_fused(pred, d_j, metrics=fused, pck_T=T, pck_threshold=mode)
        sep = ops.new_metric_state("cuda")
        for _ in range(2):
            ops.pck_epe_batch(out["coords"], d_j, T, mode, False, *sep)
        assert fused[0].cpu().tolist() == sep[0].cpu().tolist()
        assert fused[1][0].item() == pytest.approx(sep[1][0].item(), rel=1e-12)
        want = R.pck_2d_counts(out["coords"].cpu().numpy(), joints, T, mode)
        c = fused[0].cpu().tolist()
        assert (c[1] - c[0], c[1]) == (2 * want[0], 2 * want[1])


def test_fused_loss_meter_accumulates_like_average_meter(ops):
    rng = np.random.default_rng(22)
    meter = torch.zeros(2, dtype=torch.float64, device="cuda")
    m = R.AverageMeter()
    for n in (3, 5):
        joints = cu(rng.uniform(0, 256, (n, 21, 2)).astype(np.float32))
        pred = torch.rand((n, 21, 64, 64), device="cuda")
        out = ops.heatmap_fused(pred, joints, loss_meter=meter)
        m.update(out["loss"].item(), n)
    s = meter.cpu().tolist()
    assert s[1] == m.count and s[0] == pytest.approx(m.sum, rel=1e-12)


# ------------------------------------------------------------------------------------------ K5
def _counts(ops, pred, gt, T, mode, visible):
    counts, sums = ops.new_metric_state("cuda")
    ops.pck_epe_batch(cu(pred), cu(gt), T, mode, visible, counts, sums)
    return counts.cpu().tolist(), sums.cpu().tolist()


def test_pck_epe_golden_bitexact_counts(ops):
    g = golden("metrics.npz")
    pred, gt, gt3 = g["pred"], g["gt"], g["gt3"]
    n = pred.shape[0]
    for T, want in zip(g["Ts"], g["pck_prop"]):
        c, _ = _counts(ops, pred, gt, float(T), "proportion", False)
        assert c[1] == n * 21 and (c[1] - c[0]) / c[1] == want
        assert (c[1] - c[0], c[1]) == R.pck_2d_counts(pred, gt, float(T), "proportion")
    for T, want in zip([1.0, 3.0, 5.0, 10.0, 20.0], g["pck_mm"]):
        c, _ = _counts(ops, pred, gt, T, "mm", False)
        assert (c[1] - c[0]) / c[1] == want
    for T, want in zip(g["Ts"], g["pck_deg"]):                       # zero-diagonal bbox: inf / nan ratios
        c, _ = _counts(ops, g["pred_deg"], g["gt_deg"], float(T), "proportion", False)
        assert (c[1] - c[0]) / c[1] == want
    import sys
    for T, want in zip(g["Ts"], g["pck_vis_prop"]):
        c, _ = _counts(ops, pred, gt3, float(T), "proportion", True)
        assert (c[1] - c[0]) / (c[1] + sys.float_info.epsilon) == want
    for T, want in zip([1.0, 3.0, 5.0, 10.0, 20.0], g["pck_vis_mm"]):
        c, _ = _counts(ops, pred, gt3, T, "mm", True)
        assert (c[1] - c[0]) / (c[1] + sys.float_info.epsilon) == want
    c, s = _counts(ops, pred, gt, 0.2, "proportion", False)
    assert c[2] == g["epe_train"][1] and s[0] == pytest.approx(g["epe_train"][0], rel=1e-6)
    assert c[3] == n
    c, s = _counts(ops, pred, gt3, 0.2, "proportion", True)
    assert c[2] == g["epe_vis"][1] and s[0] == pytest.approx(g["epe_vis"][0], rel=1e-6)


def test_pck_counts_random_vs_oracle_and_shard_additivity(ops):
    rng = np.random.default_rng(31)
    n = 777
    gt = rng.uniform(0, 256, (n, 21, 2)).astype(np.float32)
    pred = np.floor((gt + 9 * rng.standard_normal((n, 21, 2))) / 4).astype(np.float32) * 4
    for T in (0.05, 0.1, 0.2):
        c, _ = _counts(ops, pred, gt, T, "proportion", False)
        assert (c[1] - c[0], c[1]) == R.pck_2d_counts(pred, gt, T, "proportion")
    # additive over shards (what the multi-GPU all-reduce relies on)
    counts, sums = ops.new_metric_state("cuda")
    for lo, hi in ((0, 300), (300, 301), (301, 777)):
        ops.pck_epe_batch(cu(pred[lo:hi]), cu(gt[lo:hi]), 0.2, "proportion", False, counts, sums)
    c, s = _counts(ops, pred, gt, 0.2, "proportion", False)
    assert counts.cpu().tolist() == c
    assert sums.cpu().tolist()[0] == pytest.approx(s[0], rel=1e-12)


def test_pck_curve_vs_reference_fixture(ops):
    from lighthand_b200.utils import argparser as A
    g = golden("eval.npz")
    with open(os.path.join(GOLDEN, "eval_meta.json")) as f:
        metas = json.load(f)
    for tag, T_list, method in (("mm30", [0, 30], "mm"), ("mm50", [0, 50], "mm"), ("pckb", [0.1, 0.3], "pckb")):
        out = A.pred_eval_meta(metas["eval"], T_list, method)
        assert set(out) == {k.split("/")[1] for k in g.files if k.startswith(tag + "/")}
        for k, v in out.items():
            assert np.array_equal(np.asarray(v[2]), g[f"{tag}/{k}/pck_t"])           # counts: bit-exact
            assert v[0] == pytest.approx(float(g[f"{tag}/{k}/auc"]), rel=1e-12)
            assert v[1] == pytest.approx(float(g[f"{tag}/{k}/epe"]), rel=1e-9)
    for tag, T_list, method in (("mm30", [0, 30], "mm"), ("pckb", [0.1, 0.3], "pckb")):
        auc, mean_px = A.pred_test_meta(metas["test"], T_list, method)
        assert auc == pytest.approx(float(g[f"test/{tag}/auc"]), rel=1e-12)
        assert mean_px == pytest.approx(float(g[f"test/{tag}/mean_px"]), rel=1e-9)


def test_pck_curve_bbox_computed_on_device_and_large_random(ops):
    from lighthand_b200.utils import argparser as A
    rng = np.random.default_rng(41)
    n = 5000
    gt = rng.uniform(20, 236, (n, 21, 2)).astype(np.float32)
    vis = (rng.uniform(size=(n, 21, 1)) > 0.2).astype(np.float32)
    pred = np.floor((gt + 6 * rng.standard_normal((n, 21, 2))) / 4).astype(np.float32) * 4
    cat = rng.integers(0, 4, n).astype(np.int32)
    gt3 = np.concatenate([gt, vis], -1)
    names = list(A.POSE_CATEGORIES)
    meta = {}
    for ci, name in enumerate(names):
        sel = cat == ci
        meta[name] = {"bb": R.bbox_diag_store(gt[sel]), "pred": pred[sel].tolist(), "gt": gt3[sel].tolist()}
    for T_list, method in (([0, 30], "mm"), ([0.1, 0.3], "pckb")):
        want = R.pred_eval_meta(meta, T_list, method)
        acc = A.CurveAccumulator(T_list, method)
        for lo in range(0, n, 1024):                                   # batches, bbox computed in-kernel
            acc.update(cu(pred[lo:lo + 1024]), cu(gt3[lo:lo + 1024]), None, cu(cat[lo:lo + 1024]))
        got = acc.finalize(names)
        for k in want:
            assert np.array_equal(np.asarray(got[k][2]), np.asarray(want[k][2]))
            assert got[k][0] == pytest.approx(want[k][0], rel=1e-12)
            assert got[k][1] == pytest.approx(want[k][1], rel=1e-9)


def test_transform_joints_golden_bitexact(ops):
    """lh_transform_joints against fixtures produced by the reference's augmentation() (flip mirror, right / left swap,
    trans_point2d, joint_valid in-image test; src/utils/preprocessing.py:83-90)."""
    for n_j in (42, 21):
        g = golden(f"augment{n_j}.npz")
        swap = None
        if n_j == 42:
            swap = cu(np.concatenate([np.arange(21, 42), np.arange(0, 21)]).astype(np.int32))
        out, valid = ops.transform_joints(cu(g["joints"]), cu(g["trans"].astype(np.float64)), (256, 256),
                                          flip=cu(g["flip"].astype(np.uint8)), orig_w=int(g["orig_w"]), swap=swap,
                                          valid=cu(g["valid"]))
        assert np.array_equal(out[..., :2].cpu().numpy(), g["joints_out"][..., :2])
        assert np.array_equal(valid.cpu().numpy(), g["valid_out"]) and np.array_equal(out[..., 2].cpu().numpy(), g["valid_out"])


def test_targets_and_fused_pass_with_visibility_vs_oracle(ops):
    """north_star item (2): target_weight = visibility * in-bounds (upstream HRNet; parity unpinned), masking the loss."""
    rng = np.random.default_rng(31)
    for (S, H, sigma, n) in ((256, 64, 2, 6), (224, 56, 2, 3)):
        stride = S / H
        j = edge_like(rng, n, S)
        vis = (rng.uniform(size=(n, 21, 1)) < 0.7).astype(np.float32)
        j3 = np.concatenate([j, vis], -1).astype(np.float32)
        want = [R.generate_target(j3[i], 21, (H, H), (stride, stride), sigma, use_vis=True) for i in range(n)]
        wt, ww = np.stack([w[0] for w in want]), np.stack([w[1] for w in want])
        tgt, wgt = ops.gen_targets(cu(j3), (H, H), (stride, stride), sigma, use_vis=True)
        assert np.array_equal(tgt.cpu().numpy(), wt) and np.array_equal(wgt.cpu().numpy(), ww)
        assert (ww == 0).any() and (ww == 1).any()
        # off: the reference's behaviour (visibility ignored)
        t0, w0 = ops.gen_targets(cu(j3), (H, H), (stride, stride), sigma)
        assert np.array_equal(t0.cpu().numpy(), np.stack([R.generate_target(j[i], 21, (H, H), (stride, stride), sigma)[0] for i in range(n)]))
        # fused pass: targets from the joints, loss masked by the same weights
        pred = (wt + 0.1 * rng.standard_normal(wt.shape)).astype(np.float32)
        loss, grad = R.joints_mse_loss_and_grad(pred, wt, ww, use_target_weight=True)
        out = ops.heatmap_fused(cu(pred), cu(j3), sigma=sigma, feat_stride=(stride, stride), want_target=True,
                                use_target_weight=True, use_vis=True, out_scale=stride)
        assert np.array_equal(out["target"].cpu().numpy(), wt) and np.array_equal(out["weight"].cpu().numpy(), ww)
        assert out["loss"].item() == pytest.approx(loss, rel=1e-5)
        np.testing.assert_allclose(out["grad"].cpu().numpy(), grad, rtol=1e-5, atol=1e-12)


def edge_like(rng, n, S):
    j = rng.uniform(0, S, (n, 21, 2)).astype(np.float32)
    off = rng.uniform(size=(n, 21)) < 0.15
    j[off] = rng.uniform(-S / 4, 5 * S / 4, (int(off.sum()), 2)).astype(np.float32)
    return j


def test_bbox_diag_vs_oracle(ops):
    rng = np.random.default_rng(8)
    gt = rng.uniform(0, 256, (37, 21, 3)).astype(np.float32)
    gt[3, :, :2] = 5.0                                       # zero diagonal
    got = ops.bbox_diag(cu(gt)).cpu().numpy()
    assert got.tolist() == R.bbox_diag_store(torch.from_numpy(gt))
