"""Edge cases through the C ABI: odd shapes (scalar kernel), many joints, non-finite inputs, random small
geometries (hypothesis) -- always against the CPU oracle."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from conftest import golden
from oracle import restate as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from lighthand_b200 import ops as _ops
    _ops.require_cuda()
    return _ops


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("J,H,W", [(21, 30, 62), (5, 9, 7), (40, 64, 64), (33, 16, 20), (1, 64, 64)])
def test_fused_odd_shapes_and_many_joints_vs_oracle(ops, J, H, W):
    """W % 4 != 0 takes the scalar kernel; J > 32 walks the bbox reduction in chunks."""
    rng = np.random.default_rng(J * 1000 + W)
    n = 5
    joints = rng.uniform(-8, 4 * max(H, W) + 8, (n, J, 2)).astype(np.float32)
    tgt = np.stack([R.generate_target(j, J, (H, W), (4.0, 4.0), 2)[0] for j in joints])
    pred = (tgt + 0.2 * rng.standard_normal(tgt.shape)).astype(np.float32)
    state = ops.new_metric_state("cuda")
    out = ops.heatmap_fused(cu(pred), cu(joints), want_target=True, metrics=state, pck_T=0.2)
    assert np.array_equal(out["target"].cpu().numpy(), tgt)
    loss, grad = R.joints_mse_loss_and_grad(pred, tgt)
    assert out["loss"].item() == pytest.approx(loss, rel=1e-5)
    np.testing.assert_allclose(out["grad"].cpu().numpy(), grad, rtol=1e-5, atol=1e-12)
    coords, maxvals, idx = R.decode_final(pred, 4.0)
    assert np.array_equal(out["coords"].cpu().numpy(), coords) and np.array_equal(out["argidx"].cpu().numpy(), idx)
    assert np.array_equal(out["maxvals"].cpu().numpy(), maxvals)
    correct, total = R.pck_2d_counts(coords, joints, 0.2, "proportion")
    c = state[0].cpu().tolist()
    assert (c[1] - c[0], c[1], c[3]) == (correct, total, n)
    if J >= 3:
        esum, ecnt = R.epe_train(coords, joints)
        assert c[2] == ecnt and state[1][0].item() == pytest.approx(esum, rel=1e-6)
    # and the stand-alone K5 kernel agrees
    sep = ops.new_metric_state("cuda")
    ops.pck_epe_batch(out["coords"], cu(joints), 0.2, "proportion", False, *sep)
    assert sep[0].cpu().tolist() == c


def test_fused_decode_nonfinite_maps(ops):
    """NaN is numpy's maximum (first NaN wins), -inf maps decode to index 0, +inf wins over finite values."""
    g = golden("decode.npz")
    hm = g["hm"]                                   # [3,21,16,16] with the crafted edge maps in hm[0]
    joints = np.full((3, 21, 2), 20.0, np.float32)
    out = ops.heatmap_fused(cu(hm), cu(joints), out_scale=1.0, feat_stride=(4.0, 4.0))
    assert np.array_equal(out["coords"].cpu().numpy(), g["preds"])
    assert np.array_equal(out["maxvals"].cpu().numpy(), g["maxvals"], equal_nan=True)
    assert np.array_equal(out["argidx"].cpu().numpy(), np.argmax(hm.reshape(3, 21, -1), 2))


def test_gen_targets_nonfinite_joints_give_empty_maps(ops):
    """int(nan) raises in the reference; the device path defines it: no Gaussian, weight 0.  +-inf are off-image."""
    j = np.full((1, 21, 2), 100.0, np.float32)
    j[0, 0] = [np.nan, 10]; j[0, 1] = [10, np.nan]; j[0, 2] = [np.inf, 10]; j[0, 3] = [-np.inf, 10]; j[0, 4] = [1e30, -1e30]
    tgt, w = ops.gen_targets(cu(j))
    t, w = tgt.cpu().numpy(), w.cpu().numpy()
    assert float(np.abs(t[0, :5]).max()) == 0.0 and list(w[0, :5, 0]) == [0, 0, 0, 0, 0]
    clean = np.full((21, 2), 100.0, np.float32)       # the oracle (like the reference) raises on NaN / inf joints
    assert np.array_equal(t[0, 5], R.generate_target(clean)[0][5]) and w[0, 5, 0] == 1


def test_pck_nan_and_zero_diagonal_semantics(ops):
    """correct = total - #(ratio > T): NaN ratios (0/0) count as correct, inf (d/0) as wrong (Appendix B.6)."""
    gt = np.zeros((2, 21, 2), np.float32); gt[1] = np.random.default_rng(1).uniform(0, 200, (21, 2))
    pred = gt.copy(); pred[0, 5:] += 3.0            # sample 0: zero diagonal; joints 0-4 exact (0/0), rest d/0 = inf
    c, s = ops.new_metric_state("cuda")
    ops.pck_epe_batch(cu(pred), cu(gt), 0.2, "proportion", False, c, s)
    assert (c[1].item() - c[0].item(), c[1].item()) == R.pck_2d_counts(pred, gt, 0.2, "proportion")
    assert c[0].item() == 16


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 4), st.integers(1, 6), st.integers(2, 40), st.integers(2, 40), st.integers(0, 2 ** 31 - 1))
def test_decode_random_shapes_with_ties_property(n, j, h, w, seed):
    from lighthand_b200 import ops as _ops
    rng = np.random.default_rng(seed)
    hm = rng.integers(-2, 3, (n, j, h, w)).astype(np.float32)
    if seed % 3 == 0:
        hm[rng.integers(n), rng.integers(j), rng.integers(h), rng.integers(w)] = np.nan
    coords, maxvals, idx = _ops.decode(cu(hm), 1.0)
    p, m = R.get_max_preds(hm)
    assert np.array_equal(coords.cpu().numpy(), p) and np.array_equal(maxvals.cpu().numpy(), m, equal_nan=True)
    assert np.array_equal(idx.cpu().numpy(), np.argmax(hm.reshape(n, j, -1), 2))


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 3), st.sampled_from([(16, 16), (24, 40), (56, 56), (64, 48), (7, 9)]), st.sampled_from([1, 2, 3]),
       st.integers(0, 2 ** 31 - 1))
def test_targets_random_geometry_property(n, hw, sigma, seed):
    from lighthand_b200 import ops as _ops
    rng = np.random.default_rng(seed)
    h, w = hw
    joints = rng.uniform(-30, 4 * max(h, w) + 30, (n, 21, 2)).astype(np.float32)
    tgt, wt = _ops.gen_targets(cu(joints), (h, w), (4.0, 4.0), sigma)
    for i in range(n):
        want, ww = R.generate_target(joints[i], 21, (h, w), (4.0, 4.0), sigma)
        assert np.array_equal(tgt[i].cpu().numpy(), want) and np.array_equal(wt[i].cpu().numpy(), ww)
