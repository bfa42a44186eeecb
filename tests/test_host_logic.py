"""Host-side logic that needs no GPU: the batched matrix builder against cv2, sharding arithmetic,
the curve finalisation math, and the metric all-reduce over gloo with world_size 2."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, golden
from oracle import restate as R


def test_batched_affine_builder_is_cv2_bitexact():
    import cv2
    from lighthand_b200.utils.preprocessing import (_affine_from_3pts, batch_patch_matrices, gen_trans_from_patch_batch,
                                                    gen_trans_from_patch_cv)
    rng = np.random.default_rng(0)
    n = 500
    cx, cy = rng.uniform(20, 230, n), rng.uniform(20, 230, n)
    w, h = rng.uniform(40, 250, n), rng.uniform(40, 250, n)
    scale = np.clip(rng.standard_normal(n), -1, 1) * 0.25 + 1
    rot = np.where(rng.uniform(size=n) < 0.6, np.clip(rng.standard_normal(n), -2, 2) * 45, 0.0)
    for inv in (False, True):
        got = gen_trans_from_patch_batch(cx, cy, w, h, 256, 256, scale, rot, inv=inv)
        for i in range(n):
            want = R.gen_trans_from_patch_cv(cx[i], cy[i], w[i], h[i], 256, 256, scale[i], rot[i], inv=inv)
            assert np.array_equal(got[i], want), (i, inv)
    assert np.array_equal(gen_trans_from_patch_cv(100, 120, 200, 180, 256, 256, 1.1, 33.0), R.gen_trans_from_patch_cv(100, 120, 200, 180, 256, 256, 1.1, 33.0))
    # raw 3-point solver in float64 against cv2
    src = rng.uniform(0, 256, (50, 3, 2)).astype(np.float32)
    dst = rng.uniform(0, 256, (50, 3, 2)).astype(np.float32)
    got = _affine_from_3pts(src, dst)
    for i in range(50):
        assert np.array_equal(got[i], cv2.getAffineTransform(src[i], dst[i]))
    # golden matrices of the reference
    g = golden("warp.npz")
    t, it = batch_patch_matrices(g["bbox"], g["flip"], g["scale"], g["rot"], g["imgs"].shape[1:3], g["out_hw"])
    assert np.array_equal(t, g["trans"]) and np.array_equal(it, g["inv_trans"])


def test_shard_range_partitions_exactly():
    from lighthand_b200.distributed import shard_range
    for total, ws in ((99792, 8), (99792, 4), (99792, 2), (1024, 8), (10, 3), (5, 8), (0, 4)):
        spans = [shard_range(total, r, ws) for r in range(ws)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert shard_range(99792, 3, 8) == (3 * 12474, 4 * 12474)


def test_eval_thresholds_match_reference_formulae():
    from lighthand_b200.utils.argparser import MM_PER_PX, MM_PRED_EVAL, eval_thresholds
    assert np.array_equal(eval_thresholds([0, 30], "mm"), np.linspace(0, 30, 101)[1:] * 2.83464567)
    assert np.array_equal(eval_thresholds([0, 50], "mm", MM_PER_PX), np.linspace(0, 50, 101)[1:] * 3.7795275591)
    assert np.array_equal(eval_thresholds([0.1, 0.3], "pckb"), np.linspace(0.1, 0.3, 100))
    assert np.array_equal(eval_thresholds([0, 30], "mm"), R.eval_thresholds([0, 30], "mm"))
    with pytest.raises(AssertionError):
        eval_thresholds([0, 1], "nope")


def test_average_meter_shim():
    from lighthand_b200.utils.metric_logger import AverageMeter
    g = golden("metrics.npz")["meter"]
    m = AverageMeter(); m.update(0.5, 4); m.update(0.25, 12)
    mp = AverageMeter(); mp.update_p(10.0, 4); mp.update_p(30.0, 16)
    assert [m.sum, m.count, m.avg, mp.sum, mp.count, mp.avg] == list(g)


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from lighthand_b200.distributed import all_reduce_metrics, shard_range, world
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
rank, ws = world()
lo, hi = shard_range(99792, rank, ws)
counts = torch.tensor([lo, hi - lo, 7 * (rank + 1), 1], dtype=torch.int64)
hist = torch.arange(12, dtype=torch.int64).view(3, 4) * (rank + 1)
sums = torch.tensor([0.5 * (rank + 1), 1e-3, 2.0, 0.0], dtype=torch.float64)
all_reduce_metrics([counts, hist, sums])
assert counts.tolist() == [49896, 99792, 21, 2], counts
assert hist.tolist() == (torch.arange(12).view(3, 4) * 3).tolist()
assert sums.tolist() == [1.5, 2e-3, 4.0, 0.0]
dist.destroy_process_group()
print("ok", rank)
'''


def test_metric_allreduce_gloo_world2(tmp_path):
    import subprocess
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, env={**os.environ, "OMP_NUM_THREADS": "1"}) for r in range(2)]
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out.decode()[-2000:]


def test_frei_crop_descriptors_follow_the_reference_bookkeeping():
    """utils.image_ops.crop_descriptors == the integer window arithmetic of crop() (oracle.frei_crop_geometry),
    and the packed record has the layout include/lighthand_b200.h declares."""
    import ctypes
    from lighthand_b200 import _lib
    from lighthand_b200.utils import image_ops
    from oracle import restate as R
    assert ctypes.sizeof(_lib.FreiDesc) == 112 and _lib.FreiDesc.iM.offset == 24 and _lib.FreiDesc.scratch_off.offset == 104
    rng = np.random.default_rng(21)
    n = 40
    centers = 112 + rng.uniform(-80, 80, (n, 2))
    scales = rng.uniform(0.4, 1.8, n)
    rots = np.where(rng.uniform(size=n) < 0.5, 0.0, rng.uniform(-180, 180, n))
    pns = rng.uniform(0.6, 1.4, (n, 3))
    recs, scratch, max_wh = image_ops.crop_descriptors(centers, scales, [224, 224], rots, pns)
    off = 0
    for i in range(n):
        g = R.frei_crop_geometry(centers[i], float(scales[i]), [224, 224], float(rots[i]), (224, 224))
        r = recs[i]
        assert (r.ulx, r.uly) == (int(g["ul"][0]), int(g["ul"][1])) and (r.ch, r.cw) == g["shape"]
        assert r.pad == g["pad"] and r.rotated == int(rots[i] != 0)
        assert np.allclose([r.pn[0], r.pn[1], r.pn[2]], pns[i].astype(np.float32))
        if r.rotated:
            want = R.cv2_invert_affine(R.cv2_rotation_matrix(((r.cw - 1) * 0.5, (r.ch - 1) * 0.5), float(rots[i]))).ravel()
            assert np.array_equal(np.array(list(r.iM)), want) and r.scratch_off == off
            off += 3 * (r.cw - 2 * r.pad) * (r.ch - 2 * r.pad)
    assert scratch == off and max_wh[0] > 0 and max_wh[1] > 0
    for pt in ([1, 1], [100.5, 30.25]):
        assert np.array_equal(image_ops.transform(pt, centers[0], scales[0], [224, 224], invert=1, rot=rots[1]),
                              R.frei_transform(pt, centers[0], scales[0], [224, 224], invert=1, rot=rots[1]))


def test_color_jitter_draw_follows_torchvision_stream():
    """tools.dataset.color_jitter_params consumes torch's generator exactly like ColorJitter.get_params
    (randperm(4), then brightness / contrast / saturation / hue uniforms), sample by sample."""
    import ctypes
    from torchvision import transforms
    from lighthand_b200 import _lib
    from lighthand_b200.tools.dataset import color_jitter_params
    assert ctypes.sizeof(_lib.JitterParams) == 52
    cj = transforms.ColorJitter(brightness=0.5, contrast=0.5, saturation=0.5, hue=0.5)
    flags = [True, False, True, True]
    torch.manual_seed(77)
    want = [cj.get_params(cj.brightness, cj.contrast, cj.saturation, cj.hue) if f else None for f in flags]
    torch.manual_seed(77)
    recs = color_jitter_params(flags)
    for r, w, f in zip(recs, want, flags):
        assert r.apply == int(f)
        if not f:
            assert list(r.order) == [-1] * 4
            continue
        fn_idx, b, c, s, h = w
        assert list(r.order) == [int(v) for v in fn_idx]
        assert np.allclose(list(r.factor), np.array([b, c, s, h], np.float32), rtol=0, atol=0)
        assert np.allclose(list(r.one_minus)[:3], np.array([1.0 - b, 1.0 - c, 1.0 - s]).astype(np.float32), rtol=0, atol=0)
