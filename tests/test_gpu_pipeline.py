"""End-to-end parity: the reference-shaped Python surface and the two pipeline objects against the CPU
oracle, at small sizes (oracle finishes in seconds) and through size-independent properties at
BASELINE.json's full batch sizes."""
import json
import os
import types

import numpy as np
import pytest
import torch

from conftest import golden
from oracle import restate as R

pytestmark = pytest.mark.gpu


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def synth(n, S, H, sigma, seed, src=None):
    """Synthetic batch as SURVEY 8(d) describes it (host numpy)."""
    rng = np.random.default_rng(seed)
    src = src or S
    imgs = rng.integers(0, 256, (n, src, src, 3), dtype=np.uint8)
    joints = rng.uniform(0, S, (n, 21, 2)).astype(np.float32)
    off = rng.uniform(size=(n, 21)) < 0.05
    joints[off] = rng.uniform(-S / 8, 9 * S / 8, (int(off.sum()), 2)).astype(np.float32)
    side = 0.8 * src
    bbox = np.tile(np.array([(src - side) / 2, (src - side) / 2, side, side]), (n, 1))
    scale = np.clip(rng.standard_normal(n), -1, 1) * 0.25 + 1
    rot = np.where(rng.uniform(size=n) <= 0.6, np.clip(rng.standard_normal(n), -2, 2) * 45, 0.0)
    flip = rng.uniform(size=n) <= 0.5
    return imgs, joints, bbox, scale, rot, flip


def oracle_step(imgs, joints, bbox, scale, rot, flip, pred, S, H, sigma, T=0.2):
    """The reference's per-iteration work, restated on the CPU."""
    n = imgs.shape[0]
    stride = S / H
    inp, tgt = [], []
    for i in range(n):
        chw, _, _ = R.warp_normalize_sample(imgs[i], bbox[i], bool(flip[i]), float(scale[i]), float(rot[i]), (S, S))
        inp.append(chw)
        tgt.append(R.generate_target(joints[i], 21, (H, H), (stride, stride), sigma)[0])
    inp, tgt = np.stack(inp), np.stack(tgt)
    loss, grad = R.joints_mse_loss_and_grad(pred, tgt)
    coords, maxvals, idx = R.decode_final(pred, stride)
    correct, total = R.pck_2d_counts(coords, joints, T, "proportion")
    esum, ecnt = R.epe_train(coords, joints)
    return dict(input=inp, target=tgt, loss=loss, grad=grad, coords=coords, maxvals=maxvals, idx=idx,
                correct=correct, total=total, esum=esum, ecnt=ecnt)


@pytest.mark.parametrize("S,H,sigma,n", [(256, 64, 2, 8), (224, 56, 2, 5), (512, 128, 3, 3)])
def test_device_pipeline_vs_oracle(S, H, sigma, n):
    from lighthand_b200 import ops
    from lighthand_b200.pipeline import HeatmapPipeline
    from lighthand_b200.utils.preprocessing import batch_patch_matrices
    imgs, joints, bbox, scale, rot, flip = synth(n, S, H, sigma, seed=S + n)
    trans, inv = batch_patch_matrices(bbox, flip, scale, rot, (S, S), (S, S))
    rng = np.random.default_rng(1)
    stride = S / H
    tgt = np.stack([R.generate_target(j, 21, (H, H), (stride, stride), sigma)[0] for j in joints])
    pred = (tgt + 0.1 * rng.standard_normal(tgt.shape)).astype(np.float32)
    want = oracle_step(imgs, joints, bbox, scale, rot, flip, pred, S, H, sigma)

    for use_graph in (False, True):
        pipe = HeatmapPipeline(n, (S, S), (S, S), (H, H), sigma=sigma, materialize_target=True, use_graph=use_graph,
                               exact_norm=not use_graph)
        d = [cu(imgs), cu(trans.astype(np.float64)), cu(joints), cu(pred), cu(flip.astype(np.uint8))]
        for _rep in range(2):
            out = pipe.step(d[0], d[1], d[2], d[3], flip=d[4])
        torch.cuda.synchronize()
        if use_graph:   # one-FMA normalisation: within 2 ulp of torch's three roundings
            assert np.abs(out["input"].cpu().numpy() - want["input"]).max() <= 5e-7
        else:           # table normalisation: bit-identical
            assert np.array_equal(out["input"].cpu().numpy(), want["input"])
        assert np.array_equal(out["target"].cpu().numpy(), want["target"])
        assert out["loss"].item() == pytest.approx(want["loss"], rel=1e-5)
        np.testing.assert_allclose(out["grad"].cpu().numpy(), want["grad"], rtol=1e-5, atol=1e-12)
        assert np.array_equal(out["coords"].cpu().numpy(), want["coords"])
        assert np.array_equal(out["argidx"].cpu().numpy(), want["idx"])
        m = pipe.metrics(all_reduce=False)
        assert m["n_images"] == 2 * n and m["n_total"] == 2 * want["total"]
        assert m["n_total"] - m["n_wrong"] == 2 * want["correct"]
        assert m["epe_count"] == 2 * want["ecnt"] and m["epe_sum"] == pytest.approx(2 * want["esum"], rel=1e-6)
        assert m["loss"] == pytest.approx(want["loss"], rel=1e-5)


def test_host_pipeline_c_abi_vs_device_pipeline():
    from lighthand_b200.pipeline import HeatmapPipeline, HostPipeline
    from lighthand_b200.utils.preprocessing import batch_patch_matrices
    S, H, n = 256, 64, 16
    imgs, joints, bbox, scale, rot, flip = synth(n, S, H, 2, seed=77)
    trans, _ = batch_patch_matrices(bbox, flip, scale, rot, (S, S), (S, S))
    M = np.ascontiguousarray(trans.astype(np.float64))
    pred = np.random.default_rng(3).standard_normal((n, 21, H, H)).astype(np.float32)
    fl = np.ascontiguousarray(flip.astype(np.uint8))
    pipe = HeatmapPipeline(n, materialize_target=True)
    ref = pipe.step(cu(imgs), cu(M), cu(joints), cu(pred), flip=cu(fl))
    hp = HostPipeline(n, materialize_target=True)
    outs = [hp.step(imgs, M, joints, flip=fl, pred=pred) for _ in range(3)]   # both slots + reuse
    out = outs[-1]
    assert out["loss"][0] == ref["loss"].item()
    assert np.array_equal(out["coords"], ref["coords"].cpu().numpy())
    assert np.array_equal(out["maxvals"], ref["maxvals"].cpu().numpy()[..., 0])
    assert torch.equal(hp.input, ref["input"]) and torch.equal(hp.grad, ref["grad"]) and torch.equal(hp.target, ref["target"])
    m = pipe.metrics(all_reduce=False)
    assert [int(v) for v in out["counts"]] == [3 * m["n_wrong"], 3 * m["n_total"], 3 * m["epe_count"], 3 * n]
    assert out["sums"][0] == pytest.approx(3 * m["epe_sum"], rel=1e-12)
    # asynchronous double-buffered use
    a = hp.submit(imgs, M, joints, flip=fl)          # pred=None: device-resident pred stays
    b = hp.submit(imgs, M, joints, flip=fl)
    hp.wait(a[0]); hp.wait(b[0])
    assert np.array_equal(b[1]["coords"], out["coords"])
    hp.reset_metrics()
    assert int(hp.step(imgs, M, joints, flip=fl)["counts"][3]) == n
    hp.close()


@pytest.mark.parametrize("S,H,sigma,n", [(256, 64, 2, 256), (512, 128, 3, 128)])
def test_full_size_properties(S, H, sigma, n):
    """BASELINE.json sizes (batch 256 @256/64; 128/GPU @512/128 sigma 3) through properties that need no oracle run."""
    from lighthand_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    stride = S / H
    joints = torch.rand((n, 21, 2), device="cuda", generator=g) * (S - 8 * stride) + 4 * stride   # well inside
    tgt, w = ops.gen_targets(joints, (H, H), (stride, stride), sigma)
    assert bool((w == 1).all())
    # decoding the target returns the rounded joint: mu = int(j/stride + .5)
    coords, maxvals, idx = ops.decode(tgt, 1.0)
    mu = (joints / stride + 0.5).to(torch.int32).float()
    assert torch.equal(coords, mu) and bool((maxvals == 1).all())
    # loss(pred == target) == 0 and grad == 0; loss is quadratic: loss(t + c) = c^2 / 2
    out = ops.heatmap_fused(tgt, joints, sigma=sigma, feat_stride=(stride, stride), out_scale=stride)
    assert out["loss"].item() == 0.0 and float(out["grad"].abs().max()) == 0.0
    assert torch.equal(out["coords"], mu * stride)
    out = ops.heatmap_fused(tgt + 0.5, joints, sigma=sigma, feat_stride=(stride, stride))
    assert out["loss"].item() == pytest.approx(0.125, rel=1e-6)
    assert float((out["grad"] - 0.5 / tgt.numel()).abs().max()) <= 1e-5 * 0.5 / tgt.numel()
    # PCK of a perfect prediction = 1, EPE = 0; counts add over shards
    counts, sums = ops.new_metric_state("cuda")
    ops.pck_epe_batch(joints, joints, 0.2, "proportion", False, counts, sums)
    assert counts.cpu().tolist() == [0, n * 21, n * 19, n] and sums[0].item() == 0.0
    # identity warp reproduces the image; flipping twice is the identity
    imgs = torch.randint(0, 256, (n, S, S, 3), device="cuda", dtype=torch.uint8, generator=g)
    eye = torch.tensor([[1.0, 0, 0], [0, 1.0, 0]], dtype=torch.float64, device="cuda").repeat(n, 1, 1)
    out = ops.warp_normalize(imgs, eye, (S, S), mean=None, std=None)
    assert torch.equal(out, imgs.permute(0, 3, 1, 2).float())
    fl = torch.ones(n, dtype=torch.uint8, device="cuda")
    out_f = ops.warp_normalize(imgs, eye, (S, S), flip=fl, mean=None, std=None)
    assert torch.equal(out_f, imgs.flip(2).permute(0, 3, 1, 2).float())


def test_reference_call_surface_shims():
    """The drop-in functions keep the reference's signatures and return types."""
    import lighthand_b200.utils.loss as L
    from lighthand_b200.tools.dataset import GpuCollate, generate_target
    from lighthand_b200.utils.dataset_loader import GenerateHeatmap
    from lighthand_b200.utils.preprocessing import generate_patch_image
    g = golden("loss.npz")
    target = np.stack([R.generate_target(j)[0] for j in g["joints"]])
    # JointsMSELoss: nn.Module, autograd backward == reference autograd
    p = cu(g["pred"]).requires_grad_(True)
    loss = L.JointsMSELoss(use_target_weight=False)(p, cu(target), None)
    (loss * 2).backward()
    assert loss.item() == pytest.approx(float(g["loss"]), rel=1e-5)
    np.testing.assert_allclose(p.grad.cpu().numpy(), 2 * g["grad"], rtol=1e-5, atol=1e-12)
    # get_max_preds: numpy in -> numpy out
    d = golden("decode.npz")
    preds, maxvals = L.get_max_preds(d["hm"])
    assert isinstance(preds, np.ndarray) and np.array_equal(preds, d["preds"]) and np.array_equal(maxvals, d["maxvals"], equal_nan=True)
    with pytest.raises(AssertionError):
        L.get_max_preds([1, 2, 3])
    # metrics: python floats, CPU tensors accepted like the reference
    m = golden("metrics.npz")
    tp, tg, tg3 = torch.tensor(m["pred"]), torch.tensor(m["gt"]), torch.tensor(m["gt3"])
    assert L.PCK_2d_loss(tp, tg, 0.2, "proportion") == m["pck_prop"][2]
    assert L.PCK_2d_loss_visible(tp, tg3, 0.2, "proportion") == m["pck_vis_prop"][2]
    pl, al = L.PCK_2d_loss_list(tp, tg, T_list=[0.1, 0.2], threshold="proportion", auc_t=[0.05, 0.3])
    assert np.array_equal(np.array(pl), m["pck_list"]) and np.array_equal(np.array(al), m["pck_list_auc"])
    (s, c), _ = L.EPE_train(tp, tg)
    assert c == m["epe_train"][1] and s == pytest.approx(m["epe_train"][0], rel=1e-6)
    (s, c), _ = L.EPE(tp, tg3)
    assert c == m["epe_vis"][1] and s == pytest.approx(m["epe_vis"][0], rel=1e-6)
    with pytest.raises(AssertionError):
        L.PCK_2d_loss(tp, tg, 0.2, "bogus")
    # generate_target / GenerateHeatmap: one sample, CPU in -> CPU out
    t = golden("targets.npz")
    out = generate_target(torch.tensor(t["joints"][0]))
    assert isinstance(out, torch.Tensor) and not out.is_cuda and np.array_equal(out.numpy(), t["target"][0])
    hg = GenerateHeatmap(64, 21)(t["joints"][0] / 4)
    assert isinstance(hg, np.ndarray) and np.array_equal(hg, t["hourglass_np"][0])
    # generate_patch_image: same triple as the reference
    w = golden("warp.npz")
    for i in (0, 3, 4):
        patch, trans, inv = generate_patch_image(w["imgs"][i], w["bbox"][i], bool(w["flip"][i]), float(w["scale"][i]),
                                                 float(w["rot"][i]), tuple(int(v) for v in w["out_hw"]))
        assert np.array_equal(trans, w["trans"][i]) and np.array_equal(inv, w["inv_trans"][i])
        assert np.array_equal(patch, w["patch_u8src"][i])
    # GpuCollate == __getitem__ + default collate (resize mode)
    r = golden("resize.npz")
    col = GpuCollate(image_size=64, heatmap_size=16, sigma=2)
    joints56 = np.random.default_rng(2).uniform(0, 56, (3, 21, 2)).astype(np.float32)
    images, j2d, heat = col([(r["imgs"][i], joints56[i]) for i in range(3)])
    np.testing.assert_allclose(images.cpu().numpy(), r["out"], rtol=1e-5, atol=1e-5)
    jw = (torch.tensor(joints56) * (64 / 56)).numpy()
    assert np.array_equal(j2d.cpu().numpy(), jw)
    for i in range(3):
        assert np.array_equal(heat[i].cpu().numpy(), R.generate_target(jw[i], 21, (16, 16), (4.0, 4.0), 2)[0])


def test_pred_eval_reads_pred_store_json(tmp_path, monkeypatch):
    from conftest import GOLDEN
    from lighthand_b200.utils import argparser as A
    g = golden("eval.npz")
    with open(os.path.join(GOLDEN, "eval_meta.json")) as f:
        metas = json.load(f)
    monkeypatch.chdir(tmp_path)
    os.makedirs("output/run1"); os.makedirs("final_model/run1")
    json.dump([metas["eval"]], open("output/run1/evaluation.json", "w"))       # dump() wraps in a list
    json.dump([metas["test"]], open("final_model/run1/test.json", "w"))
    bar = types.SimpleNamespace(n=0)
    bar.update = lambda k: setattr(bar, "n", bar.n + k)
    out, _ = A.pred_eval(types.SimpleNamespace(name="run1"), [0, 30], bar, "mm")
    assert bar.n == 4 and np.array_equal(out["mean_auc"][2], g["mm30/mean_auc/pck_t"])
    auc, mean_px, _ = A.pred_test(types.SimpleNamespace(name="run1"), [0.1, 0.3], bar, "pckb")
    assert auc == pytest.approx(float(g["test/pckb/auc"]), rel=1e-12)


def test_overlapped_pipeline_is_bit_identical_to_the_serial_step():
    """data_stage / loss_stage (K1 of batch i+1 on the data stream while the loss stage of batch i runs) must give,
    batch by batch, exactly what the serial ``step`` gives -- the two launches touch disjoint buffers."""
    from lighthand_b200.pipeline import HeatmapPipeline
    from lighthand_b200.utils.preprocessing import batch_patch_matrices
    S, H, n, nb = 256, 64, 32, 5
    batches = []
    for b in range(nb):
        imgs, joints, bbox, scale, rot, flip = synth(n, S, H, 2, seed=500 + b)
        trans, _ = batch_patch_matrices(bbox, flip, scale, rot, (S, S), (S, S))
        pred = np.random.default_rng(b).standard_normal((n, 21, H, H)).astype(np.float32)
        batches.append([cu(imgs), cu(trans.astype(np.float64)), cu(joints), cu(pred), cu(flip.astype(np.uint8))])
    keys = ("input", "loss", "grad", "coords", "maxvals", "argidx", "target")
    serial = HeatmapPipeline(n, materialize_target=True)
    want = []
    for d in batches:
        out = serial.step(d[0], d[1], d[2], d[3], flip=d[4])
        want.append({k: out[k].clone() for k in keys})
    m_want = serial.metrics(all_reduce=False)

    pipe = HeatmapPipeline(n, materialize_target=True)
    pipe.data_stage(batches[0][0], batches[0][1], batches[0][4])
    for b, d in enumerate(batches):
        nxt = batches[b + 1] if b + 1 < nb else None
        out = pipe.step_overlapped(nxt[0] if nxt else None, nxt[1] if nxt else None, d[2], d[3],
                                   next_flip=nxt[4] if nxt else None)
        for k in keys:
            assert torch.equal(out[k], want[b][k]), (b, k)
    assert pipe.metrics(all_reduce=False) == m_want

    # the pre-marshalled form the benchmark drives
    pipe2 = HeatmapPipeline(n, materialize_target=True)
    steps = [pipe2.prepare(d[0], d[1], d[2], d[3], flip=d[4]) for d in batches]
    steps[0].launch_data()
    for b in range(nb):
        if b + 1 < nb:
            steps[b + 1].launch_data()
        steps[b].launch_loss()
        assert torch.equal(pipe2.buffers["coords"], want[b]["coords"]) and torch.equal(pipe2.buffers["grad"], want[b]["grad"])
        assert torch.equal(pipe2.inputs[b % 2], want[b]["input"])
    assert pipe2.metrics(all_reduce=False) == m_want


def test_host_pipeline_split_submit_lets_the_backbone_run_between():
    """lh_pipeline_submit_inputs / lh_pipeline_submit_loss: a stand-in backbone enqueued on the loss stream between the
    two halves reads the slot's input and writes pred; results equal a device pipeline fed the same pred."""
    from lighthand_b200.pipeline import HeatmapPipeline, HostPipeline
    from lighthand_b200.utils.preprocessing import batch_patch_matrices
    S, H, n = 256, 64, 8
    imgs, joints, bbox, scale, rot, flip = synth(n, S, H, 2, seed=91)
    trans, _ = batch_patch_matrices(bbox, flip, scale, rot, (S, S), (S, S))
    M = np.ascontiguousarray(trans.astype(np.float64))
    fl = np.ascontiguousarray(flip.astype(np.uint8))
    hp = HostPipeline(n)
    slot, out, _keep = hp.submit_inputs(imgs, M, joints, flip=fl)
    with torch.cuda.stream(hp.loss_stream):
        inp = hp.input_slot(slot)
        hp.pred.copy_(torch.nn.functional.avg_pool2d(inp, 4).mean(1, keepdim=True).expand(-1, 21, -1, -1))   # "backbone"
    hp.submit_loss(slot)
    hp.wait(slot)
    pipe = HeatmapPipeline(n)
    ref_in = pipe.step(cu(imgs), cu(M), cu(joints), hp.pred.clone(), flip=cu(fl))
    assert torch.equal(hp.input_slot(slot), ref_in["input"])
    assert out["loss"][0] == ref_in["loss"].item()
    assert np.array_equal(out["coords"], ref_in["coords"].cpu().numpy())
    counts, sums = hp.read_metrics()
    m = pipe.metrics(all_reduce=False)
    assert [int(v) for v in counts] == [m["n_wrong"], m["n_total"], m["epe_count"], n]
    hp.close()


@pytest.mark.parametrize("S,H,sigma,n,train", [(256, 64, 2, 256, True), (224, 56, 2, 256, False), (512, 128, 3, 128, True)])
def test_benchmarked_variant_vs_oracle_at_full_baseline_batches(S, H, sigma, n, train):
    """The configuration bench.py times (materialize_target=False, one-FMA normalisation, TMA-staged K1, fused pass
    without the target store) against the CPU oracle at BASELINE.json's full batch sizes: C2 256 @ 256/64 train,
    C3 256 @ 224/56 eval, C5 128 @ 512/128 sigma 3 train."""
    from lighthand_b200.pipeline import HeatmapPipeline
    from lighthand_b200.utils.preprocessing import batch_patch_matrices
    imgs, joints, bbox, scale, rot, flip = synth(n, S, H, sigma, seed=7 * S + n)
    trans, _ = batch_patch_matrices(bbox, flip, scale, rot, (S, S), (S, S))
    stride = S / H
    rng = np.random.default_rng(S)
    tgt = np.stack([R.generate_target(j, 21, (H, H), (stride, stride), sigma)[0] for j in joints])
    pred = (tgt + 0.1 * rng.standard_normal(tgt.shape)).astype(np.float32)
    pipe = HeatmapPipeline(n, (S, S), (S, S), (H, H), sigma=sigma, train=train, materialize_target=False, exact_norm=False)
    st = pipe.prepare(cu(imgs), cu(trans.astype(np.float64)), cu(joints), cu(pred), flip=cu(flip.astype(np.uint8)))
    st.launch()
    torch.cuda.synchronize()
    # K1: the uint8 pixel is cv2's bit for bit; one-FMA normalisation within 2 ulp (values are O(1))
    got_in = pipe.input.cpu().numpy()
    for i in range(0, n, max(1, n // 16)):                    # the oracle's per-sample warp costs ~10 ms: check every 16th image
        want_chw, _, _ = R.warp_normalize_sample(imgs[i], bbox[i], bool(flip[i]), float(scale[i]), float(rot[i]), (S, S))
        assert np.abs(got_in[i] - want_chw).max() <= 5e-7, i
    loss, grad = R.joints_mse_loss_and_grad(pred, tgt) if train else (float(R.joints_mse_loss(pred, tgt)), None)
    coords, maxvals, idx = R.decode_final(pred, stride)
    correct, total = R.pck_2d_counts(coords, joints, 0.2, "proportion")
    esum, ecnt = R.epe_train(coords, joints)
    assert pipe.buffers["loss"].item() == pytest.approx(loss, rel=1e-5)
    if train:
        np.testing.assert_allclose(pipe.buffers["grad"].cpu().numpy(), grad, rtol=1e-5, atol=1e-12)
    assert np.array_equal(pipe.buffers["argidx"].cpu().numpy(), idx)
    assert np.array_equal(pipe.buffers["coords"].cpu().numpy(), coords)
    assert np.array_equal(pipe.buffers["maxvals"].cpu().numpy(), maxvals)
    m = pipe.metrics(all_reduce=False)
    assert (m["n_total"] - m["n_wrong"], m["n_total"], m["epe_count"]) == (correct, total, ecnt)
    assert m["epe_sum"] == pytest.approx(esum, rel=1e-6)
