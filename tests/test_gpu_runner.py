"""the train / valid loops (src/utils/method.py:156-287) device-resident.  The reference's
per-iteration work is replayed with the CPU oracle (model on the GPU, then .cpu() + numpy, as the reference does)
and compared with lighthand_b200.utils.method.Runner_t / Runner_v on the same model and batches."""
import copy
import types

import numpy as np
import pytest
import torch

from oracle import restate as R

pytestmark = pytest.mark.gpu


class TinyNet(torch.nn.Module):
    """Stand-in for the unchanged backbone: [N,3,64,64] -> [N,21,16,16]."""

    def __init__(self):
        super().__init__()
        self.c1 = torch.nn.Conv2d(3, 8, 3, stride=2, padding=1)
        self.c2 = torch.nn.Conv2d(8, 21, 3, stride=2, padding=1)

    def forward(self, x):
        return self.c2(torch.relu(self.c1(x)))


def make_batches(seed, nb=3, n=6):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(nb):
        images = torch.from_numpy(rng.standard_normal((n, 3, 64, 64)).astype(np.float32))
        joints = rng.uniform(4, 60, (n, 21, 2)).astype(np.float32)
        heat = np.stack([R.generate_target(j, 21, (16, 16), (4, 4), 2)[0] for j in joints])
        out.append((images, torch.from_numpy(joints), torch.from_numpy(heat)))
    return out


def reference_valid_loop(model, batches):
    """src/utils/method.py:219-287 with the oracle's functions."""
    loss_m, pck_m, epe_m = R.AverageMeter(), R.AverageMeter(), R.AverageMeter()
    model.eval()
    with torch.no_grad():
        for images, gt, heat in batches:
            pred = model(images.cuda())
            loss = R.joints_mse_loss(pred.cpu(), heat)
            pj, _ = R.get_max_preds(np.array(pred.detach().cpu()))
            pj = pj * 4
            loss_m.update(loss.item(), images.size(0))
            pck_m.update(R.pck_2d(pj, gt.numpy(), 0.2, "proportion"), images.size(0))
            s, c = R.epe_train(pj, gt.numpy())
            epe_m.update_p(s, c)
    return loss_m.avg, pck_m.avg * 100, epe_m.avg


def test_runner_v_matches_reference_loop():
    from lighthand_b200.utils.method import Runner_t, Runner_v
    torch.manual_seed(0)
    model = TinyNet().cuda()
    batches = make_batches(1)
    want_loss, want_pck, want_epe = reference_valid_loop(model, batches)
    args = types.SimpleNamespace(logging_steps=1, epoch=1, count=0, name="m/d/t")
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    tr = Runner_t(args, model, 0, batches, "TRAIN", None, None, 18, 3, 0, None, opt)
    for regen in (False, True):
        tr.regenerate_targets = regen
        rv = Runner_v(tr, model, batches, "VALID", 1.0)
        loss, count, pck, _bt = rv.run(0.0)
        assert loss == pytest.approx(want_loss, rel=1e-5)
        assert rv.epe_losses.avg == pytest.approx(want_epe, rel=1e-6)
        assert pck == pytest.approx(want_pck, rel=1e-12)          # integer counts -> identical ratio
        assert count == 0


def test_runner_t_trains_like_autograd():
    from lighthand_b200.utils.method import Runner_t
    torch.manual_seed(1)
    model = TinyNet().cuda()
    ref = copy.deepcopy(model)
    batches = make_batches(2)
    # reference-style epoch: JointsMSELoss + loss.backward() (autograd), SGD
    opt_ref = torch.optim.SGD(ref.parameters(), lr=0.5)
    ref.train()
    losses = []
    for images, _gt, heat in batches:
        pred = ref(images.cuda())
        loss = R.joints_mse_loss(pred, heat.cuda())
        losses.append(loss.item())
        opt_ref.zero_grad(); loss.backward(); opt_ref.step()
    args = types.SimpleNamespace(logging_steps=1, epoch=1, count=0, name="m/d/t")
    opt = torch.optim.SGD(model.parameters(), lr=0.5)
    tr = Runner_t(args, model, 0, batches, "TRAIN", None, None, 18, 3, 0, None, opt)
    m, o, bt = tr.run(0.0)
    assert m is model and o is opt and bt.count == 3
    assert tr.log_losses.avg == pytest.approx(float(np.mean(losses)), rel=1e-4)
    for a, b in zip(model.parameters(), ref.parameters()):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-6)
    # decoded joints of the last batch are on the device and already scaled by 4
    assert tr.pred_joint.is_cuda and tuple(tr.pred_joint.shape) == (6, 21, 2)


def test_pred_store_wire_format_and_eval_sweep(tmp_path, monkeypatch):
    """evaluation.json schema (src/utils/argparser.py:246-281, src/utils/dir.py:13-22) and the
    pck_eval_*.txt writer (src/tools/wearable_eval_2d.py:64-79)."""
    import json
    import os
    from lighthand_b200.utils import argparser as A
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(3)
    model = TinyNet().cuda().eval()
    rng = np.random.default_rng(5)
    names = list(A.POSE_CATEGORIES)
    loader = []
    for b in range(3):
        n = 5
        images = torch.from_numpy(rng.standard_normal((n, 3, 64, 64)).astype(np.float32))
        gt = rng.uniform(4, 60, (n, 21, 3)).astype(np.float32)
        gt[..., 2] = rng.uniform(size=(n, 21)) > 0.2
        anno = [names[int(k)] for k in rng.integers(0, 4, n)]
        loader.append((images, torch.from_numpy(gt), (anno, torch.arange(n))))
    bar = types.SimpleNamespace(update=lambda k: None)
    args = types.SimpleNamespace(name="tiny/run")
    meta = A.pred_store(args, loader, model, bar)
    # the reference's loop, restated with the oracle
    want = {name: {"bb": [], "pred": [], "gt": []} for name in names}
    with torch.no_grad():
        for images, gt, (anno, _i) in loader:
            pred = np.array(model(images.cuda()).detach().cpu())
            pj, _ = R.get_max_preds(pred)
            pj = torch.tensor(pj * 4)
            bb = R.bbox_diag_store(gt)
            for i, name in enumerate(anno):
                want[name]["bb"].append(bb[i]); want[name]["pred"].append(pj[i].tolist()); want[name]["gt"].append(gt[i].tolist())
    assert meta == want
    on_disk = json.load(open(os.path.join("output", "tiny/run", "evaluation.json")))
    assert isinstance(on_disk, list) and len(on_disk) == 1 and on_disk[0] == want
    out, _ = A.pred_eval(args, [0, 30], bar, "mm")
    ref = R.pred_eval_meta(want, [0, 30], "mm")
    for k in ref:
        assert np.array_equal(np.asarray(out[k][2]), np.asarray(ref[k][2])) and out[k][0] == pytest.approx(ref[k][0], rel=1e-12)
    A.write_pck_eval("pck_eval_tiny_mm_30.txt", [[out, "tiny/run"]])
    lines = open("pck_eval_tiny_mm_30.txt").read().splitlines()
    assert len(lines) == len(out)
    first = lines[0].split(";")
    k0 = next(iter(out))
    assert first[0] == k0 and first[1] == "tiny/run" and first[2] == "{:.2f}".format(out[k0][0]) and len(first) == 4 + 100 + 1


def test_gpu_collate_eval_phase_feeds_pred_store(tmp_path, monkeypatch):
    """eval_set.__getitem__ (phase "eval", src/tools/dataset.py:257-300) as a device collate: item shape, key-point
    scaling, and the batch goes straight into pred_store (src/utils/argparser.py:246-281)."""
    from torchvision import transforms
    from lighthand_b200.tools.dataset import GpuCollate
    from lighthand_b200.utils import argparser as A
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(12)
    names = list(A.POSE_CATEGORIES)
    samples = []
    for i in range(6):
        img = rng.integers(0, 256, (224, 224, 3), dtype=np.uint8)
        joint = rng.uniform(0.05, 0.95, (21, 3)).astype(np.float32)
        vis = (rng.uniform(size=21) < 0.8).astype(np.float32)
        samples.append((img, joint, vis, names[i % 4], i))
    coll = GpuCollate(image_size=64, phase="eval")
    images, joint_2d_v, (pose, idx) = coll(samples)
    trans = transforms.Compose([transforms.ToTensor(), transforms.Resize((64, 64), antialias=False),
                                transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    want_img = torch.stack([trans(s[0]) for s in samples])
    assert torch.allclose(images.cpu(), want_img, rtol=1e-5, atol=1e-5)
    want_j = torch.stack([torch.cat([torch.tensor(s[1])[:, :2] * 64, torch.tensor(s[2])[:, None]], 1) for s in samples])
    assert torch.equal(joint_2d_v.cpu(), want_j) and pose == [s[3] for s in samples] and idx == list(range(6))
    torch.manual_seed(1)
    model = TinyNet().cuda().eval()
    bar = types.SimpleNamespace(update=lambda k: None)
    meta = A.pred_store(types.SimpleNamespace(name="tiny/eval"), [(images, joint_2d_v, (pose, idx))], model, bar)
    assert sum(len(v["bb"]) for v in meta.values()) == 6
    for k, name in enumerate(pose):
        assert meta[name]["gt"][k // 4] == want_j[k].tolist()
