"""Generates tests/golden/*.npz by running the REFERENCE's own functions (imported from
/root/reference through oracle/reference_bridge.py) on seeded synthetic inputs.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Every fixture stores its inputs next to the reference outputs, so the tests never depend on a
random generator reproducing a stream.  The reference ships no golden vectors of its own
(SURVEY.md section 4); these files are what pins oracle/restate.py -- and through it the CUDA
path -- to the reference.  Versions the outputs were produced with are recorded in VERSIONS.json.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reference_bridge as rb  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def edge_joints(rng, n, size=256.0):
    """[n,21,2] float32 joints in input pixels: uniform inside, plus off-image / border / negative /
    half-integer cases that exercise truncation toward zero and the window clipping."""
    j = rng.uniform(0, size, (n, 21, 2)).astype(np.float32)
    special = np.array([
        [-3.0, 10.0], [-2.0, -2.0], [-26.0, 100.0], [-30.0, 100.0], [-25.9, 50.0], [0.0, 0.0],
        [255.0, 255.0], [256.0, 128.0], [277.9, 30.0], [278.0, 30.0], [282.0, 30.0], [300.0, 300.0],
        [2.0, 2.0], [1.99, 5.99], [6.0, 10.0], [129.99, 130.0], [130.0, 126.0], [253.9, 1.0],
        [-1.0, 254.0], [127.5, 127.5], [1000.0, -1000.0]], dtype=np.float32)
    j[0] = special
    j[1, :, 0] = np.linspace(-40, 300, 21, dtype=np.float32)
    j[1, :, 1] = np.linspace(300, -40, 21, dtype=np.float32)
    return j


def make_targets(ns):
    rng = np.random.default_rng(11)
    joints = edge_joints(rng, 6)
    tgt = np.stack([ns.generate_target(torch.tensor(j)).numpy() for j in joints])
    # GenerateHeatmap(64, 21)(joint_2d / 4)  -- src/tools/dataset.py:296
    hg = np.stack([ns.GenerateHeatmap(64, 21)(torch.tensor(j) / 4) for j in joints])
    hg_np = np.stack([ns.GenerateHeatmap(64, 21)(j / 4) for j in joints])
    save("targets.npz", joints=joints, target=tgt, hourglass=hg.astype(np.float32),
         hourglass_np=hg_np.astype(np.float32))


def make_loss(ns):
    rng = np.random.default_rng(12)
    joints = rng.uniform(0, 256, (2, 21, 2)).astype(np.float32)
    target = np.stack([ns.generate_target(torch.tensor(j)).numpy() for j in joints])
    pred = (target + 0.1 * rng.standard_normal(target.shape)).astype(np.float32)
    weight = (rng.uniform(size=(2, 21, 1)) > 0.3).astype(np.float32)
    crit = ns.JointsMSELoss(use_target_weight=False)
    p = torch.tensor(pred, requires_grad=True)
    loss = crit(p, torch.tensor(target), None)
    loss.backward()
    # the reference ignores target_weight even when told to use it (Appendix B.1)
    loss_w = ns.JointsMSELoss(use_target_weight=True)(torch.tensor(pred), torch.tensor(target), torch.tensor(weight))
    save("loss.npz", joints=joints, pred=pred, weight=weight, loss=np.float32(loss.item()),
         loss_with_weight_arg=np.float32(loss_w.item()), grad=p.grad.numpy())


def make_decode(ns):
    rng = np.random.default_rng(13)
    hm = rng.standard_normal((3, 21, 16, 16)).astype(np.float32)
    hm[0, 0] = 0.0                                  # all ties at 0 -> idx 0, coords zeroed (max <= 0)
    hm[0, 1] = -1.0                                 # negative maximum
    hm[0, 2] = 0.5; hm[0, 2, 3, 7] = 0.5            # all ties positive -> first index
    hm[0, 3, 5, 5] = 9.0; hm[0, 3, 9, 2] = 9.0      # two equal maxima -> first wins
    hm[0, 4, 15, 15] = 99.0                         # last element
    hm[0, 5, 7, 3] = np.nan                         # NaN counts as maximum
    hm[0, 6, 2, 2] = np.nan; hm[0, 6, 1, 9] = np.nan  # first NaN wins
    hm[0, 7, 4, 4] = np.inf
    hm[0, 8] = -np.inf
    hm[0, 9] = -0.0; hm[0, 9, 8, 8] = 0.0           # -0 == +0 tie -> first index (0)
    hm[0, 10, 0, 0] = 50.0
    hm[0, 11, 6, 0] = 50.0                          # x == 0
    hm[0, 12] = 1e-30                               # tiny positive ties
    hm[0, 13, 3, 3] = np.inf; hm[0, 13, 9, 9] = np.nan  # NaN beats inf
    preds, maxvals = ns.get_max_preds(hm)
    # a 64x64 case through the whole loop decode: get_max_preds then * 4 (src/utils/method.py:173-178)
    hm64 = rng.standard_normal((2, 21, 64, 64)).astype(np.float32)
    p64, m64 = ns.get_max_preds(hm64)
    idx64 = np.argmax(hm64.reshape(2, 21, -1), 2)
    save("decode.npz", hm=hm, preds=preds, maxvals=maxvals,  
         hm64=hm64, preds64=p64, maxvals64=m64, idx64=idx64)


def make_metrics(ns):
    rng = np.random.default_rng(14)
    n = 24
    gt = rng.uniform(20, 236, (n, 21, 2)).astype(np.float32)
    pred = (gt + rng.standard_normal((n, 21, 2)).astype(np.float32) * rng.choice([1.0, 8.0, 40.0], (n, 21, 1)).astype(np.float32))
    pred = np.floor(pred / 4).astype(np.float32) * 4      # decoded predictions are multiples of 4
    vis = (rng.uniform(size=(n, 21, 1)) > 0.25).astype(np.float32)
    gt3 = np.concatenate([gt, vis], -1)
    # a degenerate sample: all joints identical -> zero bbox diagonal -> inf / nan ratios (Appendix B.6)
    gt_deg = gt.copy(); gt_deg[3] = gt_deg[3, 0]
    pred_deg = pred.copy(); pred_deg[3, :5] = gt_deg[3, :5]
    out = {}
    Ts = [0.05, 0.1, 0.2, 0.3, 0.5]
    tp, tg, tg3 = torch.tensor(pred), torch.tensor(gt), torch.tensor(gt3)
    out["pck_prop"] = np.array([ns.PCK_2d_loss(tp, tg, T, "proportion") for T in Ts])
    out["pck_mm"] = np.array([ns.PCK_2d_loss(tp, tg, T, "mm") for T in [1.0, 3.0, 5.0, 10.0, 20.0]])
    out["pck_prop_gt3"] = np.array([ns.PCK_2d_loss(tp, tg3[:, :, :2], T, "proportion") for T in Ts])
    out["pck_deg"] = np.array([ns.PCK_2d_loss(torch.tensor(pred_deg), torch.tensor(gt_deg), T, "proportion") for T in Ts])
    out["pck_vis_prop"] = np.array([ns.PCK_2d_loss_visible(tp, tg3, T, "proportion") for T in Ts])
    out["pck_vis_mm"] = np.array([ns.PCK_2d_loss_visible(tp, tg3, T, "mm") for T in [1.0, 3.0, 5.0, 10.0, 20.0]])
    pl, al = ns.PCK_2d_loss_list(tp, tg, T_list=[0.1, 0.2], threshold="proportion", auc_t=[0.05, 0.3])
    out["pck_list"] = np.array(pl)
    out["pck_list_auc"] = np.array(al)
    (s, c), _ = ns.EPE_train(tp, tg)
    out["epe_train"] = np.array([s, c], dtype=np.float64)
    (s, c), _ = ns.EPE(tp, tg3)
    out["epe_vis"] = np.array([s, c], dtype=np.float64)
    # AverageMeter semantics (src/utils/metric_logger.py:7-28)
    m = ns.AverageMeter(); m.update(0.5, 4); m.update(0.25, 12)
    mp = ns.AverageMeter(); mp.update_p(10.0, 4); mp.update_p(30.0, 16)
    out["meter"] = np.array([m.sum, m.count, m.avg, mp.sum, mp.count, mp.avg])
    save("metrics.npz", pred=pred, gt=gt, gt3=gt3, pred_deg=pred_deg, gt_deg=gt_deg, Ts=np.array(Ts), **out)


def make_eval(ns):
    """pred_store's JSON schema -> pred_eval / pred_test (src/utils/argparser.py:246-438)."""
    rng = np.random.default_rng(15)
    cats = ["Standard", "Occlusion_by_Pinky", "Occlusion_by_Thumb", "Occlusion_by_Both"]
    sizes = [9, 5, 7, 3]
    meta = {}
    for c, n in zip(cats, sizes):
        gt = rng.uniform(20, 236, (n, 21, 2)).astype(np.float32)
        vis = (rng.uniform(size=(n, 21, 1)) > 0.2).astype(np.float32)
        pred = np.floor((gt + 6 * rng.standard_normal((n, 21, 2))).astype(np.float32) / 4).astype(np.float32) * 4
        bb = []
        for j in torch.tensor(gt):
            w = max(j[:, 0]) - min(j[:, 0]); h = max(j[:, 1]) - min(j[:, 1])
            bb.append(np.sqrt(w ** 2 + h ** 2).item())
        meta[c] = {"bb": bb, "pred": pred.tolist(), "gt": np.concatenate([gt, vis], -1).tolist()}
    res = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            os.makedirs("output/golden", exist_ok=True)
            ns.dump("output/golden/evaluation.json", meta)
            args = types.SimpleNamespace(name="golden")
            bar = types.SimpleNamespace(update=lambda *_a: None)
            for tag, T_list, method in (("mm30", [0, 30], "mm"), ("mm50", [0, 50], "mm"), ("pckb", [0.1, 0.3], "pckb")):
                out, _ = ns.pred_eval(args, T_list, bar, method)
                for k, v in out.items():
                    res[f"{tag}/{k}/auc"] = np.float64(v[0])
                    res[f"{tag}/{k}/epe"] = np.float64(v[1])
                    res[f"{tag}/{k}/pck_t"] = np.asarray(v[2], dtype=np.float64)
            # pred_test: meta = {"pred": [B][n,21,2], "gt": [B][n,21,2], "bb": [B][n]} with equal-size
            # batches (ragged ones break np.array() in the reference under numpy >= 1.24)
            tmeta = {"pred": [], "gt": [], "bb": []}
            for b in range(3):
                gt = rng.uniform(20, 236, (6, 21, 2)).astype(np.float32)
                pred = np.floor((gt + 6 * rng.standard_normal((6, 21, 2))).astype(np.float32) / 4).astype(np.float32) * 4
                bb = []
                for j in torch.tensor(gt):
                    w = max(j[:, 0]) - min(j[:, 0]); h = max(j[:, 1]) - min(j[:, 1])
                    bb.append(np.sqrt(w ** 2 + h ** 2).item())
                tmeta["pred"].append(pred.tolist()); tmeta["gt"].append(gt.tolist()); tmeta["bb"].append(bb)
            os.makedirs("final_model/golden", exist_ok=True)
            ns.dump("final_model/golden/test.json", tmeta)
            for tag, T_list, method in (("mm30", [0, 30], "mm"), ("pckb", [0.1, 0.3], "pckb")):
                try:
                    auc, mean_px, _ = ns.pred_test(args, T_list, bar, method)
                    res[f"test/{tag}/auc"] = np.float64(auc)
                    res[f"test/{tag}/mean_px"] = np.float64(mean_px)
                except Exception as e:
                    print("pred_test", tag, "failed in the reference:", repr(e)[:200])
                    raise
        finally:
            os.chdir(cwd)
    with open(os.path.join(HERE, "eval_meta.json"), "w") as f:
        json.dump({"eval": meta, "test": tmeta}, f)
    save("eval.npz", **res)


def make_warp(ns):
    """generate_patch_image (src/utils/preprocessing.py:144-162) on uint8 and float32 sources."""
    rng = np.random.default_rng(16)
    src_hw, out_hw = (72, 88), (64, 64)
    cases = [  # bbox (x,y,w,h), flip, scale, rot
        ([8.0, 4.0, 64.0, 64.0], False, 1.0, 0.0),
        ([8.0, 4.0, 64.0, 64.0], True, 1.0, 0.0),
        ([10.5, 6.25, 60.0, 60.0], False, 1.2, 30.0),
        ([10.5, 6.25, 60.0, 60.0], True, 0.8, -75.0),
        ([-10.0, -12.0, 90.0, 90.0], False, 1.25, 90.0),   # crop leaves the image: constant border
        ([30.0, 20.0, 20.0, 20.0], True, 0.75, 45.0),      # strong magnification
        ([0.0, 0.0, 88.0, 72.0], False, 1.0, 180.0),       # non-square bbox
    ]
    imgs = rng.integers(0, 256, (len(cases), *src_hw, 3), dtype=np.uint8)
    imgs[1, ::2] = 255; imgs[1, 1::2] = 0           # stripes: worst case for rounding
    p_u8, p_f32, tr, itr = [], [], [], []
    for img, (bbox, flip, scale, rot) in zip(imgs, cases):
        a, t, it = ns.generate_patch_image(img, bbox, flip, scale, rot, out_hw)
        b, t2, _ = ns.generate_patch_image(img.astype(np.float32), bbox, flip, scale, rot, out_hw)
        assert np.array_equal(t, t2)
        p_u8.append(a); p_f32.append(b); tr.append(t); itr.append(it)
    pts = rng.uniform(0, 64, (5, 2))
    moved = np.stack([[ns.trans_point2d(p, t) for p in pts] for t in tr])
    save("warp.npz", imgs=imgs, bbox=np.array([c[0] for c in cases]), flip=np.array([c[1] for c in cases]),
         scale=np.array([c[2] for c in cases]), rot=np.array([c[3] for c in cases]), out_hw=np.array(out_hw),
         patch_u8src=np.stack(p_u8), patch_f32src=np.stack(p_f32), trans=np.stack(tr), inv_trans=np.stack(itr),
         pts=pts, pts_moved=moved)


def make_resize():
    """ToTensor -> Resize -> Normalize exactly as src/tools/dataset.py:147-157 composes them
    (torchvision; antialias only matters when down-sampling, the reference up-samples 224 -> 256)."""
    from torchvision import transforms
    rng = np.random.default_rng(17)
    imgs = rng.integers(0, 256, (3, 56, 56, 3), dtype=np.uint8)
    trans = transforms.Compose([transforms.ToTensor(), transforms.Resize((64, 64)),
                                transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    out = np.stack([trans(i).numpy() for i in imgs])
    trans2 = transforms.Compose([transforms.ToTensor(), transforms.Resize((80, 72)),
                                 transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    out2 = np.stack([trans2(i).numpy() for i in imgs])
    save("resize.npz", imgs=imgs, out=out, out_80x72=out2)


def make_frei(ns):
    """HandMeshTSVDataset.rgb_processing / j2d_processing (src/datasets/frei_dataloader.py:134-160) and the tail
    of __getitem__ (:276-283, 334-338: Resize, Normalize, channels (2,1,0), joint rescale), called unbound on a
    stand-in ``self``.  Small geometry (img_res 56 -> 64) keeps the fixture small; crop() itself is also stored at
    the real 224 geometry for one rotated sample."""
    from torchvision import transforms
    rng = np.random.default_rng(18)
    ds = ns.HandMeshTSVDataset
    stub = types.SimpleNamespace(img_res=56)
    norm = transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
    n = 8
    imgs = rng.integers(0, 256, (n, 64, 64, 3), dtype=np.uint8)
    centers = 32 + rng.uniform(-8, 8, (n, 2))
    scales = rng.uniform(0.2, 0.4, n)
    rots = np.where(np.arange(n) % 2 == 0, 0.0, np.clip(rng.standard_normal(n) * 90, -180, 180))
    rots[1] = 90.0
    pns = rng.uniform(0.6, 1.4, (n, 3))
    kps = np.concatenate([rng.uniform(0, 64, (n, 21, 2)), np.ones((n, 21, 1))], axis=2)
    rgb, final, j2d, joint_2d = [], [], [], []
    for i in range(n):
        a = ds.rgb_processing(stub, imgs[i], centers[i], scales[i], rots[i], 0, pns[i])
        rgb.append(a)
        t = transforms.Resize((64, 64))(torch.from_numpy(a).float())
        final.append(norm(t)[(2, 1, 0), :, :].numpy())
        k = ds.j2d_processing(stub, kps[i].copy(), centers[i], scales[i], rots[i], 0)
        j2d.append(k)
        joint_2d.append(((torch.from_numpy(k).float()[:, :-1] * 100 + 112) * (256 / 224)).numpy())
    big = rng.integers(0, 256, (1, 224, 224, 3), dtype=np.uint8)
    big_par = np.array([[120.25, 100.0, 1.1, -37.5]])      # cx, cy, scale, rot
    crops = [ns.crop(big[i], big_par[i, :2], big_par[i, 2], [224, 224], rot=big_par[i, 3]).astype(np.float32)
             for i in range(1)]
    save("frei.npz", imgs=imgs, centers=centers, scales=scales, rots=rots, pns=pns, kps=kps, rgb=np.stack(rgb),
         final=np.stack(final), j2d=np.stack(j2d), joint_2d=np.stack(joint_2d), big=big, big_par=big_par,
         crops=np.stack(crops))


def make_augment(ns):
    """augmentation() (src/utils/preprocessing.py:65-92) with get_aug_config pinned to fixed draws: the transformed
    key points and joint_valid for two-hand (42 joints, right / left swap) and single-hand (21 joints) skeletons."""
    rng = np.random.default_rng(77)
    pre = ns.preprocessing
    keep = pre.get_aug_config
    cases = []
    try:
        for n_j in (42, 21):
            jt = {"right": np.arange(0, 21), "left": np.arange(21, 42)} if n_j == 42 else \
                 {"right": np.arange(0, 0), "left": np.arange(0, 0)}
            for _ in range(12):
                img = rng.integers(0, 256, (240, 320, 3), dtype=np.uint8)
                jc = rng.uniform(-20, 340, (n_j, 3)).astype(np.float32)
                jv = (rng.uniform(size=n_j) < 0.8).astype(np.float32)
                bbox = np.array([rng.uniform(0, 80), rng.uniform(0, 60), rng.uniform(120, 220), rng.uniform(120, 180)], np.float32)
                cfgd = ([float(rng.uniform(-0.15, 0.15)), float(rng.uniform(-0.15, 0.15))], float(rng.uniform(0.75, 1.25)),
                        float(rng.uniform(-90, 90)) if rng.uniform() < 0.6 else 0.0, bool(rng.uniform() < 0.5),
                        rng.uniform(0.8, 1.2, 3))
                pre.get_aug_config = lambda c=cfgd: c
                _img, jc_o, jv_o, _ht, inv = pre.augmentation(img, bbox.copy(), jc, jv.copy(), np.array([1.0, 0.0], np.float32),
                                                             "train", jt)
                b2 = bbox.copy()
                b2[0] += b2[2] * cfgd[0][0]; b2[1] += b2[3] * cfgd[0][1]
                _p, trans, _inv = ns.generate_patch_image(img, b2, cfgd[3], cfgd[1], cfgd[2], (256, 256))
                cases.append((n_j, jc, jv, trans, cfgd[3], jc_o, jv_o))
    finally:
        pre.get_aug_config = keep
    for n_j in (42, 21):
        sel = [c for c in cases if c[0] == n_j]
        save(f"augment{n_j}.npz", joints=np.stack([c[1] for c in sel]), valid=np.stack([c[2] for c in sel]),
             trans=np.stack([c[3] for c in sel]), flip=np.array([c[4] for c in sel]), orig_w=np.int64(320),
             joints_out=np.stack([c[5] for c in sel]), valid_out=np.stack([c[6] for c in sel]))


def main():
    if not rb.available():
        raise SystemExit("reference tree not mounted; fixtures can only be generated in the build container")
    ns = rb.load()
    make_targets(ns)
    make_loss(ns)
    make_decode(ns)
    make_metrics(ns)
    make_eval(ns)
    make_warp(ns)
    make_resize()
    make_frei(ns)
    make_augment(ns)
    import cv2
    import torchvision
    with open(os.path.join(HERE, "VERSIONS.json"), "w") as f:
        json.dump({"numpy": np.__version__, "torch": torch.__version__, "cv2": cv2.__version__,
                   "torchvision": torchvision.__version__, "python": sys.version.split()[0]}, f, indent=1)


if __name__ == "__main__":
    main()
