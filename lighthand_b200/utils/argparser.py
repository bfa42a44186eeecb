"""Offline evaluation sweep with the reference's call surface (src/utils/argparser.py:246-438):
PCK curve / AUC / EPE.  The counting core (float64 distances, strict '<' against 100 thresholds,
per pose category) runs in K5 on the device; ``CurveAccumulator`` keeps the histograms resident so a
sharded evaluation needs one all-reduce (lighthand_b200.distributed) and one read at the end.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

from .. import ops
from .loss import _to_cuda

MM_PRED_EVAL = 2.83464567      # src/utils/argparser.py:336
MM_PER_PX = 3.7795275591       # src/utils/argparser.py:377,399
POSE_CATEGORIES = ("Standard", "Occlusion_by_Pinky", "Occlusion_by_Thumb", "Occlusion_by_Both")  # :247-252


def _trapz(y, x):
    f = getattr(np, "trapz", None) or np.trapezoid
    return f(y, x)


def eval_thresholds(T_list, method, mm_factor=MM_PRED_EVAL):
    """src/utils/argparser.py:334-342 / :397-403."""
    if method == "mm":
        return np.array(np.linspace(T_list[0], T_list[-1], 101)[1:] * mm_factor)
    if method == "pckb":
        return np.array(np.linspace(T_list[0], T_list[-1], 100))
    assert 0, "this method is the wrong"


class CurveAccumulator:
    """Device-resident state of one (T_list, method) sweep over any number of batches / categories.

    hist int64[n_cat, K+1] (first-satisfied-threshold histogram), n_valid / n_all int64[n_cat],
    sum_d float64[n_cat].  All additive => shard the samples over ranks, all-reduce once.
    """

    def __init__(self, T_list, method, n_cat=len(POSE_CATEGORIES), use_vis=True, mm_factor=MM_PRED_EVAL,
                 device="cuda"):
        ops.require_cuda()
        self.method, self.use_vis, self.n_cat = method, use_vis, n_cat
        self.thresholds = eval_thresholds(T_list, method, mm_factor)
        self.order = np.argsort(self.thresholds, kind="stable")
        self.device = torch.device(device)
        self.thr_sorted = torch.from_numpy(np.ascontiguousarray(self.thresholds[self.order])).to(self.device)
        k = len(self.thresholds)
        # one packed int64 buffer [n_cat, K+1 | n_valid | n_all] + one float64 [n_cat]: two all-reduces
        self.ints = torch.zeros((n_cat, k + 3), dtype=torch.int64, device=self.device)
        self.sum_d = torch.zeros(n_cat, dtype=torch.float64, device=self.device)
        self.hist = torch.zeros((n_cat, k + 1), dtype=torch.int64, device=self.device)
        self.n_valid = torch.zeros(n_cat, dtype=torch.int64, device=self.device)
        self.n_all = torch.zeros(n_cat, dtype=torch.int64, device=self.device)

    def update(self, pred_xy, gt, bbox=None, category=None):
        """pred_xy [n,J,2], gt [n,J,2|3] (last column = visibility when use_vis), bbox [n] or None
        (computed like pred_store, src/utils/argparser.py:268-272), category int32 [n] or None."""
        ops.pck_curve_accum(pred_xy, gt, self.thr_sorted, self.method, self.use_vis, self.hist,
                            self.n_valid, self.n_all, self.sum_d, bbox, category)

    def tensors(self):
        """The additive state, for all-reduce."""
        return [self.hist, self.n_valid, self.n_all, self.sum_d]

    def finalize(self, names=POSE_CATEGORIES, seed_rows=971):
        """-> {name: [auc, epe_mm, pck_t[K]], ..., "mean_auc": [...]} exactly as pred_eval returns it
        (src/utils/argparser.py:370-386), including the 971x21 zero rows pooled into the mean EPE (:345)."""
        hist = self.hist.cpu().numpy()
        n_valid = self.n_valid.cpu().numpy()
        n_all = self.n_all.cpu().numpy()
        sum_d = self.sum_d.cpu().numpy()
        thr = self.thresholds
        k = len(thr)
        inv = np.empty(k, dtype=np.int64)
        inv[self.order] = np.arange(k)
        norm_factor = _trapz(np.ones_like(thr), thr)
        out = {}

        def curve(h, total):
            cnt_sorted = np.cumsum(h[:k])
            cnt = cnt_sorted[inv]
            pck_t = np.array([(int(c) / int(total)) * 100 for c in cnt])
            auc = _trapz(pck_t, thr) / (norm_factor + sys.float_info.epsilon)
            return auc, pck_t

        for c in range(self.n_cat):
            if n_all[c] == 0:
                continue
            auc, pck_t = curve(hist[c], n_valid[c])
            out[str(names[c]) if c < len(names) else str(c)] = [auc, (sum_d[c] / n_all[c]) / MM_PER_PX, pck_t]
        auc, pck_t = curve(hist.sum(0), n_valid.sum())
        j = 21
        out["mean_auc"] = [auc, (sum_d.sum() / (n_all.sum() + seed_rows * j)) / MM_PER_PX, pck_t]
        return out


def pred_store(args, dataloader, model, pbar):
    """src/utils/argparser.py:246-281: run the model over the evaluation loader, decode (x4), bucket by pose
    category and write ``output/<args.name>/evaluation.json`` in the reference's schema
    ``[{category: {"bb": [n], "pred": [n][21][2], "gt": [n][21][3]}}]``.

    The decode and the bounding-box diagonals are computed on the device (``lh_decode``, ``lh_bbox_diag``: float32
    builtin max / min / sqrt exactly like the reference's per-sample loop); results stay on the device for the whole
    sweep and ONE device->host copy at the end feeds the JSON lists.
    Returns the meta dict as well."""
    from .dir import dump
    meta = {name: {"bb": [], "pred": [], "gt": []} for name in POSE_CATEGORIES}
    dev = next(model.parameters()).device
    coords_all, bbox_all, gt_all, anno_all = [], [], [], []
    with torch.no_grad():
        for images, gt_2d_joints, annos in dataloader:
            anno, _idx = annos
            pred = model(images.to(dev, non_blocking=True))
            coords, _mv, _ai = ops.decode(pred.float().contiguous(), 4.0)
            gt = gt_2d_joints.to(dev, torch.float32)
            coords_all.append(coords)
            bbox_all.append(ops.bbox_diag(gt))                 # lh_bbox_diag: the per-sample loop of :268-272, float32
            gt_all.append(gt_2d_joints)
            anno_all.extend(anno)
            pbar.update(1)
    if coords_all:
        # ONE device->host copy for the whole sweep, then the reference's bucketing into JSON lists
        pred_l, bbox_l = torch.cat(coords_all).cpu().tolist(), torch.cat(bbox_all).cpu().tolist()
        gt_l = torch.cat([g.cpu() if isinstance(g, torch.Tensor) else torch.as_tensor(g) for g in gt_all]).tolist()
        for i, name in enumerate(anno_all):
            meta[name]["bb"].append(bbox_l[i])
            meta[name]["pred"].append(pred_l[i])
            meta[name]["gt"].append(gt_l[i])
    os.makedirs(os.path.join("output", args.name), exist_ok=True)
    dump(os.path.join("output", args.name, "evaluation.json"), meta)
    return meta


def write_pck_eval(file_name, pck_list):
    """The ``pck_eval_*.txt`` writer of src/tools/wearable_eval_2d.py:64-79: one line per (category, model):
    ``category;model;auc;epe;pck_0;...;pck_99;``.  ``pck_list`` = [[pred_eval result, model name], ...]."""
    with open(file_name, "w") as f:
        for total_pck, name in pck_list:
            for p_type in total_pck:
                f.write("{};{};{:.2f};{:.2f};".format(p_type, name, total_pck[p_type][0], total_pck[p_type][1]))
                n = len(total_pck[p_type][2])
                for idx, pck in enumerate(total_pck[p_type][2]):
                    f.write("{:.2f};".format(pck))
                    if idx == n - 1:
                        f.write("\n")


def pred_eval_meta(meta, T_list, method):
    """pred_eval on an in-memory meta dict {category: {"bb", "pred", "gt"}} (what pred_store dumps)."""
    names = list(meta.keys())
    acc = CurveAccumulator(T_list, method, n_cat=len(names), use_vis=True)
    for ci, name in enumerate(names):
        gt = _to_cuda(np.asarray(meta[name]["gt"], dtype=np.float64), torch.float32)
        pred = _to_cuda(np.asarray(meta[name]["pred"], dtype=np.float64), torch.float32)[:, :, :2].contiguous()
        bb = _to_cuda(np.asarray(meta[name]["bb"], dtype=np.float64), torch.float32)
        cat = torch.full((gt.shape[0],), ci, dtype=torch.int32, device=gt.device)
        acc.update(pred, gt, bb, cat)
    return acc.finalize(names)


def pred_eval(args, T_list, p_bar, method):
    """src/utils/argparser.py:326-388: reads output/<name>/evaluation.json, returns (pck_list, p_bar)."""
    with open(os.path.join("output", args.name, "evaluation.json"), "r") as fi:
        meta = json.load(fi)
    meta = meta[0]
    out = pred_eval_meta(meta, T_list, method)
    for _ in meta:
        p_bar.update(1)
    return out, p_bar


def pred_test_meta(meta, T_list, method):
    """pred_test on an in-memory meta {"pred": [B][n,21,2], "gt": ..., "bb": [B][n]} -> (auc, mean_px)."""
    acc = CurveAccumulator(T_list, method, n_cat=1, use_vis=False, mm_factor=MM_PER_PX)
    for pred, gt, bb in zip(meta["pred"], meta["gt"], meta["bb"]):
        gt = _to_cuda(np.asarray(gt, dtype=np.float64), torch.float32)
        pred = _to_cuda(np.asarray(pred, dtype=np.float64), torch.float32)
        acc.update(pred, gt, _to_cuda(np.asarray(bb, dtype=np.float64), torch.float32), None)
    res = acc.finalize(("all",), seed_rows=0)["all"]
    return res[0], res[1] * MM_PER_PX


def pred_test(args, T_list, pbar, method):
    """src/utils/argparser.py:391-438: reads final_model/<name>/test.json -> (auc, diff.mean(), pbar)."""
    with open(os.path.join("final_model", args.name, "test.json"), "r") as fi:
        meta = json.load(fi)
    auc, mean_px = pred_test_meta(meta[0], T_list, method)
    pbar.update(1)
    return auc, mean_px, pbar
