"""TEST / BENCH INFRASTRUCTURE ONLY -- one pass of the hot path with the reference's OWN functions.

``bench.py --impl reference`` (and its ``cpu_baseline`` leg) time the UNMODIFIED reference code where it can be had:
``__graft_entry__.build()`` mirrors the reference's pure-Python ``src/`` tree into ``baseline/_ref/`` (git-ignored,
travels to the GPU box; the reference ships no setup.py / pyproject, so ``pip install --target baseline/_ref`` has
nothing to build -- DESIGN.md section 8).  The functions executed per batch, exactly as the reference's loop calls them:

    generate_patch_image                 src/utils/preprocessing.py:144-162   (cv2.warpAffine crop / flip / rotate)
    ToTensor + Normalize                 src/tools/dataset.py:147-157         (torchvision transforms, as the dataset composes them)
    CustomDataset.generate_target        src/tools/dataset.py:165-212         (64x64 / sigma 2 geometry only: it is hard-coded)
    JointsMSELoss (+ loss.backward())    src/utils/loss.py:306-325, src/utils/method.py:169,181-183
    get_max_preds, * 4                   src/utils/loss.py:327-355, src/utils/method.py:173-178
    PCK_2d_loss, EPE_train               src/utils/loss.py:116-148, 50-67, src/utils/method.py:243-250

For geometries the reference hard-codes away (56x56, 128x128 sigma 3) ``generate_target`` comes from the oracle's
parametrised restatement; everything else is still the reference's function.
"""
from __future__ import annotations

import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VENDORED = os.path.join(ROOT, "baseline", "_ref")


def locate():
    """Reference tree to import from: the live mount in the build container, else the mirrored copy."""
    for root in (os.environ.get("LIGHTHAND_REFERENCE"), "/root/reference", VENDORED):
        if root and os.path.isdir(os.path.join(root, "src", "utils")):
            return root
    return None


def load():
    """(namespace of reference functions, kind) or (None, 'port') when no reference tree is reachable."""
    root = locate()
    if root is None:
        return None, "port"
    os.environ["LIGHTHAND_REFERENCE"] = root
    from oracle import reference_bridge as B
    B.REFERENCE_ROOT = root
    return B.load(), "reference"


def make_pass(ns, R, torch):
    """One pass of the reference's per-iteration work over one host batch; ``ns`` = reference namespace (or None:
    the oracle port ``R`` for every stage)."""
    import numpy as np
    if ns is None:
        def cpu_pass(data, S, H, sigma, train):
            imgs, joints, bbox, scale, rot, flip, pred = data
            stride = S / H
            tgt = []
            for i in range(imgs.shape[0]):
                R.warp_normalize_sample(imgs[i], bbox[i], bool(flip[i]), float(scale[i]), float(rot[i]), (S, S))
                tgt.append(R.generate_target(joints[i], 21, (H, H), (stride, stride), sigma)[0])
            tgt = np.stack(tgt)
            if train:
                R.joints_mse_loss_and_grad(pred, tgt)
            else:
                float(R.joints_mse_loss(pred, tgt))
            coords, _ = R.get_max_preds(pred)
            coords = coords * stride
            R.pck_2d(coords, joints, 0.2, "proportion")
            R.epe_train(coords, joints)
        return cpu_pass

    from torchvision import transforms
    to_input = transforms.Compose([transforms.ToTensor(),
                                   transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    criterion = ns.JointsMSELoss(use_target_weight=False)

    def cpu_pass(data, S, H, sigma, train):
        imgs, joints, bbox, scale, rot, flip, pred = data
        stride = S / H
        images, heatmaps = [], []
        for i in range(imgs.shape[0]):                      # what the DataLoader workers do per sample
            patch, _trans, _inv = ns.generate_patch_image(imgs[i], bbox[i], bool(flip[i]), float(scale[i]), float(rot[i]), (S, S))
            images.append(to_input(patch.astype(np.uint8)))
            jt = torch.from_numpy(joints[i])
            if H == 64 and sigma == 2:
                heatmaps.append(ns.generate_target(jt))
            else:
                heatmaps.append(torch.from_numpy(R.generate_target(joints[i], 21, (H, H), (stride, stride), sigma)[0]))
        _images = torch.stack(images)                       # default collate
        gt_heatmaps = torch.stack(heatmaps)
        p = torch.from_numpy(pred).requires_grad_(train)    # the backbone's output
        if train:
            loss = criterion(p, gt_heatmaps, None)
            loss.item()
            loss.backward()
        else:
            with torch.no_grad():
                criterion(p, gt_heatmaps, None).item()
        pred_joint, _ = ns.get_max_preds(np.array(p.detach().cpu()))
        pred_joint = torch.tensor(pred_joint * stride)
        gt = torch.from_numpy(joints)
        ns.PCK_2d_loss(pred_joint, gt, T=0.2, threshold="proportion")
        ns.EPE_train(pred_joint, gt)
    return cpu_pass
