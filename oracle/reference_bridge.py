"""TEST INFRASTRUCTURE ONLY -- imports the *real* reference from /root/reference.

Works only in the build container (the GPU box has no /root/reference).  Used by
``tests/golden/make_golden.py`` to generate the committed fixtures and by the
``not gpu`` test ``tests/test_oracle_vs_reference.py`` (skipped when the
reference tree is absent) to pin ``oracle/restate.py`` live.

The reference imports four third-party modules that are not installed here and
that the 2-D heatmap path never executes (matplotlib, pycocotools, easydict,
yacs); empty stand-ins are registered in ``sys.modules`` before import
(SURVEY.md Appendix C).
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LIGHTHAND_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "utils"))


class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class _CfgNode(dict):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def merge_from_file(self, *_a, **_k):
        pass

    def merge_from_list(self, *_a, **_k):
        pass

    def defrost(self):
        pass

    def freeze(self):
        pass

    def clone(self):
        return _CfgNode(self)


_loaded = None


def load():
    """Import the reference hot-path symbols; returns a namespace."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in ("matplotlib", "matplotlib.pyplot", "pycocotools", "pycocotools.coco",
                 "easydict", "yacs", "yacs.config"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["pycocotools"].coco = sys.modules["pycocotools.coco"]
    sys.modules["pycocotools.coco"].COCO = object
    sys.modules["easydict"].EasyDict = _AttrDict
    sys.modules["yacs"].config = sys.modules["yacs.config"]
    sys.modules["yacs.config"].CfgNode = _CfgNode
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    ns = types.SimpleNamespace()
    from src.utils import loss as _loss
    ns.JointsMSELoss = _loss.JointsMSELoss
    ns.get_max_preds = _loss.get_max_preds
    ns.PCK_2d_loss = _loss.PCK_2d_loss
    ns.PCK_2d_loss_visible = _loss.PCK_2d_loss_visible
    ns.PCK_2d_loss_list = _loss.PCK_2d_loss_list
    ns.EPE = _loss.EPE
    ns.EPE_train = _loss.EPE_train
    from src.utils import preprocessing as _pre
    ns.generate_patch_image = _pre.generate_patch_image
    ns.gen_trans_from_patch_cv = _pre.gen_trans_from_patch_cv
    ns.trans_point2d = _pre.trans_point2d
    ns.augmentation = _pre.augmentation
    ns.preprocessing = _pre          # module handle: tests pin get_aug_config to fixed draws
    from src.tools import dataset as _ds
    ns.CustomDataset = _ds.CustomDataset
    ns.generate_target = lambda joints: _ds.CustomDataset.generate_target(None, joints)
    from src.utils import dataset_loader as _dl
    ns.GenerateHeatmap = _dl.GenerateHeatmap
    from src.utils import metric_logger as _ml
    ns.AverageMeter = _ml.AverageMeter
    from src.utils import argparser as _ap
    ns.pred_eval = _ap.pred_eval
    ns.pred_test = _ap.pred_test
    from src.utils import dir as _dir
    ns.dump = _dir.dump
    from src.utils import image_ops as _io
    ns.crop = _io.crop
    ns.transform = _io.transform
    ns.get_transform = _io.get_transform
    from src.datasets import frei_dataloader as _frei
    ns.HandMeshTSVDataset = _frei.HandMeshTSVDataset
    _loaded = ns
    return ns
