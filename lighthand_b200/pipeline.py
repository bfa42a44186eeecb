"""The whole hot path as one object.

``HeatmapPipeline``  device-resident step  K1 warp+normalise -> fused K2+K3+K4+K5 (targets, JointsMSELoss
                     forward+backward, decode, PCK/EPE accumulation).  Two kernel launches, no
                     allocation, no host synchronisation; optionally replayed from a CUDA graph.  This is
                     what replaces, per iteration, ``__getitem__``/collate + ``JointsMSELoss`` +
                     ``.cpu()`` + ``get_max_preds`` + ``PCK_2d_loss`` + ``EPE_train`` of the reference's
                     loops (src/utils/method.py:156-287).
``HostPipeline``     the same step through the C-ABI context (lh_pipeline_*): HOST buffers in, host
                     results out, double-buffered so the H2D copy of step i+1 overlaps step i.

There is no CPU implementation behind either.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib, ops
from ._lib import LighthandError, check
from .utils.metric_logger import DeviceMeter


class HeatmapPipeline:
    def __init__(self, batch: int, src_hw=(256, 256), out_hw=(256, 256), heatmap_hw=(64, 64), num_joints: int = 21,
                 sigma: float = 2, interp: int = ops.LH_INTERP_CV2_U8, train: bool = True,
                 materialize_target: bool = False, subpixel: bool = False, map_back: bool = False,
                 pck_T: float = 0.2, mean=ops.IMAGENET_MEAN, std=ops.IMAGENET_STD, device="cuda",
                 use_graph: bool = False, exact_norm: bool = False):
        ops.require_cuda()
        self.exact_norm = exact_norm
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.N, self.J = int(batch), int(num_joints)
        self.src_hw, self.out_hw, self.hm_hw = tuple(src_hw), tuple(out_hw), tuple(heatmap_hw)
        self.sigma, self.interp, self.train = sigma, interp, train
        self.materialize_target, self.subpixel, self.map_back, self.pck_T = materialize_target, subpixel, map_back, pck_T
        self.mean, self.std = mean, std
        self.stride = (self.out_hw[1] / self.hm_hw[1], self.out_hw[0] / self.hm_hw[0])
        d, f32 = self.device, torch.float32
        n, j, (h, w) = self.N, self.J, self.hm_hw
        self.input = torch.empty((n, 3, *self.out_hw), dtype=f32, device=d)
        self.buffers = {
            "loss": torch.zeros((), dtype=f32, device=d),
            "coords": torch.empty((n, j, 2), dtype=f32, device=d),
            "maxvals": torch.empty((n, j, 1), dtype=f32, device=d),
            "argidx": torch.empty((n, j), dtype=torch.int32, device=d),
            "weight": torch.empty((n, j, 1), dtype=f32, device=d),
            "ws": torch.zeros((_lib.load().lh_heatmap_ws_bytes(n, j) + 7) // 8, dtype=torch.int64, device=d),
        }
        if train:
            self.buffers["grad"] = torch.empty((n, j, h, w), dtype=f32, device=d)
        if materialize_target:
            self.buffers["target"] = torch.empty((n, j, h, w), dtype=f32, device=d)
        self.lut = ops.build_norm_lut(mean, std, None, d) if (interp == ops.LH_INTERP_CV2_U8 and exact_norm) else None
        self.meter = DeviceMeter(d)
        self.use_graph = use_graph
        self._graphs = {}
        self.launches_per_step = 2

    # -- one step, asynchronous on the current stream -------------------------------------------
    def _enqueue(self, images_u8, fwd_affine, flip, joints, pred, inv_affine):
        ops.warp_normalize(images_u8, fwd_affine, self.out_hw, flip=flip, mean=self.mean, std=self.std,
                           lut=self.lut, interp=self.interp, out=self.input, exact_norm=self.exact_norm)
        ops.heatmap_fused(pred, joints, sigma=self.sigma, feat_stride=self.stride,
                          want_target=self.materialize_target, want_grad=self.train, want_decode=True,
                          out_scale=self.stride[0], subpixel=self.subpixel,
                          inv_affine=inv_affine if self.map_back else None,
                          loss_meter=self.meter.loss_meter, buffers=self.buffers,
                          metrics=(self.meter.counts, self.meter.sums), pck_T=self.pck_T)

    def step(self, images_u8: torch.Tensor, fwd_affine: Optional[torch.Tensor], joints: torch.Tensor,
             pred: torch.Tensor, flip: Optional[torch.Tensor] = None, inv_affine: Optional[torch.Tensor] = None) -> dict:
        """images_u8 [N,Hs,Ws,3] uint8, fwd_affine [N,2,3] float64 (None for the resize mode), joints
        [N,J,2|3] float32 in network-input pixels, pred [N,J,H,W] float32 -- all on the device.
        Returns views of the pipeline-owned outputs: input, loss, grad, coords, maxvals, argidx[, target]."""
        if images_u8.shape[0] != self.N or pred.shape[0] != self.N:
            raise LighthandError("HeatmapPipeline.step: batch size differs from the one the pipeline was built for")
        if not self.use_graph:
            self._enqueue(images_u8, fwd_affine, flip, joints, pred, inv_affine)
        else:
            key = tuple(0 if t is None else t.data_ptr() for t in (images_u8, fwd_affine, flip, joints, pred, inv_affine))
            g = self._graphs.get(key)
            if g is None:
                # warm-up outside capture (table caches, lazy module load), then capture this pointer set.
                # The metric accumulators are restored so warm-up + capture leave no trace in them.
                snap = (self.meter.counts.clone(), self.meter.sums.clone())
                s = torch.cuda.Stream(self.device)
                s.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(s):
                    self._enqueue(images_u8, fwd_affine, flip, joints, pred, inv_affine)
                torch.cuda.current_stream(self.device).wait_stream(s)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue(images_u8, fwd_affine, flip, joints, pred, inv_affine)
                self.meter.counts.copy_(snap[0])
                self.meter.sums.copy_(snap[1])
                self._graphs[key] = (g, (images_u8, fwd_affine, flip, joints, pred, inv_affine))  # keep inputs alive
                g = self._graphs[key]
            g[0].replay()
        out = dict(self.buffers)
        out["input"] = self.input
        out.pop("ws")
        return out

    def prepare(self, images_u8, fwd_affine, joints, pred, flip=None, inv_affine=None) -> "PreparedStep":
        """Validate and marshal the two C-ABI calls ONCE for a fixed set of device buffers; the returned
        object's ``launch()`` then costs two raw ctypes calls (no per-step Python argument work), which
        keeps the host ahead of a ~150 us device step without needing a CUDA graph."""
        return PreparedStep(self, images_u8, fwd_affine, joints, pred, flip, inv_affine)

    def reset_metrics(self):
        self.meter.reset()

    def metrics(self, all_reduce: bool = True) -> dict:
        """Epoch-end read: optional cross-rank sum (the path's only collective), one D2H."""
        if all_reduce:
            from .distributed import all_reduce_metrics
            all_reduce_metrics([self.meter.counts, self.meter.sums])
        return self.meter.result()

    def algorithmic_bytes_per_image(self) -> dict:
        """Compulsory HBM traffic of each launch per image."""
        sh, sw = self.src_hw
        oh, ow = self.out_hw
        hw = self.hm_hw[0] * self.hm_hw[1]
        k1 = 3 * sh * sw + 12 * oh * ow
        fused = 4 * self.J * hw * (1 + (1 if self.train else 0) + (1 if self.materialize_target else 0)) + 8 * self.J + 16 * self.J
        k5 = 16 * self.J
        return {"warp_normalize": k1, "heatmap_fused": fused, "pck_epe": k5, "step": k1 + fused + k5}


class PreparedStep:
    """Pre-marshalled K1 -> fused(K2..K5) launch sequence bound to fixed device buffers."""

    def __init__(self, pipe: HeatmapPipeline, images_u8, fwd_affine, joints, pred, flip, inv_affine):
        lib = _lib.load()
        p = pipe
        images_u8 = ops._dev(images_u8, torch.uint8, "images")
        joints = ops._dev(joints, torch.float32, "joints")
        pred = ops._dev(pred, torch.float32, "pred")
        n, sh, sw, _ = images_u8.shape
        if n != p.N or tuple(pred.shape) != (p.N, p.J, *p.hm_hw) or (sh, sw) != tuple(p.src_hw):
            raise LighthandError("prepare: tensor shapes differ from the pipeline configuration")
        if fwd_affine is not None:
            fwd_affine = ops._dev(fwd_affine, torch.float64, "fwd_affine")
        elif p.interp != ops.LH_INTERP_TORCH:
            raise LighthandError("prepare: fwd_affine required for the cv2 warp modes")
        if flip is not None:
            flip = ops._dev(flip, torch.uint8, "flip")
        if inv_affine is not None:
            inv_affine = ops._dev(inv_affine, torch.float64, "inv_affine")
        self._keep = (images_u8, fwd_affine, joints, pred, flip, inv_affine, p)
        self._ms = ops._mean_std(p.mean, p.std)
        ptr = ops._ptr
        self._k1 = (images_u8.data_ptr(), sh * sw * 3, sh, sw, sw * 3, ptr(fwd_affine), ptr(flip), None, self._ms,
                    ptr(p.lut), p.interp, ops.LH_CHAN_RGB, ops.LH_LAYOUT_CHW, p.input.data_ptr(), n, p.out_hw[0], p.out_hw[1])
        a = _lib.HeatmapArgs()
        b = p.buffers
        table, side = ops.hrnet_table(p.sigma, p.device)
        a.pred, a.joints, a.gauss_table, a.table_side = pred.data_ptr(), joints.data_ptr(), table.data_ptr(), side
        a.joint_stride = joints.shape[2]
        a.target_out = b["target"].data_ptr() if p.materialize_target else None
        a.target_weight_out = b["weight"].data_ptr()
        a.grad = b["grad"].data_ptr() if p.train else None
        a.loss_f32, a.loss_meter = b["loss"].data_ptr(), p.meter.loss_meter.data_ptr()
        a.coords, a.maxvals, a.argidx = b["coords"].data_ptr(), b["maxvals"].data_ptr(), b["argidx"].data_ptr()
        a.inv_affine = ptr(inv_affine) if p.map_back else None
        a.ws = b["ws"].data_ptr()
        a.N, a.J, a.H, a.W = p.N, p.J, p.hm_hw[0], p.hm_hw[1]
        a.style, a.use_target_weight, a.subpixel = ops.LH_TARGET_HRNET, 0, int(bool(p.subpixel))
        a.stride_x, a.stride_y, a.out_scale = float(p.stride[0]), float(p.stride[1]), float(p.stride[0])
        self._args = a
        a.metric_counts, a.metric_sums = p.meter.counts.data_ptr(), p.meter.sums.data_ptr()
        a.pck_mode, a.pck_thr = ops._pck_mode(p.pck_T, "proportion")
        self._f1, self._f2 = lib.lh_warp_normalize, lib.lh_heatmap_fused
        self._aref = C.byref(a)

    def launch(self, stream: Optional[int] = None, events=None) -> None:
        """Enqueue the step on ``stream`` (raw cudaStream_t; default: torch's current stream).  ``events``
        = (e0, e1[, e2]) torch events recorded before K1, after K1 and (if given) after the fused kernel."""
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        if events is not None:
            events[0].record()
        rc = self._f1(*self._k1, st)
        if events is not None:
            events[1].record()
        rc = rc or self._f2(self._aref, st)
        if events is not None and len(events) > 2:
            events[2].record()
        if rc:
            check(rc, "prepared step")

    def launch_split(self, stream_k1: int, stream_fused: int) -> None:
        """K1 and the fused heat-map pass touch disjoint data (images -> input; pred/joints -> grad/coords), so
        they may run concurrently: K1 of the next batch (data stage) overlaps the loss stage of the current one."""
        rc = self._f1(*self._k1, stream_k1) or self._f2(self._aref, stream_fused)
        if rc:
            check(rc, "prepared step (split streams)")


class HostPipeline:
    """C-ABI context (lh_pipeline_*, include/lighthand_b200.h): host numpy buffers in, host results out.
    ``submit``/``wait`` are double-buffered; ``step`` = submit + wait."""

    def __init__(self, batch: int, src_hw=(256, 256), out_hw=(256, 256), heatmap_hw=(64, 64), num_joints: int = 21,
                 sigma: float = 2, interp: int = ops.LH_INTERP_CV2_U8, train: bool = True,
                 materialize_target: bool = False, subpixel: bool = False, pck_T: float = 0.2,
                 mean=ops.IMAGENET_MEAN, std=ops.IMAGENET_STD, device: int = 0, exact_norm: bool = False):
        ops.require_cuda()
        self.lib = _lib.load()
        cfg = _lib.PipelineCfg()
        cfg.device, cfg.batch = int(device), int(batch)
        cfg.src_h, cfg.src_w = int(src_hw[0]), int(src_hw[1])
        cfg.out_h, cfg.out_w = int(out_hw[0]), int(out_hw[1])
        cfg.J, cfg.H, cfg.W = int(num_joints), int(heatmap_hw[0]), int(heatmap_hw[1])
        cfg.interp, cfg.train, cfg.materialize_target = int(interp), int(train), int(materialize_target)
        cfg.subpixel, cfg.sigma, cfg.pck_T = int(subpixel), float(sigma), float(np.float32(pck_T))
        cfg.exact_norm = int(bool(exact_norm))
        for i, v in enumerate((*mean, *std)):
            cfg.mean_std[i] = float(np.float32(v))
        tmp = int(sigma * 3)
        size = 2 * tmp + 1
        x = np.arange(0, size, 1, np.float32)
        table = np.ascontiguousarray(np.exp(-((x - size // 2) ** 2 + (x[:, None] - size // 2) ** 2) / (2 * sigma ** 2)).astype(np.float32))
        self.cfg, self.N, self.J = cfg, int(batch), int(num_joints)
        self.hm_hw, self.out_hw, self.src_hw = tuple(heatmap_hw), tuple(out_hw), tuple(src_hw)
        self._h = C.c_void_p()
        check(self.lib.lh_pipeline_create(C.byref(cfg), table.ctypes.data, size, C.byref(self._h)), "lh_pipeline_create")
        self._keep = {}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.lh_pipeline_destroy(self._h)
            self._h = None

    __del__ = close

    # device-resident buffers of the context, as torch views (the backbone reads .input, writes .pred)
    def _view(self, ptr, shape):
        if not ptr:
            return None
        n = int(np.prod(shape))
        iface = {"shape": (n,), "typestr": "<f4", "data": (int(ptr), False), "version": 2}
        holder = type("_Dev", (), {"__cuda_array_interface__": iface})()
        return torch.as_tensor(holder, device=f"cuda:{self.cfg.device}").view(*shape)

    @property
    def input(self):
        return self._view(self.lib.lh_pipeline_input_dev(self._h), (self.N, 3, *self.out_hw))

    @property
    def pred(self):
        return self._view(self.lib.lh_pipeline_pred_dev(self._h), (self.N, self.J, *self.hm_hw))

    @property
    def grad(self):
        return self._view(self.lib.lh_pipeline_grad_dev(self._h), (self.N, self.J, *self.hm_hw))

    @property
    def target(self):
        return self._view(self.lib.lh_pipeline_target_dev(self._h), (self.N, self.J, *self.hm_hw))

    @property
    def launches_per_step(self):
        return self.lib.lh_pipeline_launches_per_step(self._h)

    def pinned(self, shape, dtype):
        """numpy array over pinned host memory (lh_host_alloc) for full-rate DMA."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self.lib.lh_host_alloc(nbytes)
        if not p:
            raise LighthandError("lh_host_alloc failed")
        buf = (C.c_char * nbytes).from_address(p)
        arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
        self._keep[arr.ctypes.data] = p
        return arr

    def _io(self, images, fwd_affine, joints, flip, pred, out):
        io = _lib.PipelineIO()
        def ptr(a, dt):
            if a is None:
                return None
            if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags.c_contiguous):
                raise LighthandError(f"HostPipeline: expected a C-contiguous numpy array of dtype {np.dtype(dt)}")
            return a.ctypes.data
        io.images, io.fwd_affine = ptr(images, np.uint8), ptr(fwd_affine, np.float64)
        io.flip, io.joints, io.pred = ptr(flip, np.uint8), ptr(joints, np.float32), ptr(pred, np.float32)
        io.loss, io.coords, io.maxvals = ptr(out["loss"], np.float32), ptr(out["coords"], np.float32), ptr(out["maxvals"], np.float32)
        io.counts, io.sums = ptr(out["counts"], np.uint64), ptr(out["sums"], np.float64)
        return io

    def new_outputs(self):
        return {"loss": np.zeros(1, np.float32), "coords": np.zeros((self.N, self.J, 2), np.float32),
                "maxvals": np.zeros((self.N, self.J), np.float32), "counts": np.zeros(4, np.uint64),
                "sums": np.zeros(4, np.float64)}

    def submit(self, images, fwd_affine, joints, flip=None, pred=None, out=None):
        out = out if out is not None else self.new_outputs()
        io = self._io(images, fwd_affine, joints, flip, pred, out)
        slot = C.c_int(-1)
        check(self.lib.lh_pipeline_submit(self._h, C.byref(io), C.byref(slot)), "lh_pipeline_submit")
        return slot.value, out, (io, images, fwd_affine, joints, flip, pred)

    def wait(self, slot: int):
        check(self.lib.lh_pipeline_wait(self._h, int(slot)), "lh_pipeline_wait")

    def step(self, images, fwd_affine, joints, flip=None, pred=None, out=None):
        slot, out, _keep = self.submit(images, fwd_affine, joints, flip, pred, out)
        self.wait(slot)
        return out

    def reset_metrics(self):
        check(self.lib.lh_pipeline_reset_metrics(self._h), "lh_pipeline_reset_metrics")
