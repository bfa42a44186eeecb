"""The whole hot path as one object.

``HeatmapPipeline``  device-resident step  K1 warp+normalise -> fused K2+K3+K4+K5 (targets, JointsMSELoss
                     forward+backward, decode, PCK/EPE accumulation).  Two kernel launches, no
                     allocation, no host synchronisation; optionally replayed from a CUDA graph.  This is
                     what replaces, per iteration, ``__getitem__``/collate + ``JointsMSELoss`` +
                     ``.cpu()`` + ``get_max_preds`` + ``PCK_2d_loss`` + ``EPE_train`` of the reference's
                     loops (src/utils/method.py:156-287).
                     ``data_stage`` / ``loss_stage`` (or ``step_overlapped``) run the same two launches as a software
                     pipeline: K1 of the NEXT batch on an internal data stream while the loss stage of the CURRENT
                     batch runs on the caller's stream -- the reference's loop shape, where the data loader runs
                     ahead of model + loss (src/utils/method.py:160-183).  Results are bit-identical to ``step``.
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
                 use_graph: bool = False, exact_norm: bool = False, use_vis: bool = False):
        ops.require_cuda()
        self.exact_norm = exact_norm
        self.use_vis = use_vis          # target_weight = visibility * in-bounds (joints [N,J,3]); default: the reference's
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
        # network input, one buffer per pipeline slot: ``step`` uses slot 0, the overlapped form alternates so that K1 of
        # batch i+1 never overwrites what the backbone of batch i is still reading
        self.inputs = [torch.empty((n, 3, *self.out_hw), dtype=f32, device=d) for _ in range(2)]
        self.input = self.inputs[0]
        self._dstream = None            # the C stage ring (data stream + slot events) of the pipelined form, created on first use
        self._ring_h = None
        self._next_slot = 0
        self._pending = []
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
                          metrics=(self.meter.counts, self.meter.sums), pck_T=self.pck_T, use_vis=self.use_vis)

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

    # -- the same step as a two-stage software pipeline ---------------------------------------------
    # Streams and events live in the C library (lh_stage_ring_*): a step costs two library calls and no per-step torch
    # stream / event work.
    def _ring(self):
        if self._dstream is None:
            lib = _lib.load()
            h = C.c_void_p()
            check(lib.lh_stage_ring_create(self.device.index, C.byref(h)), "lh_stage_ring_create")
            self._ring_h, self._ring_lib = h, lib
            self._dstream = torch.cuda.ExternalStream(int(lib.lh_stage_ring_data_stream(h)), device=self.device)
        return self._ring_h

    def __del__(self):
        h = getattr(self, "_ring_h", None)
        if h:
            self._ring_lib.lh_stage_ring_destroy(h)
            self._ring_h = None

    def _take_slot(self) -> int:
        slot = self._next_slot
        self._next_slot ^= 1
        return slot

    def data_stage(self, images_u8: torch.Tensor, fwd_affine: Optional[torch.Tensor], flip: Optional[torch.Tensor] = None,
                   wait_for=None) -> torch.Tensor:
        """Data stage of the NEXT batch: K1 on the ring's data stream into the next input slot.  ``wait_for``: a
        torch.cuda.Event after which the images are valid (default: everything enqueued so far on the current stream;
        ``False``: nothing to wait for -- resident inputs).  Returns the slot's input tensor; it is valid on the current
        stream once ``loss_stage`` (or ``input_ready``) for this batch has been called."""
        if images_u8.shape[0] != self.N:
            raise LighthandError("HeatmapPipeline.data_stage: batch size differs from the one the pipeline was built for")
        r, lib = self._ring(), _lib.load()
        slot = self._take_slot()
        if wait_for is False:
            check(lib.lh_stage_ring_begin_data(r, slot, None, None, 0), "lh_stage_ring_begin_data")
        elif wait_for is None:
            check(lib.lh_stage_ring_begin_data(r, slot, None, torch.cuda.current_stream(self.device).cuda_stream, 2),
                  "lh_stage_ring_begin_data")
        else:
            check(lib.lh_stage_ring_begin_data(r, slot, wait_for.cuda_event, None, 1), "lh_stage_ring_begin_data")
        with torch.cuda.stream(self._dstream):
            ops.warp_normalize(images_u8, fwd_affine, self.out_hw, flip=flip, mean=self.mean, std=self.std,
                               lut=self.lut, interp=self.interp, out=self.inputs[slot], exact_norm=self.exact_norm)
        check(lib.lh_stage_ring_end_data(r, slot), "lh_stage_ring_end_data")
        self._pending.append(slot)
        return self.inputs[slot]

    def input_ready(self) -> torch.Tensor:
        """Make the current stream wait for the oldest pending data stage and return its input tensor (what the
        backbone reads).  ``loss_stage`` then closes that batch."""
        if not self._pending:
            raise LighthandError("HeatmapPipeline.input_ready: no data stage pending")
        slot = self._pending[0]
        check(self._ring_lib.lh_stage_ring_begin_loss(self._ring_h, slot, torch.cuda.current_stream(self.device).cuda_stream),
              "lh_stage_ring_begin_loss")
        return self.inputs[slot]

    def loss_stage(self, pred: torch.Tensor, joints: torch.Tensor, inv_affine: Optional[torch.Tensor] = None) -> dict:
        """Loss stage of the CURRENT batch (the oldest pending data stage) on the current stream: waits for its K1,
        then the fused K2+K3+K4+K5 launch.  Same outputs as ``step``."""
        inp = self.input_ready()
        ops.heatmap_fused(pred, joints, sigma=self.sigma, feat_stride=self.stride,
                          want_target=self.materialize_target, want_grad=self.train, want_decode=True,
                          out_scale=self.stride[0], subpixel=self.subpixel,
                          inv_affine=inv_affine if self.map_back else None,
                          loss_meter=self.meter.loss_meter, buffers=self.buffers,
                          metrics=(self.meter.counts, self.meter.sums), pck_T=self.pck_T, use_vis=self.use_vis)
        slot = self._pending.pop(0)
        check(self._ring_lib.lh_stage_ring_end_loss(self._ring_h, slot, torch.cuda.current_stream(self.device).cuda_stream),
              "lh_stage_ring_end_loss")
        out = dict(self.buffers)
        out["input"] = inp
        out.pop("ws")
        return out

    def step_overlapped(self, next_images_u8, next_fwd_affine, joints, pred, next_flip=None, inv_affine=None,
                        wait_for=None) -> dict:
        """One iteration of the pipelined loop: data stage of the NEXT batch (pass ``next_images_u8=None`` after the
        last one), then the loss stage of the CURRENT batch (whose data stage was the previous call's, or an explicit
        ``data_stage`` before the loop)."""
        if next_images_u8 is not None:
            self.data_stage(next_images_u8, next_fwd_affine, next_flip, wait_for)
        return self.loss_stage(pred, joints, inv_affine)

    def prepare(self, images_u8, fwd_affine, joints, pred, flip=None, inv_affine=None) -> "PreparedStep":
        """Validate and marshal the two C-ABI calls ONCE for a fixed set of device buffers; the returned
        object's ``launch()`` then costs two raw ctypes calls (no per-step Python argument work), which
        keeps the host ahead of a ~150 us device step without needing a CUDA graph."""
        return PreparedStep(self, images_u8, fwd_affine, joints, pred, flip, inv_affine)

    def reset_metrics(self):
        self.meter.reset()

    def metrics(self, all_reduce: bool = True) -> dict:
        """Epoch-end (or mid-epoch) read: optional cross-rank sum (the path's only collective), one D2H.
        Idempotent: the sum is taken over copies, the local accumulators keep this rank's own totals, so a
        second call in the same epoch does not re-reduce already-global numbers."""
        if all_reduce:
            from .distributed import all_reduce_metrics, is_initialized
            if is_initialized():
                counts, sums = self.meter.counts.clone(), self.meter.sums.clone()
                all_reduce_metrics([counts, sums])
                return self.meter.result(counts, sums)
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
        self._k1s = [(images_u8.data_ptr(), sh * sw * 3, sh, sw, sw * 3, ptr(fwd_affine), ptr(flip), None, self._ms,
                      ptr(p.lut), p.interp, ops.LH_CHAN_RGB, ops.LH_LAYOUT_CHW, inp.data_ptr(), n, p.out_hw[0], p.out_hw[1])
                     for inp in p.inputs]
        self._k1 = self._k1s[0]
        self._wargs, self._wrefs = [], []
        for inp in p.inputs:                       # the same K1 call as an lh_warp_args record, one per input slot
            w = _lib.WarpArgs()
            w.src, w.src_img_stride, w.src_h, w.src_w, w.src_pitch = images_u8.data_ptr(), sh * sw * 3, sh, sw, sw * 3
            w.fwd_affine, w.flip, w.chan_gain = ptr(fwd_affine), ptr(flip), None
            w.has_mean_std = int(self._ms is not None)
            if self._ms is not None:
                for i in range(6):
                    w.mean_std[i] = self._ms[i]
            w.lut, w.interp, w.chan_order, w.out_layout = ptr(p.lut), p.interp, ops.LH_CHAN_RGB, ops.LH_LAYOUT_CHW
            w.dst, w.N, w.out_h, w.out_w = inp.data_ptr(), n, p.out_hw[0], p.out_hw[1]
            self._wargs.append(w)
            self._wrefs.append(C.byref(w))
        self._ring_data, self._ring_loss = lib.lh_stage_ring_data, lib.lh_stage_ring_loss
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
        a.use_vis = int(bool(p.use_vis))
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

    def launch_data(self) -> None:
        """Pipelined form, data stage: K1 of THIS prepared batch on the ring's data stream (resident inputs: nothing to
        wait for except the slot's previous consumer).  One library call."""
        p = self._keep[6]
        r = p._ring()
        slot = p._take_slot()
        rc = self._ring_data(r, slot, self._wrefs[slot])
        if rc:
            check(rc, "prepared step (data stage)")
        p._pending.append(slot)

    def launch_loss(self) -> None:
        """Pipelined form, loss stage of THIS prepared batch on the current stream (after its ``launch_data``).  One
        library call."""
        p = self._keep[6]
        slot = p._pending.pop(0)
        rc = self._ring_loss(p._ring_h, slot, self._aref, torch.cuda.current_stream().cuda_stream)
        if rc:
            check(rc, "prepared step (loss stage)")

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

    # the two halves of ``submit`` for a host that runs the backbone in between (on ``loss_stream``)
    def submit_inputs(self, images, fwd_affine, joints, flip=None, pred=None, out=None):
        out = out if out is not None else self.new_outputs()
        io = self._io(images, fwd_affine, joints, flip, pred, out)
        slot = C.c_int(-1)
        check(self.lib.lh_pipeline_submit_inputs(self._h, C.byref(io), C.byref(slot)), "lh_pipeline_submit_inputs")
        return slot.value, out, (io, images, fwd_affine, joints, flip, pred)

    def submit_jpeg(self, streams, fwd_affine, joints, flip=None, pred=None, out=None, backend: str = "gpu_hybrid"):
        """One whole step from COMPRESSED input: ``streams`` = batch JPEG bit-streams (bytes) of src_hw; nvJPEG decodes
        them on the device into the slot's source buffer, then K1 and the loss stage as in ``submit``."""
        from .utils.tsv_file import JpegDecoder
        out = out if out is not None else self.new_outputs()
        io = self._io(None, fwd_affine, joints, flip, pred, out)
        n = len(streams)
        ptrs = (C.c_char_p * n)(*streams)
        sizes = (C.c_size_t * n)(*[len(b) for b in streams])
        slot = C.c_int(-1)
        check(self.lib.lh_pipeline_submit_inputs_jpeg(self._h, ptrs, sizes, JpegDecoder.BACKENDS[backend], C.byref(io),
                                                      C.byref(slot)), "lh_pipeline_submit_inputs_jpeg")
        check(self.lib.lh_pipeline_submit_loss(self._h, slot.value), "lh_pipeline_submit_loss")
        return slot.value, out, (io, streams, ptrs, sizes, fwd_affine, joints, flip, pred)

    def submit_loss(self, slot: int):
        check(self.lib.lh_pipeline_submit_loss(self._h, int(slot)), "lh_pipeline_submit_loss")

    @property
    def loss_stream(self) -> torch.cuda.ExternalStream:
        return torch.cuda.ExternalStream(int(self.lib.lh_pipeline_loss_stream(self._h)), device=f"cuda:{self.cfg.device}")

    def input_slot(self, slot: int):
        return self._view(self.lib.lh_pipeline_input_dev_slot(self._h, int(slot)), (self.N, 3, *self.out_hw))

    def all_reduce_metrics(self, comm) -> None:
        """SUM the context's running accumulators over the ranks of ``comm`` (distributed.MetricComm)."""
        check(self.lib.lh_pipeline_allreduce_metrics(self._h, comm._h), "lh_pipeline_allreduce_metrics")

    def read_metrics(self):
        counts, sums = np.zeros(4, np.uint64), np.zeros(4, np.float64)
        check(self.lib.lh_pipeline_read_metrics(self._h, counts.ctypes.data, sums.ctypes.data), "lh_pipeline_read_metrics")
        return counts, sums

    def step(self, images, fwd_affine, joints, flip=None, pred=None, out=None):
        slot, out, _keep = self.submit(images, fwd_affine, joints, flip, pred, out)
        self.wait(slot)
        return out

    def reset_metrics(self):
        check(self.lib.lh_pipeline_reset_metrics(self._h), "lh_pipeline_reset_metrics")
