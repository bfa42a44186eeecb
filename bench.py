#!/usr/bin/env python
"""Benchmark of the LightHand heatmap pipeline (BASELINE.json metric: heatmap-pipeline images/s; HBM
roofline fraction of the dominant kernel).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA kernels through the C ABI)
    python bench.py --impl reference [--gpus N] [--steps K] ...    the reference's CPU path (oracle port) on host cores

Workload = BASELINE.json configs[1] ("C2"): training-step pipeline, batch 256 per GPU, 256x256 uint8 RGB
source -> 256x256 float32 network input (cv2-exact affine crop/flip/rotate + mean/std), 21 x 64x64
heat-maps sigma 2: Gaussian targets, JointsMSELoss forward + backward, argmax decode (*4), PCK@0.2 + EPE
accumulation.  One "step" = one pass of that path over one batch.  Synthetic data.

N > 1 is launched by torchrun (one process per GPU); the batch is data-partitioned (weak scaling: 256
images per GPU), the only collective is the all-reduce of the packed metric counters/sums inside the
timed region.  One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "heatmap_pipeline_images_per_s"
UNIT = "images/s"
WORKLOADS = {
    # name: (batch per GPU, S, H, sigma, train)
    "C2": dict(batch=256, S=256, H=64, sigma=2, train=True,
               desc="C2 HRNet training-step heatmap pipeline: batch 256/GPU, 256x256 u8 RGB -> 256x256 f32, 21x64x64 heatmaps sigma=2"),
    "C3": dict(batch=256, S=224, H=56, sigma=2, train=False,
               desc="C3 FreiHAND-shaped eval: batch 256/GPU, 224x224 -> 56x56 heatmaps sigma=2"),
    "C5": dict(batch=128, S=512, H=128, sigma=3, train=True,
               desc="C5 high-resolution stress: batch 128/GPU, 512x512 -> 128x128 heatmaps sigma=3"),
    # C4 is a whole evaluation sweep (strong scaling), see run_eval_sweep
    "C4": dict(batch=256, S=256, H=64, sigma=2, train=False, total=99792,
               desc="C4 LightHand99K-sized eval sweep: 99,792 images batch-sharded over the ranks, 256x256 -> 64x64, "
                    "eval pipeline + PCK curve (mm 0..30) / AUC / EPE, one metric all-reduce"),
}


# ------------------------------------------------------------------------------------------------
# synthetic data (host numpy)
# ------------------------------------------------------------------------------------------------
def synth_host(n, S, seed):
    rng = np.random.default_rng(seed)
    imgs = rng.integers(0, 256, (n, S, S, 3), dtype=np.uint8)
    joints = rng.uniform(0, S, (n, 21, 2)).astype(np.float32)
    off = rng.uniform(size=(n, 21)) < 0.05
    joints[off] = rng.uniform(-S / 8, 9 * S / 8, (int(off.sum()), 2)).astype(np.float32)
    side = 0.8 * S
    bbox = np.tile(np.array([(S - side) / 2, (S - side) / 2, side, side]), (n, 1))
    scale = np.clip(rng.standard_normal(n), -1, 1) * 0.25 + 1
    rot = np.where(rng.uniform(size=n) <= 0.6, np.clip(rng.standard_normal(n), -2, 2) * 45, 0.0)
    flip = rng.uniform(size=n) <= 0.5
    return imgs, joints, bbox, scale, rot, flip


# ------------------------------------------------------------------------------------------------
# CPU path: the oracle port of the reference, one process per host core, each on its own batch
# ------------------------------------------------------------------------------------------------
def _cpu_pass(R, torch, data, S, H, sigma, train):
    """One pass of the reference's per-iteration work over one host batch (the loops the reference runs:
    per-sample warp + normalise and generate_target, then loss[/backward], decode, PCK, EPE)."""
    imgs, joints, bbox, scale, rot, flip, pred = data
    n = imgs.shape[0]
    stride = S / H
    tgt = []
    for i in range(n):
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


def _cpu_worker(args):
    wid, S, H, sigma, train, n, warm, steps, t_start = args
    import torch
    torch.set_num_threads(1)
    try:
        import cv2
        cv2.setNumThreads(1)
    except Exception:
        pass
    from oracle import restate as R
    d = synth_host(n, S, 1000 + wid)
    pred = np.random.default_rng(wid).standard_normal((n, 21, H, H)).astype(np.float32)
    data = (*d, pred)
    for _ in range(warm):
        _cpu_pass(R, torch, data, S, H, sigma, train)
    while time.time() < t_start:          # common start line
        time.sleep(0.001)
    t0 = time.perf_counter()
    for _ in range(steps):
        _cpu_pass(R, torch, data, S, H, sigma, train)
    return time.perf_counter() - t0


def cpu_throughput(wl, steps, warmup, sample_n=32, max_workers=64):
    """images/s of the oracle port with one worker process per host core (fork; call before CUDA init)."""
    import multiprocessing as mp
    import torch  # noqa: F401  (imported before the fork so the workers share the pages; CUDA is not touched)
    from oracle import restate  # noqa: F401
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, max_workers))
    ctx = mp.get_context("fork")
    t_start = time.time() + 3.0 + 0.05 * workers
    jobs = [(w, wl["S"], wl["H"], wl["sigma"], wl["train"], sample_n, warmup, steps, t_start) for w in range(workers)]
    with ctx.Pool(workers) as pool:
        times = pool.map(_cpu_worker, jobs)
    t = max(times)
    return dict(value=workers * sample_n * steps / t, seconds=t, workers=workers, cores_online=cores,
                sample=f"{workers} worker processes x {steps} passes over a {sample_n}-image batch of the same workload "
                       f"(oracle port of the reference's numpy/cv2/torch-CPU loops, 1 thread each)")


def run_reference_arm(args):
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 3))
    # bound the run: one pass over 32 images costs ~0.3-1 s per worker
    steps = min(steps, 20)
    r = cpu_throughput(wl, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * r["seconds"] / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "global_batch": r["workers"] * 32, "parallelism": f"cpu x{r['workers']} processes"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["workers"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(kernel_key):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (C2 only)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_final_traffic.json")) as f:
            t = json.load(f)["kernels"]
        for name, v in t.items():
            if kernel_key in name:
                return v["dram_read_bytes"] + v["dram_write_bytes"]
    except Exception:
        pass
    return None


def bind_to_gpu_cpus(local):
    """Multi-GPU runs: pin this rank to the CPU cores NVML reports as local to its GPU, so the pinned host buffers of
    the e2e leg (first touch) sit on the GPU's NUMA node.  Best effort: any failure leaves the affinity untouched."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cpus = {i for i in range(ncpu) if (int(words[i // 64]) >> (i % 64)) & 1} & allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def run_ours(args):
    wl = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    B, S, H, sigma, train = wl["batch"], wl["S"], wl["H"], wl["sigma"], wl["train"]
    if args.batch:
        B = args.batch

    # CPU baseline first (rank 0, N == 1 only), before CUDA is initialised: the workers are forked
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_throughput(wl, steps=args.cpu_steps, warmup=1)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["workers"], "kind": "port", "sample": r["sample"],
               "seconds": r["seconds"]}

    import torch
    import torch.distributed as dist
    from lighthand_b200 import ops
    from lighthand_b200.pipeline import HeatmapPipeline, HostPipeline
    from lighthand_b200.utils.preprocessing import batch_patch_matrices

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_cpus(local) if world > 1 else None     # pinned e2e buffers land next to this rank's GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- inputs resident in HBM: a ring of R distinct batches (> L2 in total) ------------------------
    ring = max(2, args.ring)
    pipe = HeatmapPipeline(B, (S, S), (S, S), (H, H), sigma=sigma, train=train, materialize_target=args.materialize_target,
                           device=dev)
    stride = S / H
    sets, host_sets = [], []
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    for r_i in range(ring):
        imgs, joints, bbox, scale, rot, flip = synth_host(B, S, seed=(1234 + rank) * 100 + r_i)
        trans, _ = batch_patch_matrices(bbox, flip, scale, rot, (S, S), (S, S))
        M = np.ascontiguousarray(trans.astype(np.float64))
        fl = np.ascontiguousarray(flip.astype(np.uint8))
        d_imgs = torch.from_numpy(imgs).to(dev)
        d_M, d_fl, d_j = torch.from_numpy(M).to(dev), torch.from_numpy(fl).to(dev), torch.from_numpy(joints).to(dev)
        tgt, _ = ops.gen_targets(d_j, (H, H), (stride, stride), sigma, want_weight=False)
        d_pred = tgt + 0.1 * torch.randn(tgt.shape, device=dev, generator=g)       # pred = target + noise
        del tgt
        sets.append(pipe.prepare(d_imgs, d_M, d_j, d_pred, flip=d_fl))
        if r_i < 2:
            host_sets.append((imgs, M, joints, fl))
    resident_bytes = ring * (B * S * S * 3 + B * 21 * H * H * 4)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()

    K, W = args.steps, max(3, args.warmup)
    for i in range(W):
        sets[i % ring].launch()
    torch.cuda.synchronize()
    pipe.reset_metrics()

    # ---- timed region: K steps, CUDA events on the launching stream, metric all-reduce inside ------------
    # two events per step: before K1 and between the two launches; the fused pass of step i ends where step
    # i+1 begins (one more event closes the last step)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
    ev_last = torch.cuda.Event(enable_timing=True)
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    torch.cuda.synchronize()
    t_start.record()
    for i in range(K):
        sets[i % ring].launch(events=ev[i])
    ev_last.record()
    if world > 1:
        from lighthand_b200.distributed import all_reduce_metrics
        all_reduce_metrics([pipe.meter.counts, pipe.meter.sums])
    t_end.record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = t_start.elapsed_time(t_end)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # ---- optional: the two launches on two streams (they touch disjoint data) ---------------------------
    overlap = None
    if args.overlap:
        sa, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        snap = (pipe.meter.counts.clone(), pipe.meter.sums.clone())
        cur = torch.cuda.current_stream(dev)
        for rep in range(2):                      # first pass warms the streams up
            barrier()
            torch.cuda.synchronize()
            o0.record()
            sa.wait_stream(cur); sb.wait_stream(cur)
            for i in range(K):
                sets[i % ring].launch_split(sa.cuda_stream, sb.cuda_stream)
            cur.wait_stream(sa); cur.wait_stream(sb)
            o1.record()
            torch.cuda.synchronize()
        oms = o0.elapsed_time(o1)
        if world > 1:
            t = torch.tensor([oms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            oms = float(t.item())
        overlap = {"value": world * B * K / (oms * 1e-3), "unit": UNIT, "ms_per_step": oms / K,
                   "note": "K1 and the fused pass on two streams (disjoint data): data stage of batch i+1 overlaps the loss stage of batch i"}
        pipe.meter.counts.copy_(snap[0]); pipe.meter.sums.copy_(snap[1])
    metrics = pipe.metrics(all_reduce=False)
    k1_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    fused_ms = float(np.mean([ev[i][1].elapsed_time(ev[i + 1][0] if i + 1 < K else ev_last) for i in range(K)]))
    value = world * B * K / (ms * 1e-3)

    # ---- e2e: host buffers through the C-ABI context, double-buffered, copies inside the timed region -------
    e2e = None
    if not args.no_e2e:
        hp = HostPipeline(B, (S, S), (S, S), (H, H), sigma=sigma, train=train, materialize_target=args.materialize_target,
                          device=local)
        hp.pred.copy_(sets[0]._keep[3])
        pinned = []
        for imgs, M, joints, fl in host_sets:
            a = hp.pinned(imgs.shape, np.uint8); a[...] = imgs
            b = hp.pinned(M.shape, np.float64); b[...] = M
            c = hp.pinned(joints.shape, np.float32); c[...] = joints
            d = hp.pinned(fl.shape, np.uint8); d[...] = fl
            pinned.append((a, b, c, d))
        outs = [hp.new_outputs(), hp.new_outputs()]
        Ke = max(4, min(K, args.e2e_steps))
        def run_e2e(n_steps):
            pend = None
            for i in range(n_steps):
                a, b, c, d = pinned[i % 2]
                cur = hp.submit(a, b, c, flip=d, out=outs[i % 2])
                if pend is not None:
                    hp.wait(pend[0])
                pend = cur
            hp.wait(pend[0])
        run_e2e(4)
        hp.reset_metrics()
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_e2e(Ke)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        h2d = B * S * S * 3 + B * 48 + B + B * 21 * 8
        d2h = 4 + B * 21 * 8 + B * 21 * 4 + 32 + 32
        e2e = {"value": world * B * Ke / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": Ke, "ms_per_step": 1e3 * dt / Ke,
               "note": "lh_pipeline_submit/wait (C ABI) from pinned host buffers, double-buffered; pred stays device-resident "
                       "(it is the backbone's output)"}
        hp.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = load_peaks()
    ab = pipe.algorithmic_bytes_per_image()
    k1_bytes = ab["warp_normalize"] * B
    fused_bytes = ab["heatmap_fused"] * B
    dom = ("warp_normalize", k1_bytes, k1_ms) if k1_ms >= fused_ms else ("heatmap_fused", fused_bytes, fused_ms)
    achieved = dom[1] / (dom[2] * 1e-3) / 1e9
    survey_bytes = {True: {256: 2703780, 224: 2070180, 512: 10813860}, False: {256: 2359716, 224: 1806756, 512: 9437604}}[train].get(S)
    step_ms = ms / K
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8->f32 (fixed-point bilinear in int32, heat-maps f32, reductions f64)", "data": "synthetic",
        "config": {"workload": wl["desc"], "global_batch": world * B, "batch_per_gpu": B,
                   "parallelism": f"dp{world} (batch-sharded; one all-reduce of the metric counters per timed region)",
                   "l2": f"inputs larger than L2: ring of {ring} distinct resident batches = {resident_bytes / 1e6:.0f} MB",
                   "materialize_target": bool(args.materialize_target), "train": bool(train),
                   "launch": "pre-marshalled C-ABI calls, 2 kernels/step (K1; K2+K3+K4+K5 fused)", "normalise": "one FMA per value (<=2 ulp of torch)"},
        "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "frac_of_nominal_8000GBps": achieved / 8000.0,      # north_star quotes the nominal HBM3e figure too
                     "traffic": load_traffic("warp_pairs" if dom[0] == "warp_normalize" else "heatmap_vec") if args.workload == "C2" and not args.batch else None,
                     "traffic_note": "dram__bytes_read+write per launch, ncu --set full (profiles/r01_final_traffic.json); below the algorithmic bytes because part of the output is still dirty in the 126 MB L2 when the kernel ends",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": dom[1], "avg_launch_ms": dom[2]},
        "kernels": {"warp_normalize": {"ms": k1_ms, "GBps": k1_bytes / (k1_ms * 1e-3) / 1e9, "bytes": k1_bytes},
                    "heatmap_fused": {"ms": fused_ms, "GBps": fused_bytes / (fused_ms * 1e-3) / 1e9, "bytes": fused_bytes},
                    "rest_ms": step_ms - k1_ms - fused_ms},
        "pipeline_roofline": {"moved_bytes_per_image": ab["step"], "moved_GBps": ab["step"] * B / (step_ms * 1e-3) / 1e9,
                              "moved_frac_of_peak": ab["step"] * B / (step_ms * 1e-3) / 1e9 / peak,
                              "survey_unfused_bytes_per_image": survey_bytes,
                              "survey_equiv_GBps": (survey_bytes * B / (step_ms * 1e-3) / 1e9) if survey_bytes else None},
        "two_stream": overlap,
        "cpu_baseline": cpu,
        "e2e": e2e,
        "host_affinity_cores": numa,
        "gpu_launches": pipe.launches_per_step * K,
        "clocks": clocks,
        "check": {"pck@0.2": metrics["pck"], "epe_px": metrics["epe"], "loss": metrics["loss"], "n_images": metrics["n_images"]},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_eval_sweep(args):
    """BASELINE.json configs[3]: the full 99,792-image evaluation, batch-sharded across the ranks (strong
    scaling).  Per batch: K1 warp+normalise, fused targets/loss/decode/PCK@0.2/EPE, PCK-curve histogram update;
    at the end ONE all-reduce of the packed counters and the curve state, then AUC / EPE on the host."""
    wl = WORKLOADS["C4"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    from lighthand_b200 import ops
    from lighthand_b200.distributed import all_reduce_metrics, shard_range
    from lighthand_b200.pipeline import HeatmapPipeline
    from lighthand_b200.utils.argparser import CurveAccumulator
    from lighthand_b200.utils.preprocessing import batch_patch_matrices

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_cpus(local) if world > 1 else None     # pinned e2e buffers land next to this rank's GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, S, H, sigma, total = wl["batch"], wl["S"], wl["H"], wl["sigma"], wl["total"]
    lo, hi = shard_range(total, rank, world)
    n_local = hi - lo
    full, tail = divmod(n_local, B)
    ring = max(2, args.ring)
    stride = S / H
    g = torch.Generator(device=dev).manual_seed(4321 + rank)

    def make_sets(pipe, nb, count):
        out = []
        for r_i in range(count):
            imgs, joints, bbox, scale, rot, flip = synth_host(nb, S, seed=(4321 + rank) * 100 + r_i)
            trans, _ = batch_patch_matrices(bbox, flip, scale, rot, (S, S), (S, S))
            vis = (np.random.default_rng(r_i).uniform(size=(nb, 21, 1)) < 0.8).astype(np.float32)
            d_j = torch.from_numpy(np.concatenate([joints, vis], -1)).to(dev)
            tgt, _ = ops.gen_targets(d_j, (H, H), (stride, stride), sigma, want_weight=False)
            d_pred = tgt + 0.1 * torch.randn(tgt.shape, device=dev, generator=g)
            cat = torch.randint(0, 4, (nb,), device=dev, dtype=torch.int32, generator=g)
            step = pipe.prepare(torch.from_numpy(imgs).to(dev), torch.from_numpy(np.ascontiguousarray(trans.astype(np.float64))).to(dev),
                                d_j, d_pred, flip=torch.from_numpy(np.ascontiguousarray(flip.astype(np.uint8))).to(dev))
            out.append((step, d_j, cat))
        return out

    pipe = HeatmapPipeline(B, (S, S), (S, S), (H, H), sigma=sigma, train=False, device=dev)
    sets = make_sets(pipe, B, ring)
    pipe_t = HeatmapPipeline(tail, (S, S), (S, S), (H, H), sigma=sigma, train=False, device=dev) if tail else None
    sets_t = make_sets(pipe_t, tail, 1) if tail else []
    acc = CurveAccumulator([0, 30], "mm", device=dev)

    def sweep():
        for i in range(full):
            step, d_j, cat = sets[i % ring]
            step.launch()
            acc.update(pipe.buffers["coords"], d_j, None, cat)
        if tail:
            step, d_j, cat = sets_t[0]
            step.launch()
            acc.update(pipe_t.buffers["coords"], d_j, None, cat)
            # fold the tail pipeline's counters into the main accumulators
            pipe.meter.counts.add_(pipe_t.meter.counts); pipe.meter.sums.add_(pipe_t.meter.sums)
            pipe_t.reset_metrics()
        if world > 1:
            all_reduce_metrics([pipe.meter.counts, pipe.meter.sums] + acc.tensors())

    def reset():
        pipe.reset_metrics()
        for t in acc.tensors():
            t.zero_()

    for _ in range(max(1, min(args.warmup, 2))):
        sweep()
    torch.cuda.synchronize()
    reps = max(1, min(args.steps, 5))
    times = []
    for _ in range(reps):
        reset()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sweep()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        times.append(ms)
    res = acc.finalize()
    m = pipe.meter.result()
    if rank == 0:
        ms = float(np.median(times))
        line = {
            "metric": METRIC, "value": total / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": reps, "warmup": 2,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8->f32", "data": "synthetic (ring of %d resident batches per rank, reused across the sweep)" % ring,
            "config": {"workload": wl["desc"], "global_batch": total, "batch_per_gpu": B, "images_per_rank": n_local,
                       "parallelism": f"dp{world} (contiguous shards, one all-reduce of counters + curve histograms)"},
            "check": {"n_images": m["n_images"], "n_wrong": m["n_wrong"], "n_total": m["n_total"], "pck@0.2": m["pck"],
                      "epe_px": m["epe"], "mean_auc": float(res["mean_auc"][0]), "mean_epe_mm": float(res["mean_auc"][1]),
                      "curve_count_at_T50": int(acc.hist.sum(0)[:50].sum().item())},
            "gpu_launches": 3 * (full + (1 if tail else 0)),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--ring", type=int, default=4)
    ap.add_argument("--materialize-target", action="store_true")
    ap.add_argument("--no-overlap", dest="overlap", action="store_false",
                    help="skip the extra timing of the step with K1 and the fused pass on two streams")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=70)
    ap.add_argument("--e2e-steps", type=int, default=200)
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: re-launch under torchrun, one rank per GPU (the driver does this itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "C4":
        return run_eval_sweep(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
