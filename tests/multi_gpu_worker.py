"""Worker of tests/test_gpu_multi.py, one process per GPU under torchrun: shard-invariance of the whole path on
hardware.  ONE globally indexed synthetic data set; every rank evaluates its ``shard_range`` slice, the accumulators
are summed with the C ABI's grouped NCCL all-reduce, and every rank also evaluates the WHOLE data set alone: the
integer results must be identical bit for bit, the float64 sums to 1e-12 (SURVEY.md 4(4), 8e;
semantics src/utils/method.py:241-250, src/utils/argparser.py:347-378)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from lighthand_b200.distributed import all_reduce_metrics, metric_comm, shard_range, shutdown
    from lighthand_b200.pipeline import HeatmapPipeline
    from lighthand_b200.utils.argparser import CurveAccumulator
    from lighthand_b200.utils.preprocessing import batch_patch_matrices
    from test_gpu_pipeline import synth

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    comm = metric_comm()

    # ---- 1. lh_metrics_allreduce_multi == dist.all_reduce on random buffers ------------------------------------------
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    a = torch.randint(0, 2 ** 40, (4,), dtype=torch.int64, device=dev, generator=g)
    b = torch.randint(0, 2 ** 40, (3, 101), dtype=torch.int64, device=dev, generator=g)
    c = torch.randn(7, dtype=torch.float64, device=dev, generator=g)
    want = [t.clone() for t in (a, b, c)]
    for t in want:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    comm.all_reduce([a, b, c])
    assert torch.equal(a, want[0]) and torch.equal(b, want[1])
    assert torch.allclose(c, want[2], rtol=1e-13, atol=0)

    # ---- 2. the whole path, sharded vs alone --------------------------------------------------------------------------
    S, H, sigma, B, total = 64, 16, 2, 48, 1000          # 1000 = 8 x 125: ragged tails on every rank (125 = 2 x 48 + 29)
    imgs, joints, bbox, scale, rot, flip = synth(total, S, H, sigma, seed=2024)
    trans, _ = batch_patch_matrices(bbox, flip, scale, rot, (S, S), (S, S))
    rng = np.random.default_rng(5)
    vis = (rng.uniform(size=(total, 21, 1)) < 0.8).astype(np.float32)
    cat = rng.integers(0, 4, total).astype(np.int32)
    pred = rng.standard_normal((total, 21, H, H)).astype(np.float32)
    cu = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    d_imgs, d_M, d_fl = cu(imgs), cu(trans.astype(np.float64)), cu(flip.astype(np.uint8))
    d_j, d_cat, d_pred = cu(np.concatenate([joints, vis], -1)), cu(cat), cu(pred)

    def evaluate(lo, hi):
        pipes = {}
        acc = CurveAccumulator([0, 30], "mm", device=dev)
        counts = torch.zeros(4, dtype=torch.int64, device=dev)
        sums = torch.zeros(4, dtype=torch.float64, device=dev)
        for s in range(lo, hi, B):
            e = min(s + B, hi)
            n = e - s
            if n not in pipes:
                pipes[n] = HeatmapPipeline(n, (S, S), (S, S), (H, H), sigma=sigma, train=False, device=dev)
            p = pipes[n]
            out = p.step(d_imgs[s:e], d_M[s:e], d_j[s:e], d_pred[s:e], flip=d_fl[s:e])
            acc.update(out["coords"], d_j[s:e], None, d_cat[s:e])
        for p in pipes.values():
            counts += p.meter.counts
            sums += p.meter.sums
        return counts, sums, acc

    lo, hi = shard_range(total, rank, world)
    counts, sums, acc = evaluate(lo, hi)
    all_reduce_metrics([counts, sums] + acc.tensors())          # ONE grouped NCCL launch on the compute stream
    w_counts, w_sums, w_acc = evaluate(0, total)                 # the 1-rank result, computed here
    torch.cuda.synchronize()
    assert torch.equal(counts, w_counts), (counts.tolist(), w_counts.tolist())
    assert torch.equal(acc.hist, w_acc.hist) and torch.equal(acc.n_valid, w_acc.n_valid) and torch.equal(acc.n_all, w_acc.n_all)
    # sums = [epe_sum, sum(loss_b * n_b), sum(n_b), -]: the EPE sum is a float64 sum of per-joint distances (order only:
    # 1e-12); the loss meter follows AverageMeter.update(loss.item(), n) (src/utils/method.py:171,241) -- a float32 loss PER
    # BATCH, and the shards cut the data set into different batches than the single rank does: float32 rounding, 1e-6
    assert torch.allclose(sums[[0, 2]], w_sums[[0, 2]], rtol=1e-12, atol=0), (sums.tolist(), w_sums.tolist())
    assert torch.allclose(sums[1], w_sums[1], rtol=1e-6, atol=0), (sums.tolist(), w_sums.tolist())
    assert torch.allclose(acc.sum_d, w_acc.sum_d, rtol=1e-12, atol=0)
    r1, r2 = acc.finalize(), w_acc.finalize()
    assert abs(float(r1["mean_auc"][0]) - float(r2["mean_auc"][0])) <= 1e-12

    # ---- 3. HeatmapPipeline.metrics(all_reduce=True) is idempotent ------------------------------------------
    p = HeatmapPipeline(B, (S, S), (S, S), (H, H), sigma=sigma, train=False, device=dev)
    s0 = rank * B
    p.step(d_imgs[s0:s0 + B], d_M[s0:s0 + B], d_j[s0:s0 + B], d_pred[s0:s0 + B], flip=d_fl[s0:s0 + B])
    m1, m2 = p.metrics(all_reduce=True), p.metrics(all_reduce=True)
    assert m1 == m2 and m1["n_images"] == world * B

    if rank == 0:
        print(json.dumps({"ok": True, "world": world, "counts": counts.tolist(), "sums": sums.tolist(),
                          "curve_total": int(acc.hist.sum().item())}), flush=True)
    shutdown()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
