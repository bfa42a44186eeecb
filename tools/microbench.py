#!/usr/bin/env python
"""Developer micro-benchmark: per-op device timings (CUDA events, ring of cold buffers) to see how far
each kernel variant is from the HBM copy roofline.  Not part of bench.py's contract."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lighthand_b200 import ops  # noqa: E402


def timeit(fn, ring, iters=30, warm=2):
    """Median device time of one call, us.  The ring of calls is captured into a CUDA graph so the host-side
    wrapper cost (tens of us in Python) cannot hide in the measurement."""
    for i in range(ring):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(ring):
            fn(i)
    for _ in range(warm):
        g.replay()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record()
        g.replay()
        b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2] * 1e3 / ring


def main():
    N, J, H, S, R = 256, 21, 64, 256, 4
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    joints = [torch.rand((N, J, 2), device=dev, generator=g) * S for _ in range(R)]
    pred = [torch.randn((N, J, H, H), device=dev, generator=g) for _ in range(R)]
    tgt = [ops.gen_targets(j, want_weight=False)[0] for j in joints]
    imgs = [torch.randint(0, 256, (N, S, S, 3), device=dev, dtype=torch.uint8, generator=g) for _ in range(R)]
    th = torch.rand(N, device=dev, generator=g) * 1.5 - 0.75
    sc = 0.8 * (1 + 0.25 * (torch.rand(N, device=dev, generator=g) * 2 - 1))
    M = torch.zeros((N, 2, 3), dtype=torch.float64, device=dev)
    c, s_ = torch.cos(th).double() / sc.double(), torch.sin(th).double() / sc.double()
    M[:, 0, 0], M[:, 0, 1], M[:, 1, 0], M[:, 1, 1] = c, -s_, s_, c
    ctr = S / 2
    M[:, 0, 2] = ctr - (M[:, 0, 0] * ctr + M[:, 0, 1] * ctr)
    M[:, 1, 2] = ctr - (M[:, 1, 0] * ctr + M[:, 1, 1] * ctr)
    eye = torch.tensor([[1.0, 0, 0], [0, 1.0, 0]], dtype=torch.float64, device=dev).repeat(N, 1, 1)
    out_img = torch.empty((N, 3, S, S), device=dev)
    bufs = {"grad": torch.empty_like(pred[0]), "target": torch.empty_like(pred[0])}
    hm_bytes = N * J * H * H * 4
    k1_bytes = N * S * S * 15
    copy_dst = torch.empty_like(pred[0])
    res = {}

    def rec(name, us, nbytes):
        res[name] = {"us": round(us, 2), "GBps": round(nbytes / us / 1e3, 1)}

    rec("torch_copy_88MB", timeit(lambda i: copy_dst.copy_(pred[i]), R), 2 * hm_bytes)
    rec("decode_only(read)", timeit(lambda i: ops.decode(pred[i], 4.0), R), hm_bytes)
    rec("gen_targets(write)", timeit(lambda i: ops.gen_targets(joints[i], want_weight=False, out=bufs["target"]), R), hm_bytes)
    rec("fused loss+grad+decode", timeit(lambda i: ops.heatmap_fused(pred[i], joints[i], buffers=dict(bufs)), R), 2 * hm_bytes)
    rec("fused loss+decode (no grad)", timeit(lambda i: ops.heatmap_fused(pred[i], joints[i], want_grad=False, buffers=dict(bufs)), R), hm_bytes)
    rec("fused loss+grad (no decode)", timeit(lambda i: ops.heatmap_fused(pred[i], joints[i], want_decode=False, buffers=dict(bufs)), R), 2 * hm_bytes)
    rec("fused +target", timeit(lambda i: ops.heatmap_fused(pred[i], joints[i], want_target=True, buffers=dict(bufs)), R), 3 * hm_bytes)
    rec("jmse (target_in)", timeit(lambda i: ops.jmse_loss(pred[i], tgt[i]), R), 3 * hm_bytes)
    for name, mat in (("warp rot+scale", M), ("warp identity", eye)):
        rec(name + " fma", timeit(lambda i: ops.warp_normalize(imgs[i], mat, (S, S), out=out_img, exact_norm=False), R), k1_bytes)
    const_imgs = [torch.full_like(imgs[0], 128) for _ in range(R)]
    rec("warp rot+scale fma, CONSTANT source", timeit(lambda i: ops.warp_normalize(const_imgs[i], M, (S, S), out=out_img, exact_norm=False), R), k1_bytes)
    lut = ops.build_norm_lut(device=dev)
    rec("warp rot+scale lut", timeit(lambda i: ops.warp_normalize(imgs[i], M, (S, S), out=out_img, lut=lut), R), k1_bytes)
    rec("resize torch-mode", timeit(lambda i: ops.warp_normalize(imgs[i], None, (S, S), out=out_img, interp=ops.LH_INTERP_TORCH), R), k1_bytes)
    imgs224 = [torch.randint(0, 256, (N, 224, 224, 3), device=dev, dtype=torch.uint8, generator=g) for _ in range(R)]
    rec("resize torch-mode 224->256 (LightHand99K path)",
        timeit(lambda i: ops.warp_normalize(imgs224[i], None, (S, S), out=out_img, interp=ops.LH_INTERP_TORCH), R),
        N * (224 * 224 * 3 + S * S * 12))
    # FreiHAND image path (a-3): 224x224 sources, train-time augmentation mix (40 % rotated), 224 -> 256
    import numpy as np
    from lighthand_b200.datasets import frei_dataloader as F
    from lighthand_b200.utils import image_ops
    rng = np.random.RandomState(0)
    fimgs = [torch.randint(0, 256, (N, 224, 224, 3), device=dev, dtype=torch.uint8, generator=g) for _ in range(R)]
    augm = [F.augm_params(True, rng) for _ in range(N)]
    centers = 112 + rng.uniform(-10, 10, (N, 2))
    recs, sfl, mpx = image_ops.crop_descriptors(centers, [a[3] for a in augm], [224, 224], [a[2] for a in augm], [a[1] for a in augm])
    desc = image_ops.descriptors_to_device(recs, dev)
    rgb = torch.empty((N, 3, 224, 224), device=dev)
    lib = ops._lib.load()
    scratch = torch.empty(max(sfl, 1), device=dev)
    rot = torch.tensor([i for i, r in enumerate(recs) if r.rotated], dtype=torch.int32, device=dev)
    def frei_rgb(i):
        ops.check(lib.lh_frei_rgb_processing(fimgs[i].data_ptr(), 224 * 224 * 3, 224, 224, 224 * 3, desc.data_ptr(), N, 224,
                                             rot.data_ptr(), rot.numel(), scratch.data_ptr(), mpx[0], mpx[1], rgb.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream))
    frei_bytes = N * (224 * 224 * 3 + 224 * 224 * 12)
    rec("frei rgb_processing (rotate+resize)", timeit(frei_rgb, R), frei_bytes)
    rec("frei Resize+Normalize+BGR f32", timeit(lambda i: ops.resize_normalize_f32(rgb, (S, S), chan_order=ops.LH_CHAN_BGR, out=out_img), R),
        N * (224 * 224 * 12 + S * S * 12))
    rec("frei fused rgb_processing+Resize+Normalize+BGR (lh_frei_fused)",
        timeit(lambda i: ops.frei_fused(fimgs[i], desc, 224, sfl, mpx, (S, S), chan_order=ops.LH_CHAN_BGR, rot_index=rot, out=out_img), R),
        N * (224 * 224 * 3 + S * S * 12))
    # the small entry points (latency-bound: microseconds per launch, bytes per second mean little)
    from lighthand_b200.tools.dataset import color_jitter_params
    torch.manual_seed(0)
    jp = color_jitter_params([i % 2 == 0 for i in range(N)])
    unit = torch.rand((N, 3, S, S), device=dev, generator=g)
    d_jp = torch.from_numpy(np.frombuffer(bytes(jp), dtype=np.uint8).copy()).to(dev)
    jws = torch.empty(N * ops._lib.LH_JITTER_PARTS, dtype=torch.float64, device=dev)
    ms = ops._mean_std(ops.IMAGENET_MEAN, ops.IMAGENET_STD)
    rec("color jitter + normalize (half of the batch augmented; parameters already on the device)",
        timeit(lambda i: ops.check(lib.lh_color_jitter_normalize(unit.data_ptr(), d_jp.data_ptr(), N, S, S, ms, out_img.data_ptr(),
                                                                  jws.data_ptr(), torch.cuda.current_stream().cuda_stream)), 1),
        2 * N * 3 * S * S * 4)
    jv = torch.cat([joints[0], torch.ones((N, J, 1), device=dev)], dim=2)
    flip = (torch.arange(N, device=dev) % 2).to(torch.uint8)
    rec("transform_joints", timeit(lambda i: ops.transform_joints(jv, M, (S, S), flip=flip, orig_w=S), 1), N * J * 24)
    rec("bbox_diag", timeit(lambda i: ops.bbox_diag(jv), 1), N * J * 12)
    xy = joints[1].contiguous()
    counts, sums = ops.new_metric_state(dev)
    rec("pck_epe_batch", timeit(lambda i: ops.pck_epe_batch(xy, joints[0], 0.2, "proportion", False, counts, sums), 1), N * J * 16)
    thr = torch.arange(0, 31, dtype=torch.float64, device=dev)
    hist = torch.zeros((1, 32), dtype=torch.int64, device=dev)
    nv, na, sd = (torch.zeros(1, dtype=torch.int64, device=dev), torch.zeros(1, dtype=torch.int64, device=dev),
                  torch.zeros(1, dtype=torch.float64, device=dev))
    rec("pck_curve_accum (31 thresholds)", timeit(lambda i: ops.pck_curve_accum(xy, jv, thr, "mm", False, hist, nv, na, sd), 1),
        N * J * 20)
    big_src = torch.empty((N, 3, S, S), device=dev)
    rec("torch_copy_201MB", timeit(lambda i: out_img.copy_(big_src), 1), 2 * N * 3 * S * S * 4)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
