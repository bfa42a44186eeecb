#!/usr/bin/env python
"""Host -> device copy ceiling of the box: plain pinned cudaMemcpyAsync of one step's input bytes (50.4 MB for C2), on all
ranks at once.  This is what the e2e number of bench.py is measured against (`e2e.h2d_ceiling_GBps_per_gpu`).

    python tools/h2d_ceiling.py                                              # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_ceiling.py

Prints one JSON line on rank 0: GB/s per GPU (each rank's best of three passes of 12 copies) and the sum."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = 256 * 256 * 256 * 3 + 256 * 48 + 256 + 256 * 21 * 8
    per_gpu = bench._gather_ranks(bench.h2d_ceiling(nbytes, dev, world, dist), dev, world, dist)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps({"n_gpus": world, "bytes_per_copy": nbytes, "GBps_per_gpu": per_gpu, "GBps_total": sum(per_gpu)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
