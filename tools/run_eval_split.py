#!/usr/bin/env python
"""BASELINE.json configs[4]: the evaluation path over a HAMMER-test-sized synthetic split, sharded across GPUs.

  python tools/run_eval_split.py                                   # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
      tools/run_eval_split.py [--images 120] [--check]

Per rank: its contiguous shard of the images -> polcue_eval_pass_f32 (GT depth->normals stencil, per-image masked
compute_depth_errors for the range mask and every material level of trainer.py:1389-1411, accumulators of the mean over
images: three launches, no host work in between) -> one NCCL all-reduce of 1 + 11 x 7 float64.  --check recomputes
everything unsharded on rank 0 with the CPU oracle and asserts equality (test infrastructure use of oracle/).
The runner is tools/workloads.py::cfg5_eval, which bench.py runs as well.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, ROOT)
from polcue import dist as D  # noqa: E402
import workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=120)        # 10 batches x 12, trainer.py:915-916
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--reps", type=int, default=200)
    args = ap.parse_args()
    rank, local_rank, world = D.init()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    res = workloads.cfg5_eval(rank, world, dev, images=args.images, reps=args.reps, check=args.check)
    if rank == 0:
        print(json.dumps(res), flush=True)
    D.barrier()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
