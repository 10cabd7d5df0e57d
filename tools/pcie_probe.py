#!/usr/bin/env python
"""Where is the host<->device ceiling of this box?  Plain copies, no kernels.

  python tools/pcie_probe.py                                         # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 tools/pcie_probe.py

Per rank: a `--mb` MB device buffer and pinned host buffers of three kinds -- torch's `pin_memory()` (cudaHostAlloc),
`polcue.ops.host_empty` (polcue_host_alloc_on: cudaHostAlloc under a NUMA preference for the GPU's node) and the same with
POLCUE_HOST_ALLOC=mapped (huge-page anonymous mapping + cudaHostRegister).  Device->host and host->device copies (cudaMemcpyAsync via `copy_`), each direction alone and both at once,
first with ONE rank copying (the others idle), then with ALL ranks copying at the same time.  One JSON line: GB/s per
GPU and aggregate.  This is the roofline `bench.py` reports for `e2e` and the evidence for its multi-GPU scaling.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200"))
from polcue import _lib, dist as D, ops  # noqa: E402


def read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=2048)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--chunk-mb", type=int, default=0, help="split every copy into chunks of this size (0: one copy)")
    args = ap.parse_args()
    rank, local_rank, world = D.init()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n = args.mb << 20
    d_buf = torch.empty(n, dtype=torch.uint8, device=dev)
    d_src = torch.randint(0, 255, (n,), dtype=torch.uint8, device=dev)
    def mapped():
        os.environ["POLCUE_HOST_ALLOC"] = "mapped"
        try:
            return ops.host_empty((n,), torch.uint8, dev)
        finally:
            os.environ.pop("POLCUE_HOST_ALLOC", None)

    kinds = {"torch_pin_memory": lambda: torch.empty(n, dtype=torch.uint8).pin_memory(),
             "polcue_host_alloc": lambda: ops.host_empty((n,), torch.uint8, dev),
             "polcue_host_alloc_mapped_hugepages": mapped}
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    step = (args.chunk_mb << 20) or n

    def copies(h_in, h_out, do_in, do_out):
        for lo in range(0, n, step):
            if do_in:
                with torch.cuda.stream(s_in):
                    d_buf[lo:lo + step].copy_(h_in[lo:lo + step], non_blocking=True)
            if do_out:
                with torch.cuda.stream(s_out):
                    h_out[lo:lo + step].copy_(d_src[lo:lo + step], non_blocking=True)

    def timed(h_in, h_out, do_in, do_out, active):
        torch.cuda.synchronize()
        D.barrier()
        best = float("inf")
        for _ in range(args.reps):
            D.barrier()
            t0 = time.perf_counter()
            if active:
                copies(h_in, h_out, do_in, do_out)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return D.max_over_ranks(best if active else 0.0, dev)

    result = {"n_gpus": world, "mb_per_copy": args.mb, "chunk_mb": args.chunk_mb, "cores": os.cpu_count(),
              "gpu_numa_node": int(_lib.lib().polcue_host_numa_node(local_rank)),
              "thp": read("/sys/kernel/mm/transparent_hugepage/enabled"), "numa_nodes_online": read("/sys/devices/system/node/online")}
    for name, make in kinds.items():
        h_in, h_out = make(), make()
        h_in.fill_(3)
        assert h_in.is_pinned() and h_out.is_pinned(), name
        copies(h_in, h_out, True, True)
        torch.cuda.synchronize()
        assert torch.equal(h_out[:1 << 20].to(dev), d_src[:1 << 20])
        r = {}
        for label, do_in, do_out in (("h2d", True, False), ("d2h", False, True), ("duplex", True, True)):
            t_one = timed(h_in, h_out, do_in, do_out, rank == 0)
            t_all = timed(h_in, h_out, do_in, do_out, True)
            per_dir = n / 1e9
            r[label] = {"one_rank_gbs_per_direction": per_dir / t_one, "all_ranks_gbs_per_gpu_per_direction": per_dir / t_all,
                        "all_ranks_aggregate_gbs_per_direction": world * per_dir / t_all}
        result[name] = r
        del h_in, h_out
    if rank == 0:
        print(json.dumps(result), flush=True)
    D.barrier()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
