#!/usr/bin/env python
"""BASELINE.json configs[2] and configs[3].

  --mode sequence (configs[2]): a 10k-frame synthetic HAMMER-shaped sequence (2448x2048 mosaics) sharded across the
      ranks in contiguous blocks, processed in resident chunks of 64 frames, only per-plane float64 checksums kept -- a
      by-product of the fused launch itself, the outputs (551 GB in all) are never read back; one all-reduce at the end.
      Frames come from a pool of distinct Gen-P frames resident on the device and addressed by global frame index,
      so any sharding sees the same sequence.
  --mode loader (configs[3]): the manydepth train-loader path at training resolution: four uint8 planes
      [32, 320, 480] (indoor_dataset.py:435-438) -> XOLP [32,2,320,480] -> get_normals [32,9,320,480], i.e. what
      feeds ShallowEncoder / ShallowNormalsEncoder (pre_encoders.py:49-113), timed alone and followed by torch stand-ins
      of the encoders (tools/encoder_standins.py).  Both modes live in tools/workloads.py, which bench.py runs too.

  --mode loader_full (configs[3] from the stored images): the four FULL-resolution gray images of every sample (HAMMER
      quadrants, 832x1088) in pinned host memory -> H2D -> loader front end (Pillow-exact Lanczos resize to 320x480,
      XOLP, normalizeInput, get_normals), copies double-buffered against the kernels; beside it the reference's own
      CPU path for the same samples (PIL resize x 4 + Iun_and_xolp restatement, one host core per worker).

  python tools/run_sequence.py --mode sequence [--frames 10000]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29519 tools/run_sequence.py
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from polcue import _lib, dist as D, ops, synth  # noqa: E402


def run_sequence(args, rank, world, dev):
    import workloads
    res = workloads.cfg3_sequence(rank, world, dev, frames=args.frames)
    if rank == 0:
        print(json.dumps(res), flush=True)


def run_loader(args, rank, world, dev):
    import workloads
    res = workloads.cfg4_loader(rank, world, dev, reps=args.reps)
    if rank == 0:
        print(json.dumps(res), flush=True)


def _cpu_loader_sample(seed):
    """indoor_dataset.py:335-349 + get_xolp (:430-442) for one sample on the CPU, as a loader worker runs it."""
    import numpy as np
    from PIL import Image
    sys.path.insert(0, ROOT)
    from oracle import polcue_oracle as O
    planes = synth.gen_p_planes(seed, 832, 1088)
    small = [np.asarray(Image.fromarray(p, "L").resize((synth.TRAIN_W, synth.TRAIN_H), Image.LANCZOS)) for p in planes]
    _, rho, phi = O.iun_and_xolp_lstsq(np.stack(small, axis=2), O.CANONICAL_ANGLES)
    return float(rho.sum() + phi.sum())


def run_loader_full(args, rank, world, dev):
    import time
    b, ih, iw, h, w = 32, 832, 1088, synth.TRAIN_H, synth.TRAIN_W
    pool = [torch.stack([torch.from_numpy(synth.gen_p_planes(rank * 8 + i % 8, ih, iw)[k]) for i in range(b)]).pin_memory() for k in range(4)]
    ops.lut_for(1.5, dev)
    copy_stream, main = torch.cuda.Stream(dev), torch.cuda.current_stream(dev)
    slots = [[torch.empty((b, ih, iw), dtype=torch.uint8, device=dev) for _ in range(4)] for _ in range(2)]
    outs = [{}, {}]
    ready = [torch.cuda.Event() for _ in range(2)]
    free = [torch.cuda.Event() for _ in range(2)]

    def run(reps):
        for i in range(reps):
            s = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(free[s])                      # the front end of two steps ago has read slot s
                for k in range(4):
                    slots[s][k].copy_(pool[k], non_blocking=True)
                ready[s].record(copy_stream)
            main.wait_event(ready[s])
            outs[s] = ops.loader_front_end(*slots[s], (h, w), n=1.5, normalize_xolp=ops.XOLP_MEAN_STD, out=outs[s])
            free[s].record(main)

    for e in free:
        e.record(main)
    run(4)
    torch.cuda.synchronize()
    D.barrier()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    run(args.reps)
    e.record()
    torch.cuda.synchronize()
    ms = D.max_over_ranks(a.elapsed_time(e) / args.reps, dev)
    # kernels alone (inputs resident)
    a.record()
    for _ in range(args.reps):
        ops.loader_front_end(*slots[0], (h, w), n=1.5, normalize_xolp=ops.XOLP_MEAN_STD, out=outs[0])
    e.record()
    torch.cuda.synchronize()
    ms_kernels = D.max_over_ranks(a.elapsed_time(e) / args.reps, dev)
    # the C-ABI host entry point: every output (planes, xolp, xolp_norm, normals) back in pinned host memory
    host_out = {}
    for _ in range(2):
        host_out = ops.loader_front_end_host(*pool, (h, w), n=1.5, normalize_xolp=ops.XOLP_MEAN_STD, out=host_out)
    t0 = time.perf_counter()
    host_reps = max(1, args.reps // 5)
    for _ in range(host_reps):
        host_out = ops.loader_front_end_host(*pool, (h, w), n=1.5, normalize_xolp=ops.XOLP_MEAN_STD, out=host_out)
    ms_host = D.max_over_ranks((time.perf_counter() - t0) / host_reps * 1e3, dev)
    if rank == 0:
        import multiprocessing as mp
        cores = os.cpu_count() or 1
        os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = "1"      # trainer.py:9-11
        with mp.get_context("spawn").Pool(cores) as pool_cpu:
            pool_cpu.map(_cpu_loader_sample, range(cores))                       # warm-up (imports)
            t0 = time.perf_counter()
            pool_cpu.map(_cpu_loader_sample, range(2 * cores))
            cpu_s = time.perf_counter() - t0
        print(json.dumps({"config": "cfg4 from the stored images: 4 x u8 [32,832,1088] pinned host -> H2D -> Lanczos resize + XOLP + "
                                    "normalizeInput + get_normals at 320x480 (3 launches per batch)",
                          "n_gpus": world, "us_per_batch": ms * 1e3, "samples_per_s": world * b / (ms * 1e-3),
                          "us_per_batch_kernels_only": ms_kernels * 1e3, "h2d_bytes_per_batch": 4 * b * ih * iw,
                          "h2d_gbs": 4 * b * ih * iw / (ms * 1e-3) / 1e9,
                          "host_entry_point": {"api": "polcue_loader_front_end_u8_host (all outputs returned to pinned host memory)",
                                               "us_per_batch": ms_host * 1e3, "samples_per_s": world * b / (ms_host * 1e-3),
                                               "d2h_bytes_per_batch": b * h * w * (4 + 8 + 8 + 36)},
                          "cpu_reference": {"samples_per_s": 2 * cores / cpu_s, "cores": cores,
                                            "what": "PIL Lanczos resize of the four images + lstsq XOLP per sample, one process per core"}}),
              flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=("sequence", "loader", "loader_full"), default="sequence")
    ap.add_argument("--frames", type=int, default=10000)
    ap.add_argument("--reps", type=int, default=200)
    args = ap.parse_args()
    rank, local_rank, world = D.init()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    {"sequence": run_sequence, "loader": run_loader, "loader_full": run_loader_full}[args.mode](args, rank, world, dev)
    D.barrier()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
