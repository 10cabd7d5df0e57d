#!/usr/bin/env python
"""Benchmark of the fused polarization-cue pipeline (BASELINE.json configs[1] per GPU, weak scaling).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference [...]                          # the reference's CPU algorithm (oracle port)

A step is ONE pass of the hot path over one batch: `--frames` (64) synthetic 2448x2048 polarizer mosaics per GPU
-> quadrant split -> Stokes -> DoLP/AoLP -> three physics normal candidates, one kernel launch.
  value     whole-job mosaic-Mpix/s with inputs and outputs resident in HBM (CUDA events, max over ranks)
  e2e       the same metric through the host-buffer C-ABI call (pinned host mosaics in, pinned host XOLP +
            normals out; H2D and D2H inside the timed region)
  roofline  algorithmic bytes (48 B per output pixel) / mean launch time vs the measured HBM copy peak
  cpu_baseline  the oracle's restatement of the reference chain (lstsq + table interpolation, numpy float64)
            timed on this box's host cores over a bounded sample of the same frames
Under torchrun every rank processes its own `--frames` frames (weak scaling); the only collective is the
all-reduce of an output checksum + the max-over-ranks of the timings.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200")
for _p in (PKG_DIR, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

print_result = print   # replaced in main() by a writer bound to the real stdout

FRAME_H, FRAME_W = 2048, 2448
MPIX_PER_FRAME = FRAME_H * FRAME_W / 1e6
BYTES_PER_OUT_PX = 48            # 4 B read + 8 B XOLP + 36 B normals (SURVEY 8d / BASELINE.md 3)
HBM_FALLBACK_GBS = 6650.0        # /opt/skills/guides/B200_PROFILING.md, used only if MEASURED_PEAKS.json is absent


def baseline_metric():
    try:
        with open(os.path.join(ROOT, "BASELINE.json")) as f:
            return json.load(f)["metric"]
    except Exception:
        return "Mpix/s (frames/s) fused pol-cue+normals at 1/2/4/8 B200; % of HBM roofline"


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """dram bytes per launch of the fused kernel from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "fused_traffic.json")) as f:
            d = json.load(f)
        return d
    except Exception:
        return None


class ClockSampler:
    """Samples SM clock and clock-event (throttle) reasons of one GPU through NVML while kernels run.

    NVML is polled from a thread every ~2 ms (the timed region is only tens of milliseconds long, too short for
    `nvidia-smi -lms`); `mark()` / `stop()` bracket the timed region so its samples can be told from warm-up ones.
    """
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, torch_index):
        self.samples, self.t_mark, self.running, self.err = [], None, False, None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = str(torch.cuda.get_device_properties(torch_index).uuid)
                uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(torch_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.running = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception as e:      # NVML missing: say so instead of inventing numbers
            self.err = repr(e)

    def _poll(self):
        nv = self.nv
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while self.running:
            try:
                self.samples.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                     int(get_reasons(self.h))))
            except Exception as e:
                self.err = repr(e)
                return
            time.sleep(0.002)

    def mark(self):
        self.t_mark = time.perf_counter()

    def stop(self):
        t_end = time.perf_counter()
        self.running = False
        if self.err and not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + self.err]}
        timed = [s for s in self.samples if self.t_mark is not None and self.t_mark <= s[0] <= t_end]
        use = timed if len(timed) >= 3 else self.samples      # warm-up samples run the same kernel back to back
        bits = 0
        for s in use:
            bits |= s[2]
        return {"sm_mhz": statistics.median(s[1] for s in use) if use else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(name for bit, name in self.REASONS.items() if bits & bit),
                "samples": len(use), "samples_in_timed_region": len(timed)}


# --------------------------------------------------------------------------------------------------
# CPU leg: the reference's algorithm (oracle port) on the host cores
# --------------------------------------------------------------------------------------------------
_BLAS_ENV = ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS")
_W = {}


def _cpu_worker_init(n_frames):
    """Runs once in every freshly SPAWNED worker interpreter.  The parent exported *_NUM_THREADS=1 before the spawn, so
    this is the first time numpy / OpenBLAS is loaded in the process and the pin takes effect (setting the variables in
    a forked child of a process that already imported numpy does nothing: round 1's bug).  The effective BLAS thread
    count is read back with threadpoolctl and reported by every task."""
    from oracle import polcue_oracle as O     # test infrastructure; allowed here as the measured CPU baseline only
    from polcue import synth
    import threadpoolctl
    pools = threadpoolctl.threadpool_info()
    _W["O"] = O
    _W["frames"] = [synth.gen_p_mosaic(i) for i in range(n_frames)]
    _W["blas_threads"] = max([int(p.get("num_threads", 1)) for p in pools] or [1])


def _cpu_frame(frame_index):
    mosaic = _W["frames"][frame_index % len(_W["frames"])]
    t0 = time.perf_counter()
    _, rho, _, normals = _W["O"].frame_chain_reference(mosaic, 1.5)
    return time.perf_counter() - t0, float(rho.sum()) + float(normals[2].sum()), _W["blas_threads"]


class CpuReferencePool:
    """One worker process per host core, BLAS / OpenMP pinned to ONE thread each, as the reference pins them
    (manydepth/trainer.py:9-11) and as its DataLoader workers run (num_workers, options.py:299-302)."""

    def __init__(self, frames_per_step):
        import multiprocessing as mp
        self.cores = os.cpu_count() or 1
        self.workers = max(1, min(self.cores, frames_per_step))
        saved = {k: os.environ.get(k) for k in _BLAS_ENV}
        for k in _BLAS_ENV:
            os.environ[k] = "1"                        # inherited by the spawned interpreters
        try:
            self.pool = mp.get_context("spawn").Pool(self.workers, initializer=_cpu_worker_init, initargs=(4,))
            self.step(self.workers)                    # every worker has imported numpy/scipy and built its frames
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v

    def step(self, frames):
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_frame, range(frames), chunksize=1)
        sec = time.perf_counter() - t0
        self.blas_threads = max(r[2] for r in res)
        if self.blas_threads != 1:
            raise RuntimeError(f"CPU baseline workers run BLAS with {self.blas_threads} threads; the reference pins 1")
        return sec

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_reference_run(frames_per_step, steps, warmup, budget_s=1500.0):
    """Times `steps` passes of `frames_per_step` frames through the reference chain on all host cores.
    Returns (Mpix/s, workers, seconds per step, observed BLAS threads per worker, steps actually timed)."""
    pool = CpuReferencePool(frames_per_step)
    try:
        first = None
        for _ in range(warmup):
            first = pool.step(frames_per_step)
        if first is not None and first * (steps + warmup) > budget_s:      # a slow box: keep the run inside the driver's limit
            steps = max(1, int(budget_s / first) - warmup)
        times = [pool.step(frames_per_step) for _ in range(steps)]
    finally:
        pool.close()
    sec = sum(times) / len(times)
    return frames_per_step * MPIX_PER_FRAME / sec, pool.workers, sec, pool.blas_threads, len(times)


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    frames = args.frames                                   # the SAME step as the CUDA arm: 64 frames of cfg2
    value, workers, sec, blas, steps = cpu_reference_run(frames, args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": baseline_metric(), "value": value, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": max(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2: fused split+XOLP+3 physics normal candidates on synthetic 2448x2048 mosaics (Gen-P), n=1.5",
                   "frames_per_step": frames, "frame": [FRAME_H, FRAME_W], "frames_per_s": frames / sec},
        "cpu_baseline": {"value": value, "unit": "Mpix/s", "cores": workers, "kind": "port",
                         "sample": f"{frames} frames per step x {steps} steps ({sec:.1f} s per step), one spawned process per core, "
                                   f"BLAS/OpenMP threads per process observed with threadpoolctl: {blas}; "
                                   "oracle/polcue_oracle.frame_chain_reference (lstsq XOLP + 3 table interpolations + normals)"},
        "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_result(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# GPU leg
# --------------------------------------------------------------------------------------------------
def pcie_probe(dev, h_in, d_in, h_out, d_out, D, reps=3):
    """The roofline of `e2e`: plain cudaMemcpyAsync copies (torch `copy_` from / to pinned memory) of exactly the bytes one
    e2e step moves -- the mosaics host->device and XOLP + normals device->host -- alone and then both directions at once
    on two streams (PCIe is full duplex), on all ranks at the same time.  Wall-clock around a device synchronise (each
    copy is tens of milliseconds), max over ranks.  Returns GB/s per GPU and the duplex time of one step's bytes."""
    import torch
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    d2h_bytes = sum(t.numel() * t.element_size() for t in h_out.values())
    h2d_bytes = h_in.numel() * h_in.element_size()

    s_out2 = torch.cuda.Stream(dev)

    def run(do_in, do_out, split=False):
        torch.cuda.synchronize()
        D.barrier()
        best = float("inf")
        for _ in range(reps):
            D.barrier()
            t0 = time.perf_counter()
            if do_in:
                with torch.cuda.stream(s_in):
                    d_in.copy_(h_in, non_blocking=True)
            if do_out and not split:
                with torch.cuda.stream(s_out):
                    for key, t in h_out.items():
                        t.copy_(d_out[key], non_blocking=True)
            elif do_out:                                        # the same bytes as frame-sized pieces alternating between two streams
                k = 0                                           # (how the host pipeline returns its chunks)
                for key, t in h_out.items():
                    for f in range(t.shape[0]):
                        with torch.cuda.stream(s_out if k % 2 == 0 else s_out2):
                            t[f].copy_(d_out[key][f], non_blocking=True)
                        k += 1
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return D.max_over_ranks(best, dev)

    run(True, True)                                             # warm-up
    t_in, t_out, t_both = run(True, False), run(False, True), run(True, True)
    t_out_split = run(False, True, split=True)
    return {"h2d_gbs_alone": h2d_bytes / t_in / 1e9, "d2h_gbs_alone": d2h_bytes / min(t_out, t_out_split) / 1e9,
            "d2h_gbs_alone_whole_tensors_one_stream": d2h_bytes / t_out / 1e9,
            "d2h_gbs_alone_frame_pieces_two_streams": d2h_bytes / t_out_split / 1e9,
            "duplex_ms_for_one_step": t_both * 1e3, "d2h_gbs_duplex": d2h_bytes / t_both / 1e9,
            "h2d_bytes": h2d_bytes, "d2h_bytes": d2h_bytes}


def run_polcue_arm(args, rank, local_rank, world):
    import torch
    from polcue import _lib, dist as D, ops, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (polcue has no CPU path); use --impl reference for the CPU baseline")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    D.init("nccl")
    ops.set_default_trig(args.trig)               # which table handle (sincos variant) this process hands to the library

    B, H, W = args.frames, FRAME_H, FRAME_W
    hs, ws = H // 2, W // 2
    out_px = B * hs * ws
    first = rank * B                                            # every rank owns its own frames of the sequence
    if args.gen == "P":
        mosaic = synth.gen_p_batch_torch(first, B, H, W, device=dev)
    else:
        mosaic = torch.stack([torch.from_numpy(synth.gen_u_mosaic(first + i, H, W)) for i in range(B)]).to(dev)
    out = {"xolp": torch.empty((B, 2, hs, ws), dtype=torch.float32, device=dev),
           "normals": torch.empty((B, 9, hs, ws), dtype=torch.float32, device=dev)}
    ops.lut_for(1.5, dev)
    torch.cuda.synchronize()

    def step():
        ops.fused_mosaic(mosaic, 1.5, out=out)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    t_warm = time.perf_counter()
    done = 0
    while done < max(args.warmup, 3) or time.perf_counter() - t_warm < args.warmup_seconds:   # clocks settle under load
        step()
        done += 1
        if done % 16 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    D.barrier()
    if sampler:
        sampler.mark()
    launches0 = _lib.launch_count()
    events = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    torch.cuda.synchronize()
    events[0].record()
    for i in range(args.steps):
        step()
        events[i + 1].record()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - launches0
    D.barrier()
    total_ms = events[0].elapsed_time(events[-1])
    per_launch_ms = [events[i].elapsed_time(events[i + 1]) for i in range(args.steps)]
    ms_per_step = D.max_over_ranks(total_ms / args.steps, dev)
    clocks = sampler.stop() if sampler else None

    # ---- the same step after `--sustained-seconds` of continuous load (clocks under the 1 kW power cap) ----
    sustained = None
    if args.sustained_seconds > 0:
        s_sampler = ClockSampler(local_rank) if rank == 0 else None
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < args.sustained_seconds:
            step()
            n += 1
            if n % 16 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        D.barrier()
        if s_sampler:
            s_sampler.mark()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(args.steps):
            step()
        ev[1].record()
        torch.cuda.synchronize()
        s_ms = D.max_over_ranks(ev[0].elapsed_time(ev[1]) / args.steps, dev)
        sustained = {"ms_per_step": s_ms, "value": world * B * MPIX_PER_FRAME / (s_ms * 1e-3), "unit": "Mpix/s",
                     "after_seconds_of_load": args.sustained_seconds, "clocks": s_sampler.stop() if s_sampler else None}

    # output checksum: the one collective of the sequence benchmark
    chk = torch.stack((out["xolp"].double().sum(), out["normals"].double().sum()))
    D.all_reduce_sums(chk)

    # ---- e2e: host buffers through the public host entry point -------------------------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h_mosaic = ops.host_empty((B, H, W), torch.uint8, dev)       # pinned, on the NUMA node of this rank's GPU (polcue_host_alloc_on)
    h_mosaic.copy_(mosaic)
    h_out = {"xolp": ops.host_empty((B, 2, hs, ws), torch.float32, dev),
             "normals": ops.host_empty((B, 9, hs, ws), torch.float32, dev)}
    ops.fused_mosaic_host(h_mosaic, 1.5, out=h_out)             # warm-up: allocates the device ring
    torch.cuda.synchronize()
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ops.fused_mosaic_host(h_mosaic, 1.5, out=h_out)         # blocks until the results are in host memory
    e2e_ms = D.max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps, dev)
    same = True
    for key in ("xolp", "normals"):                             # EVERY frame of the host path equals the device path, bit for bit
        back = h_out[key].to(dev)
        same = same and bool(torch.equal(back, out[key]))
        del back
    pcie = pcie_probe(dev, h_mosaic, mosaic, h_out, out, D)     # plain cudaMemcpyAsync of the step's bytes: the e2e roofline

    # ---- informational: the training-loader usage -- host mosaics in, outputs stay in HBM for the encoders, the host
    #      reads back only the per-plane float64 checksums (11 doubles) ---------------------------------------------
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        mosaic.copy_(h_mosaic, non_blocking=True)              # H2D from pinned memory on the compute stream
        plane_sums = ops.fused_mosaic(mosaic, 1.5, out=out, want_stats=True)["stats13"].cpu()   # checksums: by-product of the launch; D2H + sync
    res_ms = D.max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps, dev)
    del plane_sums

    # ---- the other BASELINE configs, driver-visible: cfg3 (10k-frame sequence), cfg4 (train-loader path + encoders),
    #      cfg5 (evaluation split).  tools/workloads.py holds the runners; each block carries its own parity check. ----
    extra = {}
    if not args.no_extra_configs:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import workloads
        del h_out, h_mosaic
        try:
            extra["cfg3"] = workloads.cfg3_sequence(rank, world, dev, frames=args.sequence_frames, pool=mosaic if args.gen == "P" and rank == 0 and world == 1 else None)
            extra["cfg4"] = workloads.cfg4_loader(rank, world, dev, reps=200)
            extra["cfg5"] = workloads.cfg5_eval(rank, world, dev, images=120, reps=200, check=True)
        except Exception as e:                                  # never lose the headline line over an auxiliary block
            extra["error"] = repr(e)

    if rank != 0:
        return
    peak, peak_src = hbm_peak()
    mean_launch_ms = sum(per_launch_ms) / len(per_launch_ms)
    achieved = BYTES_PER_OUT_PX * out_px / (mean_launch_ms * 1e-3) / 1e9
    traffic = recorded_traffic()
    value = world * B * MPIX_PER_FRAME / (ms_per_step * 1e-3)
    line = {
        "metric": baseline_metric(), "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: fused split+XOLP+3 physics normal candidates on synthetic 2448x2048 mosaics "
                               f"(Gen-{args.gen}), n=1.5", "frames_per_gpu_per_step": B, "frame": [H, W],
                   "frames_per_s": world * B / (ms_per_step * 1e-3), "outputs": "xolp f32 [B,2,Hs,Ws] + normals f32 [B,9,Hs,Ws]",
                   "l2_policy": f"working set {BYTES_PER_OUT_PX * out_px / 1e9:.2f} GB per step >> 126 MB L2 (no flush needed)",
                   "trig": args.trig, "checksum": [float(chk[0]), float(chk[1])]},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (traffic or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": BYTES_PER_OUT_PX * out_px, "kernel": "fused_mosaic_kernel<4>",
                     "launch_ms": mean_launch_ms},
        "e2e": {"value": world * B * MPIX_PER_FRAME / (e2e_ms * 1e-3), "unit": "Mpix/s", "ms_per_step": e2e_ms,
                "steps": e2e_steps, "h2d_bytes_per_step": B * H * W, "d2h_bytes_per_step": 44 * out_px,
                "api": "polcue.ops.fused_mosaic_host -> polcue_fused_mosaic_u8_host (pinned host buffers from polcue_host_alloc_on)",
                "matches_device_path": same, "matches_device_path_frames_compared": B,
                "roofline": {"bound": "pcie", "achieved": 44 * out_px / (e2e_ms * 1e-3) / 1e9, "peak": pcie["d2h_gbs_alone"],
                             "unit": "GB/s per GPU, device->host (the dominant direction: 3.53 of the step's 3.85 GB)",
                             "frac": 44 * out_px / (e2e_ms * 1e-3) / 1e9 / pcie["d2h_gbs_alone"],
                             "probe": pcie, "ranks_copying_concurrently": world,
                             "host_numa_node_of_gpu": int(_lib.lib().polcue_host_numa_node(local_rank)),
                             "note": "peak = best plain cudaMemcpyAsync of the step's device->host bytes (whole tensors on one stream, or "
                                     "frame-sized pieces on two) with nothing else running on this GPU's link, all ranks copying "
                                     "concurrently; the probe also times both directions at once"}},
        "e2e_device_resident_outputs": {
            "value": world * B * MPIX_PER_FRAME / (res_ms * 1e-3), "unit": "Mpix/s", "ms_per_step": res_ms, "steps": e2e_steps,
            "h2d_bytes_per_step": B * H * W, "d2h_bytes_per_step": 13 * 8,
            "note": "informational, not the headline: pinned host mosaics -> H2D -> fused kernel -> outputs stay in HBM (as the "
                    "encoders consume them, pre_encoders.py:89-97) -> per-plane float64 checksums (a by-product of the same launch) "
                    "read back to the host"},
        "gpu_launches": launches,
        "clocks": clocks,
        "sustained": sustained,
    }
    line.update(extra)
    if world == 1 and not args.no_cpu_baseline:
        frames = max(1, min(2 * (os.cpu_count() or 1), B))      # bounded sample: two frames per core (~10-20 s of CPU work)
        v, workers, sec, blas, passes = cpu_reference_run(frames, 2, 1)
        line["cpu_baseline"] = {"value": v, "unit": "Mpix/s", "cores": workers, "kind": "port",
                                "sample": f"{frames} of the step's {B} Gen-P frames per pass x {passes} passes ({sec:.1f} s per pass), one "
                                          f"spawned process per core, BLAS/OpenMP threads per process observed with threadpoolctl: "
                                          f"{blas}; oracle/polcue_oracle.frame_chain_reference"}
    print_result(json.dumps(line))


def claim_stdout():
    """Keep stdout for the ONE JSON line: everything else written to fd 1 during the run (NCCL's version banner,
    library chatter) is sent to stderr.  Returns a file object on the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    real_stdout = claim_stdout()
    global print_result
    print_result = lambda line: (real_stdout.write(line + "\n"), real_stdout.flush())
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("polcue", "reference"), default="polcue")
    ap.add_argument("--frames", type=int, default=64, help="frames per GPU per step (BASELINE configs[1]: 64)")
    ap.add_argument("--gen", choices=("P", "U"), default="P", help="synthetic generator: physical (P) or uniform stress (U)")
    ap.add_argument("--trig", choices=("poly", "mufu"), default="mufu")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--warmup-seconds", type=float, default=0.0,
                    help="extra warm-up time before the timed region (0: exactly --warmup steps, like the peak measurement)")
    ap.add_argument("--sustained-seconds", type=float, default=2.0,
                    help="also report ms/step after this long under continuous load (power-capped clocks); 0 disables")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the cfg3 / cfg4 / cfg5 blocks")
    ap.add_argument("--sequence-frames", type=int, default=10000, help="cfg3: frames of the sequence (sharded over the GPUs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
    else:
        run_polcue_arm(args, rank, local_rank, world)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
