"""The measured workloads of BASELINE.json configs[2..4], shared by bench.py (driver-run) and the tools/run_*.py CLIs.

Each function runs on every rank of a torchrun job (one process per GPU), times on the device with CUDA events, takes
the max over ranks and returns a small dict on every rank (rank 0 prints it).  Nothing here touches oracle/ except the
explicitly named `check` helpers, which are test infrastructure (the checker, never the thing measured).
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)
from polcue import _lib, dist as D, ops, synth  # noqa: E402

HBM_FALLBACK_GBS = 6650.0


def hbm_peak():
    try:
        import json
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return HBM_FALLBACK_GBS


# --------------------------------------------------------------------------------------------------
# cfg3: the 10k-frame sequence, sharded, chunks of 64, checksums as a by-product of the fused launch
# --------------------------------------------------------------------------------------------------
def cfg3_sequence(rank, world, dev, frames=10000, chunk=64, pool=None):
    """BASELINE configs[2]: frames [lo, hi) of a `frames`-frame synthetic HAMMER-shaped sequence per rank (contiguous
    shards, strong scaling), resident chunks of 64 mosaics -> fused kernel with the statistics by-product (13 float64
    sums per launch, accumulated on the device; the 551 GB of outputs are overwritten chunk after chunk and never read
    back), one all-reduce of the sums at the end.  Frame f of the sequence is pool[f % 64] (64 distinct Gen-P frames
    resident on every rank), so any sharding sees the same sequence."""
    pool_n = 64
    if pool is None:
        pool = synth.gen_p_batch_torch(0, pool_n, device=dev)
    lo, hi = D.shard_range(frames, rank, world)
    hs, ws = synth.FRAME_H // 2, synth.FRAME_W // 2
    out = {"xolp": torch.empty((chunk, 2, hs, ws), dtype=torch.float32, device=dev),
           "normals": torch.empty((chunk, 9, hs, ws), dtype=torch.float32, device=dev)}
    sums = torch.zeros(13, dtype=torch.float64, device=dev)
    ops.lut_for(1.5, dev)

    def run_chunk(first, n, acc):
        """Frames [first, first + n) of the sequence.  The sequence cycles through the resident pool, so a chunk is one or
        two contiguous VIEWS of it (no gather copy): pool[o : o + n] and, when it wraps, the head of the pool."""
        done = 0
        while done < n:
            o = (first + done) % pool_n
            m = min(n - done, pool_n - o)
            res = ops.fused_mosaic(pool[o:o + m], 1.5, want_stats=True,
                                   out={"xolp": out["xolp"][done:done + m], "normals": out["normals"][done:done + m], "stats13": out.get("stats13")})
            out["stats13"] = res["stats13"]
            acc += res["stats13"]
            done += m

    scratch = torch.zeros_like(sums)
    for _ in range(3):                                    # warm-up: clocks, table upload, workspace allocation
        run_chunk(0, chunk, scratch)
    torch.cuda.synchronize()
    D.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.launch_count()
    a.record()
    for first in range(lo, hi, chunk):
        run_chunk(first, min(chunk, hi - first), sums)
    b.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - launches0
    ms = D.max_over_ranks(a.elapsed_time(b), dev)
    D.all_reduce_sums(sums)
    # size-independent parity property: the sequence repeats the 64-frame pool, so its checksum is frames/64 x the pool's
    one = torch.zeros_like(sums)
    run_chunk(0, chunk, one)
    whole, rest = divmod(frames, pool_n)
    expect = one * whole
    if rest:
        part = torch.zeros_like(sums)
        run_chunk(whole * pool_n, rest, part)
        expect = expect + part
    rel = float(((sums - expect).abs() / expect.abs().clamp_min(1e-300)).max())
    mpix = frames * synth.FRAME_H * synth.FRAME_W / 1e6
    return {"workload": f"cfg3: {frames}-frame synthetic sequence of 2448x2048 mosaics, contiguous shards over {world} GPU(s), chunks of "
                        f"{chunk}, per-plane float64 checksums as a by-product of the fused launch (outputs never re-read)",
            "frames": frames, "n_gpus": world, "scaling": "strong", "seconds": ms / 1e3, "frames_per_s": frames / (ms / 1e3),
            "value": mpix / (ms / 1e3), "unit": "Mpix/s", "launches_rank0": launches,
            "roofline_frac": 48.0 * (hi - lo) * hs * ws / (ms * 1e-3) / 1e9 / hbm_peak(),
            "checksums": [float(v) for v in sums.cpu()],
            "parity": f"checksum == frames/64 x the 64-frame pool's checksum: max relative difference {rel:.1e}",
            "parity_ok": bool(rel < 1e-12)}


# --------------------------------------------------------------------------------------------------
# cfg4: the train-loader path at training resolution feeding the encoders, batch 32
# --------------------------------------------------------------------------------------------------
def cfg4_loader(rank, world, dev, reps=200, with_encoders=True):
    """BASELINE configs[3]: four uint8 planes [32, 320, 480] (indoor_dataset.py:435-438) -> XOLP [32,2,320,480] +
    get_normals [32,9,320,480] in ONE launch (`polcue_fused_planes_u8`), alone and followed by the forward passes of
    the encoders they feed (pre_encoders.py:49-97, resnet_encoder.py:783-822; torch stand-ins with the reference's layer
    geometry, random weights, eval mode -- their time is measured, they are not part of the path)."""
    b, h, w = 32, synth.TRAIN_H, synth.TRAIN_W
    planes = [torch.stack([torch.from_numpy(synth.gen_p_planes(rank * b + i, h, w)[k]) for i in range(b)]).to(dev) for k in range(4)]
    ops.lut_for(1.5, dev)
    out = {}

    def kernel():
        nonlocal out
        out = ops.fused_planes(*planes, n=1.5, out=out)
        return out

    def timed(fn, n):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        D.barrier()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return D.max_over_ranks(a.elapsed_time(e) / n, dev)

    ms_kernel = timed(kernel, reps)
    px = b * h * w
    res = {"workload": "cfg4: train-loader path at training resolution, planes u8 [32,4,320,480] -> xolp [32,2,320,480] + normals "
                       "[32,9,320,480] (1 fused launch) -> ShallowEncoder / ShallowNormalsEncoder / ShallowResnetEncoder forward",
           "n_gpus": world, "scaling": "weak", "kernel_us": ms_kernel * 1e3, "value": world * b / (ms_kernel * 1e-3), "unit": "samples/s",
           "roofline_frac": 48.0 * px / (ms_kernel * 1e-3) / 1e9 / hbm_peak(),
           "note": "4.9 Mpx per launch: 236 MB, launch-latency sized"}
    if with_encoders:
        from encoder_standins import PolarEncoders
        enc = PolarEncoders().to(dev).eval()
        rgb = torch.rand((b, 3, h, w), device=dev)

        @torch.no_grad()
        def encoders_only():
            return enc(out["xolp"], out["normals"], rgb)

        @torch.no_grad()
        def both():
            o = kernel()
            return enc(o["xolp"], o["normals"], rgb)

        n = max(10, reps // 10)
        ms_enc = timed(encoders_only, n)
        ms_both = timed(both, n)
        res.update({"encoders_ms": ms_enc, "kernel_plus_encoders_ms": ms_both, "kernel_share_of_step": ms_kernel / ms_both,
                    "samples_per_s_with_encoders": world * b / (ms_both * 1e-3)})
    return res


# --------------------------------------------------------------------------------------------------
# cfg5: the evaluation path over a HAMMER-test-sized split
# --------------------------------------------------------------------------------------------------
def cfg5_eval(rank, world, dev, images=120, reps=200, check=False, graph=True):
    """BASELINE configs[4]: per rank its contiguous shard of `images` 320x480 images -> GT depth->normals stencil +
    per-image masked compute_depth_errors for the range mask and every material level (11 groups, one launch) ->
    accumulators of the mean over images; one sum over ranks of 1 + 11 x 7 float64.  `polcue_eval_pass_peer_f32`: three
    launches, the last of which also exchanges the accumulators over NVLink peer memory; no host work in between; timed
    eagerly, replayed from a CUDA graph (the 120-image pass is launch-latency sized), and with NCCL's all-reduce instead."""
    lo, hi = D.shard_range(images, rank, world)
    n_local = hi - lo
    groups = [None] + list(synth.MATERIAL_LEVELS)
    n_acc = 1 + 7 * len(groups)
    if n_local > 0:
        gt, pred, inst, k = (torch.from_numpy(a).to(dev) for a in synth.gen_depth_batch(lo, n_local))
    bufs = {}
    peer, peer_note = None, None
    if world > 1:
        try:
            peer = D.PeerExchange(dev)                         # NVLink peer-memory exchange (one node); raises on every rank or on none
        except RuntimeError as e:
            peer_note = str(e)
    zeros, zeros_out = torch.zeros(n_acc, dtype=torch.float64, device=dev), torch.zeros(n_acc, dtype=torch.float64, device=dev)

    # prepared launchers (ops.EvalPass): arguments checked and buffers bound once, a run only enqueues the three launches
    plain = ops.EvalPass(gt, pred, inst, k, 0.1, 2.0, groups) if n_local > 0 else None
    fused = ops.EvalPass(gt, pred, inst, k, 0.1, 2.0, groups, peer=peer) if (n_local > 0 and peer is not None) else None
    bufs = (fused or plain).out if n_local > 0 else {}

    def local_pass():
        return plain.run()["mean_acc"] if plain is not None else zeros

    def evaluate_nccl():
        acc = local_pass().clone()
        D.all_reduce_sums(acc)                                # library all-reduce of 78 float64: a launch + a protocol round trip
        return acc

    def evaluate():
        """The product path: the sum over ranks happens inside the pass's last kernel (polcue_eval_pass_peer_f32)."""
        if peer is None:
            return evaluate_nccl() if world > 1 else local_pass()
        if fused is not None:
            return fused.run()["mean_acc_all"]
        return peer.all_reduce(zeros, out=zeros_out)          # a rank without images still takes part

    def evaluate_unprepared():
        """the same through ops.eval_pass (argument checks and buffer look-ups on every call)"""
        nonlocal bufs
        if n_local > 0 and (peer is not None or world == 1):
            bufs = ops.eval_pass(gt, pred, inst, k, 0.1, 2.0, groups, out=bufs, peer=peer)
            return bufs["mean_acc_all" if peer is not None else "mean_acc"]
        return evaluate()

    def timed(fn, n):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        D.barrier()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return D.max_over_ranks(a.elapsed_time(e) / n, dev)

    acc = evaluate().clone()
    acc_nccl = evaluate_nccl()
    torch.cuda.synchronize()
    assert torch.allclose(acc, acc_nccl, rtol=1e-13, atol=0, equal_nan=True), (acc, acc_nccl)   # same sum, rank order vs NCCL's order
    means = (acc[1:] / acc[0]).reshape(len(groups), 7)
    ms_eager = timed(evaluate, reps)
    ms_unprepared = timed(evaluate_unprepared, reps)
    ms_nccl = timed(evaluate_nccl, reps) if world > 1 else None
    ms_local = timed(local_pass, reps)
    def replayed(fn):
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                fn()
        torch.cuda.current_stream(dev).wait_stream(side)
        return timed(g.replay, reps)

    ms_graph = ms_graph_local = us_exchange = None
    if graph and (world == 1 or peer is not None):
        ms_graph = replayed(evaluate)                          # the exchange is captured with the pass
        torch.cuda.synchronize()
        assert torch.equal(evaluate().clone(), acc)            # replays and eager calls give the same (rank-ordered) sums
        if peer is not None:
            ms_graph_local = replayed(local_pass)
            us_exchange = 1e3 * replayed(lambda: peer.all_reduce(zeros, out=zeros_out))   # the exchange alone: launch + one NVLink trip
    if peer is not None:
        made, failed = peer.status()
        assert failed == 0, f"peer exchange {failed} timed out"
    ms_best = min(ms_eager, ms_graph) if ms_graph is not None else ms_eager
    px = n_local * synth.TRAIN_H * synth.TRAIN_W
    bytes_alg = px * (16 + 9)                                  # stencil 4 R + 12 W, metrics 4 + 4 + 1 R (gt is read by both kernels)
    chk = (bufs["normals"].double().sum().reshape(1) if n_local > 0 else torch.zeros(1, dtype=torch.float64, device=dev))
    D.all_reduce_sums(chk)
    res = {"workload": f"cfg5: evaluation path, {images} synthetic 320x480 images over {world} GPU(s): GT depth->normals + per-image "
                       "masked depth errors for 11 mask groups + mean over images, summed over ranks inside the last kernel (polcue_eval_pass_peer_f32, 3 launches)",
           "images": images, "n_gpus": world, "scaling": "strong", "ms_per_pass": ms_best, "value": images / (ms_best * 1e-3),
           "ms_per_pass_is": "CUDA-graph replay of the whole pass, exchange included" if ms_graph is not None and ms_graph <= ms_eager else "eager launches through ops.EvalPass",
           "ms_per_pass_eager_prepared_launcher": ms_eager,
           "unit": "images/s", "collective": ("sum of %d float64 over NVLink peer memory inside the pass's last kernel (polcue_eval_pass_peer_f32), "
                                              "rank-ordered; checked against the NCCL all-reduce" % n_acc) if peer is not None else
                                             ("NCCL all-reduce (peer memory unavailable: %s)" % peer_note if world > 1 else "none (one rank)"),
           "ms_per_pass_cuda_graph_replay": ms_graph, "ms_per_pass_through_ops_eval_pass": ms_unprepared, "ms_per_pass_with_nccl_all_reduce": ms_nccl, "ms_local_launches_only": ms_local,
           "ms_local_cuda_graph_replay": ms_graph_local, "us_peer_exchange_alone_graph_replay": us_exchange,
           "roofline_frac_of_pass": bytes_alg / (ms_best * 1e-3) / 1e9 / hbm_peak(),
           "abs_rel_all": float(means[0, 0]), "a1_all": float(means[0, 4]), "normals_checksum": float(chk[0])}
    if check and rank == 0:
        from oracle import polcue_oracle as O                  # the checker (test infrastructure), never timed
        g_, p_, i_, kk = synth.gen_depth_batch(0, images)
        t0 = time.perf_counter()
        worst = 0.0
        for gi, level in enumerate(groups):
            _, mean = O.depth_errors_per_image(g_, p_, 0.1, 2.0, i_ if level is not None else None, level)
            got = means[gi].cpu().numpy()
            assert np.array_equal(np.isnan(got), np.isnan(mean)), (level, got, mean)
            assert np.allclose(got, mean, rtol=5e-6, equal_nan=True), (level, got, mean)
            worst = max(worst, float(np.nanmax(np.abs(got / mean - 1))))
        res["parity"] = (f"sharded mean over images == unsharded oracle for {len(groups)} mask groups, worst relative difference "
                         f"{worst:.1e} ({time.perf_counter() - t0:.1f} s on the CPU)")
        res["parity_ok"] = True
    if peer is not None:
        peer.close()
    return res
