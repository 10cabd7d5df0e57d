#!/usr/bin/env python
"""Per-kernel roofline readings for every CUDA entry point of libpolcue.so other than the headline fused kernel
(bench.py covers that one).  For each kernel: algorithmic bytes per launch / mean launch time (CUDA
events on the launching stream) against the measured HBM copy peak.  Working sets smaller than ~2x the L2 are
preceded by an L2 flush (a 512 MB read, so that no dirty lines are left for the kernel to write back) inside the loop;
the separately measured cost of the flushes is subtracted.

  python tools/bench_kernels.py [--reps 10] [--only stencil,metrics] > profiles/kernels_rNN.jsonl
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200"))
from polcue import _lib, ops, synth  # noqa: E402


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def timeit(fn, reps, flush):
    """Mean launch time with the GPU kept busy: `reps` calls are enqueued back to back between two events, so host
    launch overhead overlaps with the previous kernel (as it does in a real pipeline).  With `flush`, an L2-sized
    write precedes every call inside the loop and the separately measured cost of those writes is subtracted."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()

    def run(body):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            body()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    if flush is None:
        return run(fn), run(fn)

    # The flush READS 512 MB (a reduction), so L2 ends up full of CLEAN lines of the flush buffer.  A write flush would
    # leave ~126 MB of dirty lines whose write-back the measured kernel then pays for (about 20 us of DRAM traffic).
    view = flush.view(torch.int32)

    def both():
        view.sum()
        fn()

    t_flush = min(run(lambda: view.sum()) for _ in range(2))
    t_both = min(run(both) for _ in range(2))
    return t_both - t_flush, t_both - t_flush


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default="")
    ap.add_argument("--once", action="store_true", help="call every kernel exactly once (for ncu captures), no timing")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peak, peak_src = hbm_peak()
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    rng = np.random.default_rng(0)

    def emit(name, config, alg_bytes, fn, small=False):
        if only and name.split("/")[0] not in only:
            return
        if args.once:
            fn()
            torch.cuda.synchronize()
            return
        l0 = _lib.launch_count()
        fn()
        launches = _lib.launch_count() - l0
        mean_ms, min_ms = timeit(fn, args.reps, flush_buf if small else None)
        achieved = alg_bytes / (mean_ms * 1e-3) / 1e9
        print(json.dumps({"kernel": name, "config": config, "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": mean_ms,
                          "best_ms": min_ms, "achieved_gbs": achieved, "peak_gbs": peak, "peak_source": peak_src,
                          "frac": achieved / peak, "l2_flushed": bool(small), "kernels_per_call": launches}), flush=True)

    # ---- XOLP-only and normals-only members of the fused family (cfg2 geometry) ----
    B, hs, ws = 64, 1024, 1224
    px = B * hs * ws
    planes = [torch.from_numpy(rng.integers(0, 256, (B, hs, ws), dtype=np.uint8)).to(dev) for _ in range(4)]
    stack = torch.stack(planes, dim=3).contiguous()
    emit("xolp_stack_u8", f"stack u8 [{B},{hs},{ws},4] -> xolp f32", 12 * px, lambda: ops.xolp_from_stack(stack, None, want_iun=False))
    emit("xolp_planes_u8", f"4 planes u8 [{B},{hs},{ws}] -> xolp f32", 12 * px, lambda: ops.xolp_from_planes(*planes))
    _, xolp = ops.xolp_from_stack(stack, None, want_iun=False)
    del stack, planes
    emit("stats/channel_stats", f"xolp f32 [{B},2,{hs},{ws}] -> per-channel sum, sum of squares (float64)", 8 * px,
         lambda: ops.channel_stats(xolp))
    emit("normals_from_xolp", f"xolp f32 [{B},2,{hs},{ws}] -> normals f32 [{B},9,..]", 44 * px, lambda: ops.get_normals(xolp, 1.5))
    mosaic = torch.from_numpy(rng.integers(0, 256, (B, 2 * hs, 2 * ws), dtype=np.uint8)).to(dev)
    emit("split_pol", f"mosaic u8 [{B},{2 * hs},{2 * ws}] -> 4 quadrants", 2 * mosaic.numel(), lambda: ops.split_pol_batch(mosaic))
    emit("fused_xolp_only", f"mosaic u8 [{B},{2 * hs},{2 * ws}] -> xolp f32 (no normals)", 12 * px,
         lambda: ops.fused_mosaic(mosaic, 1.5, want_normals=False))
    del xolp, mosaic
    torch.cuda.empty_cache()

    # ---- depth -> normals stencil and metrics (cfg5 geometry and a 16x scaled split) ----
    for n_img in (120, 1920):
        h, w = 320, 480
        base = synth.gen_depth_batch(0, 8, h, w)
        reps_ = n_img // 8
        gt = torch.from_numpy(base[0]).to(dev).repeat(reps_, 1, 1)
        pred = torch.from_numpy(base[1]).to(dev).repeat(reps_, 1, 1)
        inst = torch.from_numpy(base[2]).to(dev).repeat(reps_, 1, 1)
        k = torch.from_numpy(base[3]).to(dev).repeat(reps_, 1, 1)
        npx = gt.numel()
        small = npx * 16 < 300e6
        depth = gt[:, None].contiguous()
        emit("stencil/depth_to_normals", f"depth f32 [{n_img},1,{h},{w}] -> normals f32 [{n_img},3,{h},{w}]", 16 * npx,
             lambda: ops.depth_to_normals(depth, k), small)
        emit("metrics/per_image", f"gt,pred f32 [{n_img},{h},{w}], range mask", 8 * npx,
             lambda: ops.depth_errors_per_image(gt, pred, 0.1, 2.0), small)
        emit("metrics/per_image_inst", f"gt,pred f32 + inst u8 [{n_img},{h},{w}], material filter", 9 * npx,
             lambda: ops.depth_errors_per_image(gt, pred, 0.1, 2.0, inst, 40), small)
        emit("metrics/groups11", f"gt,pred f32 + inst u8 [{n_img},{h},{w}], 11 mask groups in one launch (bytes counted once)", 9 * npx,
             lambda: ops.depth_errors_groups(gt, pred, inst, 0.1, 2.0, [None] + list(synth.MATERIAL_LEVELS)), small)
        smooth = torch.where(gt > 0, gt, torch.full_like(gt, 0.7))[:, None].contiguous()
        predd = pred[:, None].contiguous().requires_grad_(True)
        maskf = ((gt >= 0.1) & (gt <= 2.0)).float()[:, None].contiguous()
        if n_img * ((h + 15) // 16) * ((w + 127) // 128) <= (1 << 18):
            # straight through the C ABI with preallocated buffers (the autograd wrapper costs more host time than the kernels)
            import ctypes as C
            L = _lib.lib()
            wsb = torch.zeros(int(L.polcue_normals_loss_workspace_bytes()), dtype=torch.uint8, device=dev)
            sums2 = torch.empty(2, dtype=torch.float64, device=dev)
            lossv, one = torch.empty((), device=dev), torch.ones((), device=dev)
            gradp = torch.empty_like(predd)
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            pd = predd.detach()

            def fwd():
                assert L.polcue_normals_loss_fwd_f32(smooth.data_ptr(), pd.data_ptr(), k.data_ptr(), maskf.data_ptr(), n_img, h, w,
                                                     wsb.data_ptr(), sums2.data_ptr(), lossv.data_ptr(), st) == 0

            def bwd():
                assert L.polcue_normals_loss_bwd_f32(smooth.data_ptr(), pd.data_ptr(), k.data_ptr(), maskf.data_ptr(), n_img, h, w,
                                                     sums2.data_ptr(), one.data_ptr(), gradp.data_ptr(), st) == 0
            emit("loss/normals_fwd", f"depth_gt, depth_pred, mask f32 [{n_img},1,{h},{w}] -> loss (C ABI)", 12 * npx, fwd, small)
            emit("loss/normals_bwd", f"-> d loss / d depth_pred f32 [{n_img},1,{h},{w}] (C ABI)", 16 * npx, bwd, small)

            # the trainer's own call (trainer.py:1240-1251): range mask derived from the GT tile, L1 depth loss folded in; no mask tensor
            sums3 = torch.empty(3, dtype=torch.float64, device=dev)
            losses2 = torch.empty(2, device=dev)
            gtc = gt[:, None].contiguous()

            def sup_fwd():
                assert L.polcue_supervised_losses_fwd_f32(gtc.data_ptr(), pd.data_ptr(), k.data_ptr(), C.c_float(0.1), C.c_float(2.0), n_img, h, w,
                                                          wsb.data_ptr(), sums3.data_ptr(), losses2.data_ptr(), st) == 0

            def sup_bwd():
                assert L.polcue_supervised_losses_bwd_f32(gtc.data_ptr(), pd.data_ptr(), k.data_ptr(), C.c_float(0.1), C.c_float(2.0), n_img, h, w,
                                                          sums3.data_ptr(), one.data_ptr(), one.data_ptr(), gradp.data_ptr(), st) == 0
            emit("loss/supervised_fwd", f"depth_gt (with zero-depth holes), depth_pred f32 [{n_img},1,{h},{w}] -> normals loss + L1 depth loss (C ABI)",
                 8 * npx, sup_fwd, small)
            emit("loss/supervised_bwd", f"-> d (both losses) / d depth_pred f32 [{n_img},1,{h},{w}] (C ABI)", 12 * npx, sup_bwd, small)

            def fwd_bwd():
                predd.grad = None
                ops.normals_loss(smooth, predd, k, maskf).backward()
            emit("loss/normals_autograd", "ops.normals_loss(...).backward(): forward + backward through torch.autograd (12 + 16 B/px)",
                 28 * npx, fwd_bwd, small)
        m = gt > 0
        gflat, pflat = gt[m].contiguous(), pred[m].contiguous()
        emit("metrics/flat", f"gt,pred f32 [{gflat.numel()}] (compacted)", 8 * gflat.numel(),
             lambda: ops.depth_error_sums(gflat, pflat), small)
        del gt, pred, inst, depth, gflat, pflat
        torch.cuda.empty_cache()

    # ---- loader front end: Pillow-exact Lanczos resize of the four polarizer images + XOLP + normals (cfg4 geometry) ----
    B, ih, iw, oh, ow = 32, 832, 1088, 320, 480            # HAMMER quadrants -> training resolution
    full = [torch.from_numpy(np.stack([synth.gen_p_planes(f % 4, ih, iw)[k] for f in range(4)] * (B // 4))).to(dev) for k in range(4)]
    stacked = torch.stack(full, dim=1).contiguous()
    ws = torch.empty((4 * B * ih + 32) * ow, dtype=torch.uint8, device=dev)
    in_bytes, opx = 4 * B * ih * iw, B * oh * ow
    emit("loader/lanczos_resize", f"u8 [{B},4,{ih},{iw}] -> u8 [{B},4,{oh},{ow}] (2 launches; integer-MAC bound, bytes = in + out)",
         in_bytes + 4 * opx, lambda: ops.lanczos_resize(stacked, (oh, ow), workspace=ws), True)
    if not args.once and (not only or "loader" in only):
        import ctypes as C
        L = _lib.lib()
        L.polcue_debug_resize_pass_times(1, None, None)
        th, tv = [], []
        for _ in range(8):
            flush_buf.fill_(1)
            ops.lanczos_resize(stacked, (oh, ow), workspace=ws)
            a, b = C.c_float(), C.c_float()
            L.polcue_debug_resize_pass_times(1, C.byref(a), C.byref(b))
            th.append(a.value)
            tv.append(b.value)
        L.polcue_debug_resize_pass_times(0, None, None)
        print(json.dumps({"kernel": "loader/lanczos_resize passes", "horizontal_ms": sorted(th)[len(th) // 2], "vertical_ms": sorted(tv)[len(tv) // 2],
                          "note": "CUDA events between the two launches of one call, L2 flushed before the call"}), flush=True)
    keep = {}
    keep.update(ops.loader_front_end(*full, (oh, ow)))
    emit("loader/front_end", f"4 x u8 [{B},{ih},{iw}] -> planes u8 + xolp f32 + normals f32 at {oh}x{ow} (3 launches)",
         in_bytes + (4 + 44) * opx, lambda: ops.loader_front_end(*full, (oh, ow), out=keep), True)
    small = keep["planes"]
    p4 = [small[:, k].contiguous() for k in range(4)]
    keep2 = {}
    keep2.update(ops.fused_planes(*p4))
    emit("loader/fused_planes", f"4 x u8 [{B},{oh},{ow}] -> xolp + normals (the front end's last launch alone)", 48 * opx,
         lambda: ops.fused_planes(*p4, out=keep2), True)


if __name__ == "__main__":
    main()
