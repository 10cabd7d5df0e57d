#!/usr/bin/env python
"""One tiny call of every CUDA entry point, for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_all.py`."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200"))
from polcue import ops, synth  # noqa: E402

dev = torch.device("cuda", 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
for (h, w) in ((64, 96), (34, 66), (10, 14)):                       # VEC = 4, 1 (odd half width), 1
    m = t(synth.gen_batch("U", 0, 3, h, w))
    out = ops.fused_mosaic(m, 1.5, want_iun=True, want_planes=True)
    ops.fused_mosaic(m, 1.5, want_normals=False)
    ops.get_normals(out["xolp"], 1.5)
    ops.split_pol_batch(m)
    ops.channel_stats(out["normals"])
    ops.xolp_from_planes(*[out["planes"][:, k].contiguous() for k in range(4)], want_iun=True)
    ops.xolp_from_stack(out["planes"].permute(0, 2, 3, 1).contiguous(), None)
    ops.xolp_from_stack(out["planes"].permute(0, 2, 3, 1).contiguous().float(), np.array([5., 50., 95., 140.]) * np.pi / 180)
    rho, phi = out["xolp"][:, 0].contiguous(), out["xolp"][:, 1].contiguous()
    ops.calc_normals(phi, ops.rho_diffuse(rho, 1.5))
    ops.rho_spec(rho, 1.5)
    st = out["planes"][0].permute(1, 2, 0).contiguous().float()
    r, p, _ = ops.stokes_channel(st, torch.ones(st.shape[:2], dtype=torch.uint8, device=dev))
    ops.calc_normals_channel(p, torch.nan_to_num(r), torch.ones_like(p, dtype=torch.uint8))
host = ops.fused_mosaic_host(torch.from_numpy(synth.gen_batch("P", 0, 3, 64, 96)).pin_memory(), 1.5, chunk_frames=2)
for (h, w) in ((64, 96), (37, 131), (5, 3)):                        # TMA path, manual path, tiny
    gt, pred, inst, k = (t(a) for a in synth.gen_depth_batch(0, 3, h, w))
    ops.depth_to_normals(gt[:, None], k)
    ops.depth_errors_per_image(gt, pred, 0.1, 2.0, inst, 40)
    ops.depth_errors_groups(gt, pred, inst, 0.1, 2.0, [None, 20, 40, 200])
    mk = gt > 0
    ops.depth_error_sums(gt[mk], pred[mk])
    dp = pred[:, None].clone().requires_grad_(True)
    ops.normals_loss(torch.where(gt > 0, gt, torch.full_like(gt, 0.7))[:, None], dp, k, (gt > 0).float()[:, None]).backward()
rng = np.random.default_rng(0)
for (ih, iw, oh, ow) in ((96, 128, 40, 60), (61, 83, 23, 31), (200, 16, 7, 16), (12, 10, 30, 41), (256, 300, 8, 9)):
    # TMA-pipelined + cp.async passes, manual staging, identity axis, upscaling, generic tap counts
    full = [t(rng.integers(0, 256, (3, ih, iw), dtype=np.uint8)) for _ in range(4)]
    ops.loader_front_end(*full, (oh, ow), flip=[False, True, False], want_iun=True, normalize_xolp=ops.XOLP_MEAN_STD)
    ops.lanczos_resize(full[0], (oh, ow), flip=True)
m = t(synth.gen_batch("U", 0, 2, 60, 144))
ops.fused_mosaic(m, 1.8)                                            # steep end segment: float64 out-of-line path
ops.fused_mosaic(m, 1.5, superpixel=(2, 1, 3, 0), want_planes=True)
torch.cuda.synchronize()
print("sanitize_all: every entry point ran")
