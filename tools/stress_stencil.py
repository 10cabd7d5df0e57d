#!/usr/bin/env python
"""Randomized differential stress of the depth->normals stencil and the normals loss on ground truth WITH zero-depth holes.

  python tools/stress_stencil.py [--cases 120] [--seed 7]

Per case: random batch / shape (widths that are and are not multiples of 4), random camera, a smooth surface with a step,
a random fraction (0-40 %) of invalid pixels plus random missing regions; then
  * the stencil under the every-pixel conditioning protocol of tests/parity.py (1e-3 rad wherever float32 can reach it,
    exact zero vectors inside holes, counted round-off pixels elsewhere), TMA-staged and hand-staged kernels bit-equal;
  * loss, sums and gradient of the packed kernels bit-identical to the scalar kernels (rows that are 16-byte multiples);
  * loss and gradient against the float64 autograd oracle (slack of 2 / sum(mask) per counted round-off normal, gradient
    compared outside their 3 x 3 neighbourhoods).
The oracle is test infrastructure: this tool is a checker, never the thing measured.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import parity as P  # noqa: E402
from oracle import polcue_oracle as O  # noqa: E402
from polcue import ops, synth  # noqa: E402


def misaligned(t):
    flat = torch.empty(t.numel() + 1, dtype=t.dtype, device=t.device)
    view = flat[1:].view(t.shape)
    view.copy_(t)
    return view


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=120)
    ap.add_argument("--seed", type=int, default=7)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    worst = {"stencil": 0.0, "ill_fraction": 0.0, "loss": 0.0, "grad": 0.0}
    for case in range(args.cases):
        b, h = int(rng.integers(1, 4)), int(rng.integers(2, 90))
        w = int(rng.integers(2, 300))
        if case % 2 == 0:
            w = 4 * max(1, w // 4)
        v, u = np.mgrid[0:h, 0:w].astype(np.float64)
        depth = np.stack([0.3 + rng.uniform(0.2, 1.5) * (1 + 0.3 * np.sin(u / rng.uniform(5, 60) + rng.uniform(0, 6)) *
                                                          np.cos(v / rng.uniform(5, 60))) + rng.uniform(-2e-3, 2e-3) * u
                          + 0.2 * (u > w * rng.uniform(0.2, 0.8)) for _ in range(b)]).astype(np.float32)
        gt = np.where(rng.random(depth.shape) < rng.uniform(0.0, 0.4), 0.0, depth).astype(np.float32)
        if min(h, w) >= 8 and case % 3:
            gt = synth.add_hole_regions(gt, case)
        k = np.stack([np.array([[rng.uniform(200, 900), 0, rng.uniform(0, w)], [0, rng.uniform(200, 900), rng.uniform(0, h)], [0, 0, 1]])
                      for _ in range(b)]).astype(np.float32)
        g, kk = dev(gt)[:, None], dev(k)
        n1 = ops.depth_to_normals(g, kk)
        assert torch.equal(n1, ops.depth_to_normals(misaligned(g), kk)), case
        err, ill = P.assert_stencil_normals_close(n1.cpu().numpy(), gt[:, None], k, O, what=f"case {case}", max_ill_fraction=1.0)
        worst["stencil"] = max(worst["stencil"], err)
        worst["ill_fraction"] = max(worst["ill_fraction"], ill / gt.size)
        # loss: smooth prediction, random mask
        pred = (np.where(gt > 0, gt, 0.7) * (1 + 0.05 * np.sin(u / 23.0) * np.cos(v / 17.0))).astype(np.float32)
        mask = (rng.random(gt.shape) < 0.8).astype(np.float32)
        if mask.sum() == 0:
            continue
        pr, m = dev(pred)[:, None], dev(mask)[:, None]
        d1 = pr.clone().requires_grad_(True)
        loss = ops.normals_loss(g, d1, kk, m)
        loss.backward()
        if w % 4 == 0:
            d2 = pr.clone().requires_grad_(True)
            loss2 = ops.normals_loss(misaligned(g), d2, kk, m)
            loss2.backward()
            assert torch.equal(loss.detach(), loss2.detach()) and torch.equal(d1.grad, d2.grad), f"packed != scalar in case {case}"
        _, bound, dead = O.depth_to_normals_conditioning(gt[:, None], k)
        _, bound_p, _ = O.depth_to_normals_conditioning(pred[:, None], k)
        bad = ((P.STENCIL_ROUNDINGS * 2.0 ** -24 * np.maximum(bound, bound_p) > P.NORMAL_TOL) & ~dead)
        near = bad.copy()
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                near |= np.roll(np.roll(bad, dy, axis=1), dx, axis=2)
        t64 = lambda a: torch.from_numpy(a.astype(np.float64))
        p64 = t64(pred)[:, None].requires_grad_(True)
        ref = O.normals_loss_torch(t64(gt)[:, None], p64, t64(k), t64(mask)[:, None])
        ref.backward()
        slack = 2.0 * float((bad & (mask > 0)).sum()) / float(mask.sum())
        dl = abs(float(loss.detach()) - float(ref.detach()))
        assert dl <= 3e-5 * abs(float(ref.detach())) + slack, (case, dl, slack)
        worst["loss"] = max(worst["loss"], max(0.0, dl - slack))
        if min(h, w) >= 3 and (~near).any():
            rg = p64.grad[:, 0].numpy()
            gg = d1.grad[:, 0].cpu().numpy().astype(np.float64)
            scale = np.abs(rg).max() + 1e-30
            e = (np.abs(gg - rg) / scale)[~near].max()
            assert e < 4e-3, (case, e)
            worst["grad"] = max(worst["grad"], float(e))
    print(f"{args.cases} cases ok; worst: {worst}")


if __name__ == "__main__":
    main()
