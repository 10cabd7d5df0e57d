#!/usr/bin/env python
"""Stress run for the resize kernels (mbarrier/TMA pipeline, cp.async double buffer): many random geometries against
the CPU restatement, and repeated runs of one large batch that must be bit-identical.  Not part of the test suite.
  python tools/stress_resize.py [--cases 300] [--repeats 200]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200"))
sys.path.insert(0, ROOT)
from oracle import polcue_oracle as O  # noqa: E402
from polcue import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=300)
    ap.add_argument("--repeats", type=int, default=200)
    args = ap.parse_args()
    rng = np.random.default_rng(2027)
    bad = 0
    for case in range(args.cases):
        ih, iw = int(rng.integers(1, 300)), int(rng.integers(1, 400))
        if case % 3 == 0:
            iw = 16 * max(1, iw // 16)
        oh, ow = int(rng.integers(1, 120)), int(rng.integers(1, 200))
        if case % 2 == 0:
            ow = 4 * max(1, ow // 4)
        n = int(rng.integers(1, 10))
        imgs = rng.integers(0, 256, (n, ih, iw), dtype=np.uint8)
        flips = [bool(v) for v in rng.integers(0, 2, n)]
        got = ops.lanczos_resize(torch.from_numpy(imgs).cuda(), (oh, ow), flip=flips).cpu().numpy()
        for k in range(n):
            src = np.ascontiguousarray(imgs[k][:, ::-1]) if flips[k] else imgs[k]
            if not np.array_equal(got[k], O.resize_lanczos_u8(src, (oh, ow))):
                bad += 1
                print("MISMATCH", case, (ih, iw, oh, ow), k, flips[k], flush=True)
    print(f"{args.cases} random geometries: {bad} mismatching images")
    x = torch.randint(0, 256, (32, 4, 832, 1088), dtype=torch.uint8, device="cuda")
    ref = ops.lanczos_resize(x, (320, 480), flip=[bool(i & 1) for i in range(128)])
    diff = 0
    for _ in range(args.repeats):
        diff += int(not torch.equal(ops.lanczos_resize(x, (320, 480), flip=[bool(i & 1) for i in range(128)]), ref))
    print(f"{args.repeats} repeats of 128 images 832x1088 -> 320x480: {diff} differing runs")
    sys.exit(1 if (bad or diff) else 0)


if __name__ == "__main__":
    main()
