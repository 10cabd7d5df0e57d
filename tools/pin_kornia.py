#!/usr/bin/env python
"""Pins the depth->normals oracle on kornia itself -- run this wherever kornia 0.5.11 IS importable.

The reference calls `kornia.geometry.depth.depth_to_normals` (manydepth/trainer.py:37,1305-1306,1477,1484; pinned
kornia==0.5.11 at environment.yml:41).  kornia is neither vendored in the reference nor installable in the authoring
container (no network), so `oracle.depth_to_normals` restates its published algorithm and DESIGN.md says "parity
unpinned" for that one function.  This script closes the gap on any machine that has kornia:

    pip install kornia==0.5.11        # any machine with network access; CPU is enough
    python tools/pin_kornia.py        # writes tests/golden/kornia_outputs.npz

It evaluates kornia (float64 and float32, CPU) on the seeded depth maps the tests use -- hole-free surfaces, GT with
10 % random invalid pixels, GT with missing regions / invalid borders / an isolated valid pixel, random cameras -- and
stores inputs + outputs.  `tests/test_oracle_golden.py::test_depth_to_normals_oracle_against_kornia_outputs` consumes the
file when present (skips otherwise): the float64 oracle must equal kornia's float64 result to 1e-12, and
`tests/test_gpu_parity.py::test_depth_to_normals_against_kornia_outputs` holds the CUDA kernel to 1e-3 rad against it.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200"))
from polcue import synth  # noqa: E402


def cases():
    """(name, depth [B,1,H,W] float32, K [B,3,3] float32) -- the inputs of the stencil tests."""
    out = []
    for h, w in ((320, 480), (64, 96), (37, 131), (5, 3), (1, 1), (9, 260)):
        gt, _, _, k = synth.gen_depth_batch(3, 3, h, w)
        holes = synth.add_hole_regions(gt, 3) if min(h, w) >= 8 else gt
        smooth = np.where(gt > 0, gt, 0.7).astype(np.float32)
        for tag, d in (("smooth", smooth), ("random_holes", gt), ("region_holes", holes)):
            out.append((f"{tag}_{h}x{w}", d[:, None].astype(np.float32), k.astype(np.float32)))
    rng = np.random.default_rng(99)
    for case in range(4):                                          # random cameras and tilted surfaces
        h, w = int(rng.integers(2, 70)), int(rng.integers(2, 200))
        v, u = np.mgrid[0:h, 0:w].astype(np.float64)
        d = (0.3 + rng.uniform(0.2, 1.5) * (1 + 0.3 * np.sin(u / rng.uniform(5, 60)) * np.cos(v / rng.uniform(5, 60)))
             + rng.uniform(-2e-3, 2e-3) * u).astype(np.float32)[None, None]
        k = np.array([[[rng.uniform(200, 900), 0, rng.uniform(0, w)], [0, rng.uniform(200, 900), rng.uniform(0, h)], [0, 0, 1]]], np.float32)
        out.append((f"camera_{case}", d, k))
    return out


def main():
    import kornia
    import torch
    from kornia.geometry.depth import depth_to_normals
    store = {"kornia_version": np.array(kornia.__version__), "torch_version": np.array(torch.__version__)}
    for name, depth, k in cases():
        store[name + "/depth"] = depth
        store[name + "/K"] = k
        store[name + "/normals_f64"] = depth_to_normals(torch.from_numpy(depth).double(), torch.from_numpy(k).double()).numpy()
        store[name + "/normals_f32"] = depth_to_normals(torch.from_numpy(depth), torch.from_numpy(k)).numpy()
    path = os.path.join(ROOT, "tests", "golden", "kornia_outputs.npz")
    np.savez_compressed(path, **store)
    print(f"wrote {path}: {len(store) // 4} cases with kornia {kornia.__version__}")


if __name__ == "__main__":
    main()
