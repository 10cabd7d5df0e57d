#!/usr/bin/env python
"""One-off differential stress of the headline kernels against the CPU oracle: many random (shape, generator, refractive
index, layout) cases through polcue_fused_mosaic_u8 / _planes_u8 / _superpixel_u8 and polcue_normals_from_xolp_f32.
Not part of the test suite.   python tools/stress_fused.py [--cases 200]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity as P  # noqa: E402
from oracle import polcue_oracle as O  # noqa: E402
from polcue import _lib, ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=200)
    args = ap.parse_args()
    rng = np.random.default_rng(4096)
    worst = {"rho": 0.0, "phi": 0.0, "normals": 0.0}
    for case in range(args.cases):
        hs, ws = int(rng.integers(1, 160)), int(rng.integers(1, 260))
        if case % 3 == 0:
            ws = 4 * max(1, ws // 4)
        b = int(rng.integers(1, 4))
        n = float(rng.uniform(1.02, 3.0)) if case % 4 else 1.5
        kind = case % 3
        if kind == 0:
            mosaics = rng.integers(0, 256, (b, 2 * hs, 2 * ws), dtype=np.uint8)
        elif kind == 1:
            mosaics = np.stack([synth.tile_mosaic(synth.gen_p_planes(case * 7 + f, hs, ws)) for f in range(b)])
        else:       # few grey levels: many ties and degenerate pixels
            mosaics = (rng.integers(0, 4, (b, 2 * hs, 2 * ws)) * 85).astype(np.uint8)
        with ops.trig("mufu" if case & 1 else "poly"):
            out = ops.fused_mosaic(torch.from_numpy(mosaics).cuda(), n, want_iun=True, want_planes=True)
            for f in range(b):
                stack = O.stack_quadrants(mosaics[f])
                assert np.array_equal(out["planes"][f].cpu().numpy(), stack.transpose(2, 0, 1)), case
                iun, rho, phi = O.iun_and_xolp_closed(stack)
                g_rho, g_phi = out["xolp"][f, 0].cpu().numpy(), out["xolp"][f, 1].cpu().numpy()
                P.assert_dolp_close(out["iun"][f].cpu().numpy(), iun, "Iun")
                P.assert_dolp_close(g_rho, rho, "rho")
                P.assert_aolp_close(g_phi, phi)
                ref = O.get_normals(np.stack((rho, phi))[None], n).reshape(3, 3, hs, ws)
                got = out["normals"][f].cpu().numpy().reshape(3, 3, hs, ws)
                worst["normals"] = max(worst["normals"], P.assert_normals_close(got, ref, axis=1, what=f"case {case} n={n:.4f}"))
                worst["rho"] = max(worst["rho"], float(np.max(np.abs(g_rho - rho) / (np.abs(rho) + 1e-2))))
                worst["phi"] = max(worst["phi"], float(np.max(P.aolp_error(g_phi, phi))))
            # the other entry points give the same bits
            planes = [out["planes"][:, k].contiguous() for k in range(4)]
            fp = ops.fused_planes(*planes, n=n)
            gn = ops.get_normals(out["xolp"], n)
            raw = torch.zeros_like(torch.from_numpy(mosaics)).cuda()
            pattern = [int(v) for v in rng.permutation(4)]
            for pos, angle in enumerate(pattern):
                raw[:, pos // 2::2, pos % 2::2] = planes[angle]
            sp = ops.fused_mosaic(raw, n, superpixel=pattern)
            assert torch.equal(fp["normals"], out["normals"]) and torch.equal(sp["normals"], out["normals"]) and torch.equal(sp["xolp"], out["xolp"]), case
            ref_x = O.get_normals(out["xolp"].cpu().numpy(), n)          # get_normals sees the float32 XOLP, as in the reference
            P.assert_normals_close(gn.cpu().numpy().reshape(b, 3, 3, hs, ws), ref_x.reshape(b, 3, 3, hs, ws), axis=2, what=f"get_normals case {case}")
    print(f"{args.cases} cases ok; worst errors: {worst}")


if __name__ == "__main__":
    main()
