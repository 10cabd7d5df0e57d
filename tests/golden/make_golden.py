"""Generate tests/golden/*.npz by executing the REFERENCE functions in the authoring container.

Run once, here (the GPU box has no /root/reference):   python tests/golden/make_golden.py
The reference has no golden vectors or tests of its own, so these outputs of the
reference itself are what pins the oracle (oracle/polcue_oracle.py) and, through it, the kernels.

Reference entry points executed (paths relative to /root/reference):
  polarisation/pol_split_and_save.py:split_pol            polarisation/xolp.py:Iun_and_xolp
  polarisation/xolp_and_normals.py:{rho_diffuse,rho_spec,calc_normals}       (stub matplotlib)
  ppp_code/physical_normals_channels.py:{PolarisationImage_channel,rho_*_channel,calc_normals_channel}
  manydepth/normals_vec.py:{rho_diffuse,rho_spec,calc_normals}
  manydepth/networks/pre_encoders.py:ShallowNormalsEncoder.get_normals
  manydepth/layers.py:{compute_depth_errors,compute_depth_errors_numpy}
kornia.depth_to_normals is NOT available (parity unpinned, see oracle docstring).
"""
import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("POLCUE_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200"))
sys.path.insert(0, REF)
sys.path.insert(1, os.path.join(REF, "polarisation"))   # xolp_and_normals.py:10 imports its sibling by bare name

# matplotlib is absent; the two numpy scripts import it but their functions never call it
for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.image"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].use = lambda *a, **k: None
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["matplotlib"].image = sys.modules["matplotlib.image"]

from polarisation.pol_split_and_save import split_pol            # noqa: E402
from polarisation.xolp import Iun_and_xolp                       # noqa: E402
from polarisation import xolp_and_normals as ref_np              # noqa: E402
from ppp_code import physical_normals_channels as ref_ppp        # noqa: E402
from manydepth import normals_vec as ref_vec                     # noqa: E402
from manydepth.networks.pre_encoders import ShallowNormalsEncoder  # noqa: E402
from manydepth.layers import compute_depth_errors, compute_depth_errors_numpy  # noqa: E402

from polcue import synth                                         # noqa: E402

ANGLES = np.array([0, 45, 90, 135]) * np.pi / 180


def quiet(fn, *a):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a)


def main():
    out = {}

    # ---- known-answer pixels ------------------------------------------------
    kat = np.array([[0, 0, 0, 0], [10, 20, 30, 40], [200, 100, 0, 100], [255, 0, 0, 0], [0, 255, 0, 0],
                    [0, 0, 255, 0], [0, 0, 0, 255], [100, 100, 0, 0], [255, 255, 255, 255],
                    [1, 0, 0, 0], [0, 0, 1, 1], [7, 200, 9, 13], [128, 64, 128, 63]], dtype=np.uint8)
    iun, rho, phi = Iun_and_xolp(kat.reshape(1, -1, 4), ANGLES)
    out["kat_in"], out["kat_iun"], out["kat_rho"], out["kat_phi"] = kat, iun[0], rho[0], phi[0]

    # ---- split_pol ---------------------------------------------------------------------
    rng = np.random.default_rng(7)
    for tag, shape in (("gray", (8, 12)), ("bgr", (6, 10, 3))):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        q = quiet(split_pol, img)
        out[f"split_{tag}_in"] = img
        for name, arr in zip(("im00", "im10", "im01", "im11"), q):
            out[f"split_{tag}_{name}"] = np.ascontiguousarray(arr)

    # ---- XOLP + get_normals on seeded stacks -------------------------------------------
    stacks = {
        "u": np.random.default_rng(20260101).integers(0, 256, (64, 96, 4), dtype=np.uint8),
        "p": np.stack(synth.gen_p_planes(3, 64, 96), axis=2),
    }
    for tag, st in stacks.items():
        iun, rho, phi = Iun_and_xolp(st, ANGLES)
        out[f"xolp_{tag}_in"] = st
        out[f"xolp_{tag}_iun"], out[f"xolp_{tag}_rho"], out[f"xolp_{tag}_phi"] = iun, rho, phi
        x32 = torch.from_numpy(np.stack((rho, phi))[None]).float()        # trainer.py:510 `.float()`
        out[f"getn_{tag}_x32"] = x32.numpy()
        out[f"getn_{tag}_n"] = ShallowNormalsEncoder.get_normals(x32, 1.5).numpy()
        # the all-numpy float64 chain of polarisation/xolp_and_normals.py:116-120
        th_d = ref_np.rho_diffuse(rho, 1.5)
        th_1, th_2 = ref_np.rho_spec(rho, 1.5)
        out[f"chain_{tag}_nd"] = ref_np.calc_normals(phi, th_d)
        out[f"chain_{tag}_n1"] = ref_np.calc_normals(phi + np.pi / 2, th_1)
        out[f"chain_{tag}_n2"] = ref_np.calc_normals(phi + np.pi / 2, th_2)

    # non-canonical polarizer angles
    ang2 = np.array([5.0, 50.0, 95.0, 140.0]) * np.pi / 180
    iun, rho, phi = Iun_and_xolp(stacks["p"], ang2)
    out["xolp_ang2"], out["xolp_ang2_iun"], out["xolp_ang2_rho"], out["xolp_ang2_phi"] = ang2, iun, rho, phi

    # float-valued stack (any real dtype is accepted by the reference)
    fl = stacks["p"].astype(np.float32) * np.float32(0.37) + np.float32(1.5)
    iun, rho, phi = Iun_and_xolp(fl, ANGLES)
    out["xolp_f32_in"], out["xolp_f32_iun"], out["xolp_f32_rho"], out["xolp_f32_phi"] = fl, iun, rho, phi

    # ---- table inversions ---------------------------------------------------------------
    rq = np.concatenate(([0.0, 1e-6, 1e-4, 0.01, 0.1, 0.3, 0.3846, 0.384615384615, 0.5, 0.9, 0.99, 0.9999, 1.0, 1.2, 2.0],
                         np.random.default_rng(11).uniform(0, 1, 200) ** 2,
                         np.random.default_rng(12).uniform(0, 2, 50)))
    out["tab_rho"] = rq
    for n in (1.3, 1.5, 1.8):
        t = torch.from_numpy(rq)[None, None]
        out[f"tab_d_{n}"] = ref_vec.rho_diffuse(t, n).numpy().ravel()
        t1, t2 = ref_vec.rho_spec(t, n)
        out[f"tab_s1_{n}"], out[f"tab_s2_{n}"] = t1.numpy().ravel(), t2.numpy().ravel()
        assert np.array_equal(out[f"tab_d_{n}"], ref_np.rho_diffuse(rq, n))
        assert np.array_equal(out[f"tab_s1_{n}"], ref_ppp.rho_spec_channel(rq, n)[0])

    # ---- ppp channel variant -------------------------------------------------------------
    st = stacks["p"]
    mask = np.zeros(st.shape[:2], dtype=bool)
    mask[8:50, 10:80] = True
    mask[20:24, 30:40] = False
    images = np.zeros(st.shape, dtype=np.float64)
    for k in range(4):
        images[:, :, k][mask] = st[:, :, k][mask]                 # physical_normals_channels.py:119-123
    rho2, phi2, iun2 = ref_ppp.PolarisationImage_channel(images, ANGLES, mask)
    out["ppp_images"], out["ppp_mask"] = images, mask
    out["ppp_rho"], out["ppp_phi"], out["ppp_iun"] = rho2, phi2, iun2
    th_d = ref_ppp.rho_diffuse_channel(rho2, 1.5)
    th_1, th_2 = ref_ppp.rho_spec_channel(rho2, 1.5)
    out["ppp_nd"] = ref_ppp.calc_normals_channel(phi2, th_d, mask)
    out["ppp_n1"] = ref_ppp.calc_normals_channel(phi2 + np.pi / 2, th_1, mask)
    out["ppp_n2"] = ref_ppp.calc_normals_channel(phi2 + np.pi / 2, th_2, mask)

    # ---- depth metrics ---------------------------------------------------------------------
    g = torch.Generator().manual_seed(1234)
    gt = torch.rand(4096, generator=g) * 1.9 + 0.1
    pr = (gt * (1 + 0.2 * (torch.rand(4096, generator=g) - 0.5))).clamp(0.1, 2)
    out["met_gt"], out["met_pred"] = gt.numpy(), pr.numpy()
    out["met_torch"] = np.array([float(v) for v in compute_depth_errors(gt, pr)])
    out["met_numpy"] = np.array([float(v) for v in compute_depth_errors_numpy(gt.numpy(), pr.numpy())])
    gtb, prb, instb, _ = synth.gen_depth_batch(0, 3, 64, 96)
    rows = []
    for b in range(3):                                            # trainer.py:1376-1422, object == "all"
        m = np.logical_and(gtb[b] > 0.1, gtb[b] < 2.0)
        p = prb[b][m].copy()
        p[p < 0.1] = 0.1
        p[p > 2.0] = 2.0
        rows.append([float(v) for v in compute_depth_errors_numpy(gtb[b][m], p)])
    out["met_img_rows"] = np.array(rows)

    path = os.path.join(HERE, "reference_outputs.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
