"""GPU parity tests: every CUDA entry point (through the C ABI) against the CPU oracle and the golden
reference outputs, with BASELINE.json's tolerances (tests/parity.py)."""
import os

import numpy as np
import pytest
import torch

import parity as P
from oracle import polcue_oracle as O

pytestmark = pytest.mark.gpu

from polcue import ops, synth  # noqa: E402
from polcue import _lib  # noqa: E402
from polcue.compat import (trainer as c_depth, layers as c_layers, normals_vec as c_nv,  # noqa: E402
                           physical_normals_channels as c_ppp, pol_split_and_save as c_split,
                           xolp as c_xolp, xolp_and_normals as c_xn)


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def check_fused(mosaics, n=1.5, mufu=False):
    """mosaics: B x H x W uint8 numpy.  Full protocol of SURVEY 8c against the closed-form oracle."""
    with ops.trig("mufu" if mufu else "poly"):
        out = ops.fused_mosaic(dev(mosaics), n, want_iun=True, want_planes=True)
        torch.cuda.synchronize()
    worst = 0.0
    for b in range(mosaics.shape[0]):
        stack = O.stack_quadrants(mosaics[b])
        hs, ws = stack.shape[:2]
        assert np.array_equal(out["planes"][b].cpu().numpy(), stack.transpose(2, 0, 1)), "split not bit-exact"
        iun, rho, phi = O.iun_and_xolp_closed(stack)
        _, degenerate = O.xolp_tie_masks(stack)
        g_rho = out["xolp"][b, 0].cpu().numpy()
        g_phi = out["xolp"][b, 1].cpu().numpy()
        P.assert_dolp_close(out["iun"][b].cpu().numpy(), iun, "Iun")
        P.assert_dolp_close(g_rho, rho, "rho")
        P.assert_aolp_close(g_phi, phi)   # closed form vs closed form: no exclusions needed
        ref = O.get_normals(np.stack((rho, phi))[None], n).reshape(3, 3, hs, ws)
        got = out["normals"][b].cpu().numpy().reshape(3, 3, hs, ws)
        worst = max(worst, P.assert_normals_close(got, ref, axis=1))
        assert np.abs(np.linalg.norm(got, axis=1) - 1.0).max() < 3e-6
    return worst


@pytest.mark.parametrize("mufu", [False, True])
@pytest.mark.parametrize("kind", ["P", "U"])
def test_fused_small_batches(kind, mufu):
    check_fused(synth.gen_batch(kind, 0, 3, 128, 192), mufu=mufu)


@pytest.mark.parametrize("shape", [(2, 2), (2, 6), (6, 2), (10, 14), (12, 20), (34, 66), (130, 250)])
def test_fused_ragged_shapes_exercise_every_vector_width(shape):
    h, w = shape
    rng = np.random.default_rng(h * 1000 + w)
    check_fused(rng.integers(0, 256, (2, h, w), dtype=np.uint8))


def test_fused_randomized_shapes_generators_and_refractive_indices():
    """Differential sweep: 36 (shape, generator, n) combinations against the oracle, full protocol each."""
    rng = np.random.default_rng(2026)
    for case in range(36):
        hs, ws = int(rng.integers(1, 90)), int(rng.integers(1, 150))
        if case % 3 == 0:
            ws = 4 * max(1, ws // 4)                              # exercise the 16-byte path as well
        b = int(rng.integers(1, 4))
        n = float(rng.choice([1.2, 1.33, 1.5, 1.5, 1.5, 1.8, 2.4]))
        if case % 2:
            mosaics = rng.integers(0, 256, (b, 2 * hs, 2 * ws), dtype=np.uint8)
        else:
            mosaics = np.stack([synth.tile_mosaic(synth.gen_p_planes(case * 10 + f, hs, ws)) for f in range(b)])
        check_fused(mosaics, n=n, mufu=bool(case % 4 < 2))


def test_fused_edge_values():
    # flat, black, saturated, ties, single-channel pixels
    vals = np.array([[0, 0, 0, 0], [255, 255, 255, 255], [255, 0, 0, 0], [0, 255, 0, 0], [0, 0, 255, 0], [0, 0, 0, 255],
                     [1, 0, 0, 0], [0, 0, 1, 1], [10, 20, 30, 40], [200, 100, 0, 100], [100, 100, 0, 0], [5, 5, 9, 5]],
                    dtype=np.uint8)
    hs, ws = 4, vals.shape[0]
    planes = [np.tile(vals[:, k], (hs, 1)) for k in range(4)]
    check_fused(synth.tile_mosaic(planes)[None])


def test_fused_vs_golden_reference_outputs(golden):
    """Kernel vs the REFERENCE's own outputs (lstsq + scipy), tie pixels handled per SURVEY 8c."""
    for tag in ("u", "p"):
        st = golden[f"xolp_{tag}_in"]
        mosaic = synth.tile_mosaic([st[..., k] for k in range(4)])
        out = ops.fused_mosaic(dev(mosaic)[None], 1.5, want_iun=True)
        torch.cuda.synchronize()
        sign_tie, degenerate = O.xolp_tie_masks(st)
        P.assert_dolp_close(out["iun"][0].cpu().numpy(), golden[f"xolp_{tag}_iun"], "Iun")
        P.assert_dolp_close(out["xolp"][0, 0].cpu().numpy(), golden[f"xolp_{tag}_rho"], "rho")
        P.assert_aolp_close(out["xolp"][0, 1].cpu().numpy(), golden[f"xolp_{tag}_phi"], exclude=degenerate)
        got = out["normals"][0].cpu().numpy().reshape(3, 3, *st.shape[:2])
        for k, key in enumerate(("nd", "n1", "n2")):
            ref = golden[f"chain_{tag}_{key}"].transpose(2, 0, 1)
            skip = degenerate if key == "n2" else (degenerate & False)
            P.assert_normals_close(got[k], ref, axis=0, twin_ok=sign_tie | degenerate if key != "n2" else sign_tie, skip=skip,
                                   what=f"{tag}/{key}")
        ref_n2z = np.abs(golden[f"chain_{tag}_n2"][..., 2])
        assert np.abs(np.abs(got[2, 2]) - ref_n2z)[degenerate].max(initial=0) < 1e-3


def test_fused_full_frame_properties():
    """BASELINE full size (2448 x 2048): oracle on a strip + size-independent properties on the whole frame."""
    mosaic = synth.gen_batch("P", 5, 2)
    out = ops.fused_mosaic(dev(mosaic), 1.5, want_iun=True, want_planes=True)
    torch.cuda.synchronize()
    hs, ws = mosaic.shape[1] // 2, mosaic.shape[2] // 2
    # re-tiling the planes reproduces the mosaic bit-exactly
    pl = out["planes"].cpu().numpy()
    for b in range(2):
        assert np.array_equal(synth.tile_mosaic(list(pl[b])), mosaic[b])
    nrm = out["normals"]
    norms = nrm.view(2, 3, 3, hs, ws).norm(dim=2)
    assert float((norms - 1).abs().max()) < 3e-6
    phi = out["xolp"][:, 1]
    assert float(phi.abs().max()) <= np.pi / 2 + 1e-6
    # frames are independent: batch result == per-frame result, bit for bit
    single = ops.fused_mosaic(dev(mosaic[1:2]), 1.5)
    assert torch.equal(single["normals"][0], nrm[1]) and torch.equal(single["xolp"][0], out["xolp"][1])
    # oracle on a 64-row strip of frame 0 (rows 480..543 of each quadrant)
    stack = O.stack_quadrants(mosaic[0])[480:544]
    iun, rho, phi_ref = O.iun_and_xolp_closed(stack)
    P.assert_dolp_close(out["xolp"][0, 0, 480:544].cpu().numpy(), rho)
    P.assert_aolp_close(out["xolp"][0, 1, 480:544].cpu().numpy(), phi_ref)
    ref = O.get_normals(np.stack((rho, phi_ref))[None], 1.5).reshape(3, 3, 64, ws)
    P.assert_normals_close(nrm[0, :, 480:544].cpu().numpy().reshape(3, 3, 64, ws), ref, axis=1)


@pytest.mark.parametrize("kind", ["P", "U"])
def test_fused_benchmarked_configuration_64_full_frames(kind):
    """BASELINE configs[1] exactly as bench.py launches it: B = 64 mosaics of 2448 x 2048 in ONE launch (78 336 tiles handed
    out through cluster launch control).  (a) frames 0, 31, 63 against the oracle on three row strips each (first rows,
    middle, last rows); (b) EVERY frame bit-equal to its own single-frame launch; (c) the host entry point
    (polcue_fused_mosaic_u8_host, what `e2e` times) bit-equal on all frames."""
    B, H, W = 64, synth.FRAME_H, synth.FRAME_W
    hs, ws = H // 2, W // 2
    if kind == "P":
        mosaic = synth.gen_p_batch_torch(0, B, H, W, device="cuda")           # bench.py's generator
    else:
        gen = torch.Generator(device="cuda")
        gen.manual_seed(64)
        mosaic = torch.randint(0, 256, (B, H, W), dtype=torch.uint8, device="cuda", generator=gen)
    out = ops.fused_mosaic(mosaic, 1.5)
    torch.cuda.synchronize()
    # (a) oracle strips
    for f in (0, 31, 63):
        frame = mosaic[f].cpu().numpy()
        stack = O.stack_quadrants(frame)
        for r0 in (0, 480, hs - 48):
            rows = slice(r0, r0 + 48)
            _, rho, phi = O.iun_and_xolp_closed(stack[rows])
            P.assert_dolp_close(out["xolp"][f, 0, rows].cpu().numpy(), rho, f"rho frame {f} rows {r0}")
            P.assert_aolp_close(out["xolp"][f, 1, rows].cpu().numpy(), phi)
            ref = O.get_normals(np.stack((rho, phi))[None], 1.5).reshape(3, 3, 48, ws)
            P.assert_normals_close(out["normals"][f, :, rows].cpu().numpy().reshape(3, 3, 48, ws), ref, axis=1,
                                   what=f"normals frame {f} rows {r0}")
    # (b) every frame == its single-frame launch
    one = {}
    for f in range(B):
        one = ops.fused_mosaic(mosaic[f:f + 1], 1.5, out=one)
        assert torch.equal(one["xolp"][0], out["xolp"][f]) and torch.equal(one["normals"][0], out["normals"][f]), f
    # (c) host entry point on all frames (NUMA-local pinned buffers from the library)
    h_mosaic = ops.host_empty((B, H, W), torch.uint8)
    h_mosaic.copy_(mosaic)
    host = ops.fused_mosaic_host(h_mosaic, 1.5)
    for key in ("xolp", "normals"):
        assert torch.equal(host[key].cuda(), out[key]), key
    # unit length everywhere, AoLP in range
    norms = out["normals"].view(B, 3, 3, hs, ws).norm(dim=2)
    assert float((norms - 1).abs().max()) < 3e-6
    assert float(out["xolp"][:, 1].abs().max()) <= np.pi / 2 + 1e-6
    del host, h_mosaic, out, norms, mosaic
    torch.cuda.empty_cache()


def test_fused_without_optional_outputs_and_bad_args():
    mosaic = dev(synth.gen_u_mosaic(1, 64, 96))[None]
    full = ops.fused_mosaic(mosaic, 1.5, want_iun=True)
    lean = ops.fused_mosaic(mosaic, 1.5, want_normals=False)
    assert set(lean) == {"xolp"} and torch.equal(lean["xolp"], full["xolp"])
    with pytest.raises(ValueError):
        ops.fused_mosaic(mosaic[:, :63], 1.5)
    with pytest.raises(TypeError):
        ops.fused_mosaic(mosaic.cpu(), 1.5)
    L = _lib.lib()
    assert L.polcue_fused_mosaic_u8(None, 1, 64, 96, None, None, None, None, None, None) == _lib.EINVAL
    assert L.polcue_fused_mosaic_u8(mosaic.data_ptr(), 1, 63, 96, None, None, None, full["xolp"].data_ptr(), None, None) == _lib.EINVAL
    assert L.polcue_fused_mosaic_u8(mosaic.data_ptr(), 0, 64, 96, None, None, None, full["xolp"].data_ptr(), None, None) == 0


@pytest.mark.parametrize("shape", [(2, 2), (2, 6), (10, 14), (34, 66), (130, 250), (64, 8200), (2048, 2448)])
def test_xolp_only_streaming_kernel_equals_the_full_kernel(shape):
    """`want_normals=False` takes the plain streaming member of the fused family (four groups per thread, no table, no tile
    loop): its XOLP, Iun and planes must equal the full kernel's bit for bit on every vector width, ragged tails included,
    for the quadrant layout, the raw super-pixel layout and four separate planes."""
    h, w = shape
    rng = np.random.default_rng(h * 7 + w)
    b = 1 if h * w > 1_000_000 else 3
    mosaic = dev(rng.integers(0, 256, (b, h, w), dtype=np.uint8))
    full = ops.fused_mosaic(mosaic, 1.5, want_iun=True, want_planes=True)
    lean = ops.fused_mosaic(mosaic, 1.5, want_iun=True, want_planes=True, want_normals=False)
    assert set(lean) == {"xolp", "iun", "planes"}
    for key in lean:
        assert torch.equal(lean[key], full[key]), key
    sp_full = ops.fused_mosaic(mosaic, 1.5, superpixel=(2, 1, 3, 0))
    sp_lean = ops.fused_mosaic(mosaic, 1.5, superpixel=(2, 1, 3, 0), want_normals=False)
    assert torch.equal(sp_lean["xolp"], sp_full["xolp"])
    planes = [full["planes"][:, k].contiguous() for k in range(4)]
    _, x = ops.xolp_from_planes(*planes)
    assert torch.equal(x, full["xolp"])
    assert torch.equal(ops.fused_planes(*planes, want_normals=False)["xolp"], full["xolp"])


def test_fused_host_entry_point_matches_device_entry_point():
    mosaic = torch.from_numpy(synth.gen_batch("P", 9, 5, 128, 192)).pin_memory()
    host = ops.fused_mosaic_host(mosaic, 1.5, want_iun=True, chunk_frames=2)   # 3 chunks, ragged tail
    devo = ops.fused_mosaic(mosaic.cuda(), 1.5, want_iun=True)
    torch.cuda.synchronize()
    for key in ("xolp", "normals", "iun"):
        assert torch.equal(host[key], devo[key].cpu()), key
    host2 = ops.fused_mosaic_host(mosaic, 1.5, want_iun=True)                  # default chunking
    assert torch.equal(host2["normals"], host["normals"])


# ------------------------------------------------------------------------------------------------
def test_xolp_stack_u8_and_f32_vs_golden(golden):
    for tag in ("u", "p"):
        st = golden[f"xolp_{tag}_in"]
        iun, rho, phi = c_xolp.Iun_and_xolp(st, O.CANONICAL_ANGLES)
        assert iun.dtype == np.float64 and iun.shape == st.shape[:2]
        _, degenerate = O.xolp_tie_masks(st)
        P.assert_dolp_close(iun, golden[f"xolp_{tag}_iun"], "Iun")
        P.assert_dolp_close(rho, golden[f"xolp_{tag}_rho"], "rho")
        P.assert_aolp_close(phi, golden[f"xolp_{tag}_phi"], exclude=degenerate)
    fl = golden["xolp_f32_in"]
    iun, rho, phi = c_xolp.Iun_and_xolp(fl, O.CANONICAL_ANGLES)
    s1, s2 = fl[..., 0] - fl[..., 2], fl[..., 1] - fl[..., 3]
    P.assert_dolp_close(iun, golden["xolp_f32_iun"], "Iun")
    P.assert_dolp_close(rho, golden["xolp_f32_rho"], "rho")
    P.assert_aolp_close(phi, golden["xolp_f32_phi"], exclude=(s1 == 0) & (s2 == 0))


def test_xolp_known_answers(golden):
    iun, rho, phi = c_xolp.Iun_and_xolp(golden["kat_in"][None], O.CANONICAL_ANGLES)
    _, degenerate = O.xolp_tie_masks(golden["kat_in"][None])
    P.assert_dolp_close(iun[0], golden["kat_iun"])
    P.assert_dolp_close(rho[0], golden["kat_rho"])
    P.assert_aolp_close(phi[0], golden["kat_phi"], exclude=degenerate[0])
    assert rho[0, 0] == 0.0 and phi[0, 0] == 0.0          # all-zero pixel: the inf/nan scrub of xolp.py:26-29


def test_xolp_noncanonical_angles(golden):
    iun, rho, phi = c_xolp.Iun_and_xolp(golden["xolp_p_in"], golden["xolp_ang2"])
    P.assert_dolp_close(iun, golden["xolp_ang2_iun"], "Iun")
    P.assert_dolp_close(rho, golden["xolp_ang2_rho"], "rho")
    P.assert_aolp_close(phi, golden["xolp_ang2_phi"])


@pytest.mark.parametrize("shape", [(320, 480), (64, 96), (33, 47), (10, 14)])
def test_fused_planes_matches_fused_mosaic(shape):
    hs, ws = shape
    planes = [np.stack([synth.gen_p_planes(f, hs, ws)[k] for f in range(3)]) for k in range(4)]
    mosaic = np.stack([synth.tile_mosaic([planes[k][f] for k in range(4)]) for f in range(3)])
    a = ops.fused_planes(*(dev(p) for p in planes), want_iun=True)
    m = ops.fused_mosaic(dev(mosaic), 1.5, want_iun=True)
    for key in ("xolp", "normals", "iun"):
        assert torch.equal(a[key], m[key]), key
    # views into one [B,4,H,W] tensor and single images work too
    packed = dev(np.stack(planes, axis=1))
    v = ops.fused_planes(*(packed[:, k].contiguous() for k in range(4)))
    assert torch.equal(v["normals"], m["normals"])
    one = ops.fused_planes(*(dev(p[0]) for p in planes), want_normals=False)
    assert torch.equal(one["xolp"][0], m["xolp"][0])


@pytest.mark.parametrize("pattern", [(0, 1, 2, 3), (2, 1, 3, 0), (3, 0, 1, 2)])
@pytest.mark.parametrize("shape", [(64, 96), (33, 47), (10, 14), (1, 1)])
def test_fused_superpixel_layout_matches_planes(shape, pattern):
    hs, ws = shape
    planes = [np.stack([synth.gen_p_planes(f, hs, ws)[k] for f in range(2)]) for k in range(4)]
    raw = np.zeros((2, 2 * hs, 2 * ws), np.uint8)
    for pos, angle in enumerate(pattern):
        raw[:, pos // 2::2, pos % 2::2] = planes[angle]
    a = ops.fused_mosaic(dev(raw), 1.5, want_iun=True, want_planes=True, superpixel=pattern)
    ref = ops.fused_planes(*(dev(p) for p in planes), want_iun=True)
    for key in ("xolp", "normals", "iun"):
        assert torch.equal(a[key], ref[key]), key
    assert np.array_equal(a["planes"].cpu().numpy(), np.stack(planes, axis=1))       # de-interleave is bit-exact
    with pytest.raises(ValueError):
        ops.fused_mosaic(dev(raw), 1.5, superpixel=(0, 0, 1, 2))


def test_xolp_planes_matches_stack_path():
    planes = synth.gen_p_planes(2, 96, 128)
    _, x1 = ops.xolp_from_planes(*(dev(p)[None] for p in planes), want_iun=False)
    iun2, x2 = ops.xolp_from_stack(dev(np.stack(planes, axis=2)), None)
    iun1, _ = ops.xolp_from_planes(*(dev(p) for p in planes), want_iun=True)
    assert torch.equal(x1, x2) and torch.equal(iun1, iun2)
    with pytest.raises(ValueError):
        c_xolp.Iun_and_xolp(np.zeros((4, 4, 3), np.uint8), O.CANONICAL_ANGLES)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1.3, 1.5, 1.8])
def test_table_inversions_vs_golden(golden, n):
    rq = golden["tab_rho"]
    rho = dev(rq, torch.float32)[None, None]
    rq32 = rho.cpu().numpy().astype(np.float64).ravel()      # what the kernel actually saw
    knots = O.sorted_knots(n)

    def slope_at(name, q):
        xk, yk = knots[name]
        hi = np.clip(np.searchsorted(xk, q), 1, len(xk) - 1)
        return np.abs((yk[hi] - yk[hi - 1]) / (xk[hi] - xk[hi - 1]))

    td = c_nv.rho_diffuse(rho, n).cpu().numpy().ravel()
    t1, t2 = (t.cpu().numpy().ravel() for t in c_nv.rho_spec(rho, n))
    for name, got, key in (("diffuse", td, "tab_d"), ("spec1", t1, "tab_s1"), ("spec2", t2, "tab_s2")):
        ref = O.interp_linear_extrap(*knots[name], rq32)
        # float32 evaluation: 2 ulp of theta plus the float32 spacing of rho times the local slope
        tol = 4e-7 * (1 + np.abs(ref)) + slope_at(name, rq32) * 1.3e-7 * np.maximum(np.abs(rq32), np.abs(1 - rq32))
        assert (np.abs(got - ref) <= tol).all(), (name, np.abs(got - ref).max())
        # and against the reference's own float64 outputs for the float64 queries, where the table is not steep
        tame = slope_at(name, rq) < 20
        assert np.abs(got - golden[f"{key}_{n}"])[tame].max() < 5e-6


def test_get_normals_vs_golden(golden):
    for tag in ("u", "p"):
        x = dev(golden[f"getn_{tag}_x32"])
        got = c_nv.get_normals(x, 1.5)
        assert got.shape == (1, 9, 64, 96) and got.dtype == torch.float32
        ref = golden[f"getn_{tag}_n"]
        P.assert_normals_close(got.cpu().numpy().reshape(3, 3, 64, 96), ref.reshape(3, 3, 64, 96), axis=1)
        # composition of the three fine-grained mirrors gives the same normals
        rho, phi = x[:, 0], x[:, 1]
        parts = torch.cat((c_nv.calc_normals(phi, c_nv.rho_diffuse(rho, 1.5)),
                           c_nv.calc_normals(phi + np.pi / 2, c_nv.rho_spec(rho, 1.5)[0]),
                           c_nv.calc_normals(phi + np.pi / 2, c_nv.rho_spec(rho, 1.5)[1])), dim=1)
        P.assert_normals_close(parts.cpu().numpy().reshape(3, 3, 64, 96), ref.reshape(3, 3, 64, 96), axis=1)


def test_get_normals_odd_sizes_and_extreme_rho():
    rng = np.random.default_rng(3)
    x = np.stack((rng.uniform(-0.05, 2.0, (2, 7, 13)), rng.uniform(-np.pi / 2, np.pi / 2, (2, 7, 13))), axis=1).astype(np.float32)
    got = ops.get_normals(dev(x), 1.5).cpu().numpy()
    ref = O.get_normals(x, 1.5)
    P.assert_normals_close(got.reshape(2, 3, 3, 7, 13), ref.reshape(2, 3, 3, 7, 13), axis=2)


@pytest.mark.parametrize("mufu", [False, True])
def test_steep_end_segment_is_evaluated_in_float64(mufu):
    """n = 1.8: the second specular branch ends in a segment of slope -10797 (tests/test_cpu_host.py); beyond it theta
    reaches -1e4 rad, which only float64 holds to 1e-3 rad.  Every kernel family, both vector widths."""
    n = 1.8
    xk, yk = O.sorted_knots(n)["spec2"]
    rng = np.random.default_rng(18)
    with ops.trig("mufu" if mufu else "poly"):
        for shape in ((2, 16, 64), (1, 7, 13)):
            rho = np.concatenate((rng.uniform(0.99999, 1.00001, 500), rng.uniform(1.0, 2.5, 500), rng.uniform(0, 1, 1000),
                                  np.float32(xk[-3:]).astype(np.float64), np.nextafter(np.float32(xk[-2:]), np.float32(2)).astype(np.float64)))
            rho = rng.choice(rho, shape).astype(np.float32)
            x = np.stack((rho, rng.uniform(-np.pi / 2, np.pi / 2, shape).astype(np.float32)), axis=1)
            got = ops.get_normals(dev(x), n).cpu().numpy()
            ref = O.get_normals(x, n)
            b, h, w = shape
            P.assert_normals_close(got.reshape(b, 3, 3, h, w), ref.reshape(b, 3, 3, h, w), axis=2)
            # the fine-grained mirror returns float32 angles: the float64 line rounded once
            t2 = c_nv.rho_spec(dev(rho), n)[1].cpu().numpy()
            ref2 = O.interp_linear_extrap(xk, yk, rho.astype(np.float64))
            assert np.abs(ref2).max() > 1e4
            assert (np.abs(t2 - ref2) <= 6.0e-8 * np.abs(ref2) + 7e-6).all()
        # fused kernel: uniform bytes put ~10 % of the pixels beyond rho = 1; flat-topped pixels sit exactly on rho = 1
        mosaics = rng.integers(0, 256, (2, 60, 144), dtype=np.uint8)
        mosaics[0, :4, :8] = 200          # I0 = 200 ...
        mosaics[0, 30:34, :8] = 0         # ... I90 = 0, I45 = I135: rho = 1 exactly
        mosaics[0, :4, 72:80] = 100
        mosaics[0, 30:34, 72:80] = 100
        check_fused(mosaics, n=n, mufu=mufu)
        check_fused(mosaics[:, :, :142], n=n, mufu=mufu)       # Ws = 71: scalar path
        planes = [np.ascontiguousarray(O.stack_quadrants(mosaics[0])[..., k]) for k in range(4)]
        assert torch.equal(ops.fused_planes(*(dev(p)[None] for p in planes), n=n)["normals"],
                           ops.fused_mosaic(dev(mosaics[:1]), n)["normals"])


def test_get_normals_random_refractive_indices():
    """Differential sweep over n (incl. 1.5548, whose second specular branch ends with slope -53 006, and n near 1):
    XOLP-fed and byte-fed kernels against the float64 oracle on rho in [0, 2.2], both sincos variants."""
    rng = np.random.default_rng(23)
    for i, n in enumerate(np.concatenate((rng.uniform(1.02, 3.0, 8), [1.5548, 1.01]))):
        shape = (2, int(rng.integers(3, 40)), int(rng.integers(3, 60)))
        rho = np.where(rng.random(shape) < 0.5, rng.uniform(0, 1, shape) ** 2, rng.uniform(0, 2.2, shape))
        rho = np.where(rng.random(shape) < 0.1, 1 - rng.uniform(0, 1, shape) ** 3 * 1e-3, rho).astype(np.float32)
        x = np.stack((rho, rng.uniform(-np.pi / 2, np.pi / 2, shape).astype(np.float32)), axis=1)
        with ops.trig("mufu" if i & 1 else "poly"):
            got = ops.get_normals(dev(x), float(n)).cpu().numpy()
        ref = O.get_normals(x, float(n))
        b, h, w = shape
        P.assert_normals_close(got.reshape(b, 3, 3, h, w), ref.reshape(b, 3, 3, h, w), axis=2, what=f"n={n:.4f}")
        check_fused(rng.integers(0, 256, (1, 2 * h, 2 * w), dtype=np.uint8), n=float(n), mufu=bool(i & 1))


def test_numpy_chain_mirror(golden):
    st = golden["xolp_p_in"]
    mosaic = synth.tile_mosaic([st[..., k] for k in range(4)])
    iun, rho, phi, nd, n1, n2 = c_xn.process_frame(mosaic, 1.5)
    sign_tie, degenerate = O.xolp_tie_masks(st)
    P.assert_dolp_close(rho, golden["xolp_p_rho"])
    P.assert_normals_close(nd, golden["chain_p_nd"], axis=2, twin_ok=sign_tie | degenerate)
    P.assert_normals_close(n1, golden["chain_p_n1"], axis=2, twin_ok=sign_tie | degenerate)
    P.assert_normals_close(n2, golden["chain_p_n2"], axis=2, twin_ok=sign_tie, skip=degenerate)
    th = c_xn.rho_diffuse(golden["xolp_p_rho"], 1.5)
    assert th.dtype == np.float64 and np.abs(th - O.rho_diffuse(golden["xolp_p_rho"], 1.5)).max() < 5e-6
    nn = c_xn.calc_normals(golden["xolp_p_phi"], th)
    assert nn.shape == st.shape[:2] + (3,)
    P.assert_normals_close(nn, O.calc_normals_hw3(golden["xolp_p_phi"], th), axis=2)


def test_ppp_channel_variant_vs_golden(golden):
    rho, phi, iun = c_ppp.PolarisationImage_channel(golden["ppp_images"], O.CANONICAL_ANGLES, golden["ppp_mask"])
    assert np.array_equal(np.isnan(rho), np.isnan(golden["ppp_rho"]))
    ok = ~np.isnan(golden["ppp_rho"])
    P.assert_dolp_close(rho[ok], golden["ppp_rho"][ok])
    P.assert_dolp_close(iun, golden["ppp_iun"])
    s1 = golden["ppp_images"][..., 0] - golden["ppp_images"][..., 2]
    s2 = golden["ppp_images"][..., 1] - golden["ppp_images"][..., 3]
    P.assert_aolp_close(phi, golden["ppp_phi"], exclude=(s1 == 0) & (s2 == 0))
    mask = golden["ppp_mask"]
    th_d = c_ppp.rho_diffuse_channel(np.nan_to_num(golden["ppp_rho"]), 1.5)
    got = c_ppp.calc_normals_channel(golden["ppp_phi"], th_d, mask)
    ref = O.calc_normals_channel(golden["ppp_phi"], O.rho_diffuse(np.nan_to_num(golden["ppp_rho"]), 1.5), mask)
    assert np.array_equal(got[~mask], np.zeros_like(got[~mask]))
    P.assert_normals_close(got[mask], ref[mask], axis=1)


# ------------------------------------------------------------------------------------------------
def test_split_pol_bit_exact(golden):
    for tag in ("gray", "bgr"):
        q = c_split.split_pol(golden[f"split_{tag}_in"])
        for name, arr in zip(("im00", "im10", "im01", "im11"), q):
            assert np.array_equal(arr, golden[f"split_{tag}_{name}"]), (tag, name)
    with pytest.raises(ValueError):
        c_split.split_pol(np.zeros((5, 4), np.uint8))
    # full-size frame, every copy width (uint8 gray: 1224 % 8 == 0; BGR; float32; odd half-width)
    rng = np.random.default_rng(0)
    for shape, dtype in (((2048, 2448), np.uint8), ((64, 96, 3), np.uint8), ((32, 40), np.float32), ((6, 10), np.uint8),
                         ((6, 14), np.uint16)):
        img = rng.integers(0, 200, shape).astype(dtype)
        for got, ref in zip(c_split.split_pol(img), O.split_pol(img)):
            assert np.array_equal(got, ref)
    batch = rng.integers(0, 256, (3, 16, 24), dtype=np.uint8)
    for k, got in enumerate(ops.split_pol_batch(dev(batch))):
        for b in range(3):
            assert np.array_equal(got[b].cpu().numpy(), O.split_pol(batch[b])[k])


# ------------------------------------------------------------------------------------------------
def torch_sobel_restatement(depth, K):
    """float32 torch restatement of kornia's op sequence (pad replicate + conv with sobel/8), for cross-checking."""
    import torch.nn.functional as F
    b, _, h, w = depth.shape
    u = torch.arange(w, device=depth.device, dtype=depth.dtype)[None, None, :].expand(b, h, w)
    v = torch.arange(h, device=depth.device, dtype=depth.dtype)[None, :, None].expand(b, h, w)
    fx, fy, cx, cy = K[:, 0, 0, None, None], K[:, 1, 1, None, None], K[:, 0, 2, None, None], K[:, 1, 2, None, None]
    z = depth[:, 0]
    xyz = torch.stack(((u - cx) / fx * z, (v - cy) / fy * z, z), 1)
    kx = torch.tensor([[-1., 0., 1.], [-2., 0., 2.], [-1., 0., 1.]], device=depth.device, dtype=depth.dtype) / 8
    ker = torch.stack((kx, kx.t()))[:, None]
    g = F.conv2d(F.pad(xyz.reshape(b * 3, 1, h, w), (1, 1, 1, 1), mode="replicate"), ker).view(b, 3, 2, h, w)
    return F.normalize(torch.cross(g[:, :, 0], g[:, :, 1], dim=1), dim=1, p=2, eps=1e-12)


def holey_depth(first, batch, h, w):
    """GT depth as the reference feeds it (HAMMER: 0 = invalid): 10 % random invalid pixels, a depth step, plus -- on images
    large enough -- contiguous missing regions, an invalid border strip and an isolated valid pixel."""
    gt, pred, inst, k = synth.gen_depth_batch(first, batch, h, w)
    if min(h, w) >= 8:
        gt = synth.add_hole_regions(gt, first)
    return gt, pred, inst, k


@pytest.mark.parametrize("shape", [(320, 480), (64, 96), (37, 131), (5, 3), (1, 1), (9, 260), (832, 1088)])
def test_depth_to_normals(shape):
    """Hole-free surface: 1e-3 rad on every pixel.  GT with zero-depth holes (the data trainer.py:1305 actually passes):
    every pixel checked by the conditioning-aware protocol of parity.assert_stencil_normals_close -- 1e-3 rad wherever the
    float64 model says float32 can reach it, exact zero vectors inside holes, an explicit count bound on the rest."""
    h, w = shape
    batch = 3 if h * w < 500_000 else 1
    gt, _, _, k = holey_depth(3, batch, h, w)
    valid_only = np.where(gt > 0, gt, 0.7).astype(np.float32)          # smooth surface, no invalid holes
    torch.backends.cudnn.allow_tf32 = False                             # the float32 restatement below must not use TF32 convolutions
    got = c_depth.depth_to_normals(dev(valid_only)[:, None], dev(k)).cpu().numpy()
    ref = O.depth_to_normals(valid_only[:, None], k)
    P.assert_normals_close(got, ref, axis=1, tol=1e-3)
    assert np.abs(np.linalg.norm(got, axis=1) - np.linalg.norm(ref, axis=1)).max() < 1e-5     # unit length (zero on one-pixel-wide images)
    for depth in (gt, synth.gen_depth_batch(3, batch, h, w)[0]):        # with region holes, and with the random 10 % only
        got = c_depth.depth_to_normals(dev(depth)[:, None], dev(k)).cpu().numpy()
        worst, ill = P.assert_stencil_normals_close(got, depth[:, None], k, O, max_ill_fraction=3e-3 if h * w > 100_000 else 4e-2)
        # the float32 torch restatement of kornia's op sequence agrees too wherever float32 can reach the bound
        ref32 = torch_sobel_restatement(dev(depth)[:, None], dev(k)).cpu().numpy()
        _, bound, dead = O.depth_to_normals_conditioning(depth[:, None], k)
        well = (P.STENCIL_ROUNDINGS * 2.0 ** -24 * bound <= P.NORMAL_TOL) & ~dead
        err32 = P.angular_error(got, ref32, axis=1)
        assert (err32[well] <= 1e-3).all(), float(err32[well].max())
        assert (ref32[np.broadcast_to(dead[:, None], ref32.shape)] == 0).all()


KORNIA_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kornia_outputs.npz")


@pytest.mark.skipif(not os.path.exists(KORNIA_GOLDEN), reason="tests/golden/kornia_outputs.npz absent: run tools/pin_kornia.py where kornia 0.5.11 is installed")
def test_depth_to_normals_against_kornia_outputs():
    """The CUDA stencil against outputs of kornia 0.5.11 itself (fixture written by tools/pin_kornia.py), every pixel, under
    the conditioning-aware protocol; the reference's own float32 result must lie inside the same bounds."""
    data = np.load(KORNIA_GOLDEN)
    for name in sorted({k.split("/")[0] for k in data.files if "/" in k}):
        depth, km = data[name + "/depth"], data[name + "/K"]
        got = ops.depth_to_normals(dev(depth), dev(km)).cpu().numpy()
        _, bound, dead = O.depth_to_normals_conditioning(depth, km)
        lim = np.maximum(P.NORMAL_TOL, P.STENCIL_ROUNDINGS * 2.0 ** -24 * bound)
        for ref, slack in ((data[name + "/normals_f64"], 1.0), (data[name + "/normals_f32"], 2.0)):   # float32 vs float32: both carry round-off
            err = np.where(dead | ~np.isfinite(lim), 0.0, P.angular_error(got, ref, axis=1))
            assert not (~(err <= slack * lim)).any(), (name, float(np.nanmax(err)))
        assert (got.transpose(0, 2, 3, 1)[dead] == 0).all(), name


def test_stencil_metrics_and_loss_randomized_shapes_and_cameras():
    """Differential sweep: 16 random (batch, shape, camera, depth range, mask) cases through the stencil, the per-image
    metrics (range + material filter), the flat metric sums and the normals loss with its gradient."""
    rng = np.random.default_rng(99)
    for case in range(16):
        b, h, w = int(rng.integers(1, 4)), int(rng.integers(1, 70)), int(rng.integers(1, 200))
        if case % 4 == 0:
            w = 4 * max(1, w // 4)                                     # 16-byte rows: TMA staging
        v, u = np.mgrid[0:h, 0:w].astype(np.float64)
        depth = np.stack([0.3 + rng.uniform(0.2, 1.5) * (1 + 0.3 * np.sin(u / rng.uniform(5, 60) + rng.uniform(0, 6)) *
                                                          np.cos(v / rng.uniform(5, 60))) + rng.uniform(-2e-3, 2e-3) * u
                          for _ in range(b)]).astype(np.float32)
        k = np.stack([np.array([[rng.uniform(200, 900), 0, rng.uniform(0, w)], [0, rng.uniform(200, 900), rng.uniform(0, h)], [0, 0, 1]])
                      for _ in range(b)]).astype(np.float32)
        got = ops.depth_to_normals(dev(depth)[:, None], dev(k)).cpu().numpy()
        P.assert_normals_close(got, O.depth_to_normals(depth[:, None], k), axis=1, tol=1e-3, what=f"stencil case {case}")
        # metrics: prediction = noisy depth, 15 % invalid ground truth, random range and material
        pred = (depth * (1 + 0.1 * rng.standard_normal(depth.shape))).clip(0.05, 4).astype(np.float32)
        gt = np.where(rng.random(depth.shape) < 0.15, 0, depth).astype(np.float32)
        inst = (20 * rng.integers(0, 11, depth.shape)).astype(np.uint8)
        lo, hi = np.float32(rng.uniform(0.05, 0.4)), np.float32(rng.uniform(1.0, 3.0))
        for inst_id in (None, int(20 * rng.integers(0, 11))):
            sums, metrics = ops.depth_errors_per_image(dev(gt), dev(pred), float(lo), float(hi), dev(inst) if inst_id is not None else None, inst_id)
            rows, _ = O.depth_errors_per_image(gt, pred, lo, hi, inst, inst_id)
            gm = metrics.cpu().numpy().astype(np.float64)
            assert np.array_equal(np.isnan(gm), np.isnan(rows)), case
            assert np.allclose(gm, rows, rtol=1e-5, equal_nan=True), (case, np.nanmax(np.abs(gm / rows - 1)))
        m = gt > 0
        if m.any():
            flat = [float(x) for x in ops.compute_depth_errors(dev(gt[m]), dev(pred[m]))]
            assert np.allclose(flat, O.compute_depth_errors(gt[m], pred[m]), rtol=2e-5), case
        # loss + gradient against the float64 autograd oracle
        mask = (rng.random(depth.shape) < 0.8).astype(np.float32)
        if mask.sum() == 0 or min(h, w) < 2:
            continue        # one-pixel-wide images: every normal is the zero vector and the reference's gradient is round-off noise
        t64 = lambda a: torch.from_numpy(a.astype(np.float64))
        p64 = t64(pred)[:, None].requires_grad_(True)
        ref_loss = O.normals_loss_torch(t64(depth)[:, None], p64, t64(k), t64(mask)[:, None])
        ref_loss.backward()
        ref_grad = p64.grad[:, 0].numpy()
        dp = dev(pred)[:, None].clone().requires_grad_(True)
        loss = ops.normals_loss(dev(depth)[:, None], dp, dev(k), dev(mask)[:, None])
        loss.backward()
        assert abs(float(loss.detach()) - float(ref_loss.detach())) < 3e-5 * max(1.0, abs(float(ref_loss.detach()))), case
        g = dp.grad[:, 0].cpu().numpy().astype(np.float64)
        scale = np.abs(ref_grad).max() + 1e-30
        err = np.abs(g - ref_grad) / scale          # float32 kernel vs float64 autograd, as in the fixed-shape test below
        assert err.max() < 2e-3 and np.sqrt((err ** 2).mean()) < 2e-4, (case, float(err.max()))


def test_depth_to_normals_uses_tma_staging_when_rows_are_16_byte_multiples():
    L = _lib.lib()
    gt, _, _, k = synth.gen_depth_batch(0, 2, 64, 96)
    before = L.polcue_debug_stencil_tma_launches()
    ops.depth_to_normals(dev(gt)[:, None], dev(k))
    assert L.polcue_debug_stencil_tma_launches() == before + 1
    gt, _, _, k = synth.gen_depth_batch(0, 2, 37, 131)
    ops.depth_to_normals(dev(gt)[:, None], dev(k))              # odd width: manually staged tile
    assert L.polcue_debug_stencil_tma_launches() == before + 1


def _loss_case(shape, batch=3, first=0):
    h, w = shape
    gt, pred, _, k = synth.gen_depth_batch(first, batch, h, w)
    smooth = np.where(gt > 0, gt, 0.7).astype(np.float32)
    # a smooth prediction: noise-free surface with a different low-frequency error per image (what a network outputs)
    vv, uu = np.mgrid[0:h, 0:w].astype(np.float32)
    pred_s = (smooth * (1.0 + 0.05 * np.sin(uu / 23.0 + np.arange(batch)[:, None, None]) * np.cos(vv / 17.0))).astype(np.float32)
    mask = ((gt >= 0.1) & (gt <= 2.0)).astype(np.float32)          # trainer.py:1242-1243
    return smooth, pred_s, mask, k


@pytest.mark.parametrize("shape", [(320, 480), (64, 96), (37, 131), (16, 128), (17, 129), (5, 3), (1, 1), (2, 260)])
def test_normals_loss_forward_and_backward_vs_autograd_oracle(shape):
    gt, pred, mask, k = _loss_case(shape)
    if mask.sum() == 0:
        mask[:] = 1.0
    t64 = lambda a: torch.from_numpy(a.astype(np.float64))
    p64 = t64(pred)[:, None].requires_grad_(True)
    ref = O.normals_loss_torch(t64(gt)[:, None], p64, t64(k), t64(mask)[:, None])
    ref.backward()
    ref_grad = p64.grad[:, 0].numpy()

    d_pred = dev(pred)[:, None].requires_grad_(True)
    loss = ops.normals_loss(dev(gt)[:, None], d_pred, dev(k), dev(mask)[:, None])
    assert loss.dim() == 0 and loss.dtype == torch.float32
    assert abs(float(loss) - float(ref)) < 2e-5 * max(1.0, abs(float(ref)))
    (3.0 * loss).backward()                                           # upstream gradient 3
    got = d_pred.grad[:, 0].cpu().numpy() / 3.0
    scale = np.abs(ref_grad).max() + 1e-30
    err = np.abs(got - ref_grad)
    # float32 kernel vs float64 autograd: gradients of normalised cross products cancel heavily where the surface is flat
    assert err.max() < 2e-3 * scale, (err.max(), scale)
    assert np.sqrt((err ** 2).mean()) < 2e-4 * scale


@pytest.mark.parametrize("shape", [(320, 480), (37, 131), (16, 128), (2, 5)])
def test_supervised_depth_and_normals_losses_fused(shape):
    """trainer.py:1240-1251 for one scale: mask from the GT range, masked L1 depth loss and the normals loss, both with
    gradients, against float64 torch autograd; and the normals part bit-identical to the stand-alone kernels."""
    from polcue.compat import trainer as c_tr
    gt, pred, _, k = _loss_case(shape)
    gt = gt.copy()
    gt[:, : max(1, shape[0] // 8)] += 1.2                             # beyond max_depth: outside the range mask
    lo, hi = 0.1, 1.6
    t64 = lambda a: torch.from_numpy(a.astype(np.float64))
    g64, p64 = t64(gt)[:, None], t64(pred)[:, None].requires_grad_(True)
    m64 = (g64 >= lo).double() * (g64 <= hi).double()
    ref_depth = ((g64 - p64).abs() * m64).sum() / m64.sum()
    ref_normals = O.normals_loss_torch(g64, p64, t64(k), m64)
    (0.7 * ref_depth + 1.3 * ref_normals).backward()
    ref_grad = p64.grad[:, 0].numpy()

    k44 = torch.eye(4).repeat(gt.shape[0], 1, 1)
    k44[:, :3, :3] = torch.from_numpy(k)
    d_pred = dev(pred)[:, None].requires_grad_(True)
    depth_loss, normals_loss = c_tr.compute_supervised_losses(dev(gt)[:, None], d_pred, k44.cuda(), lo, hi)
    assert abs(float(depth_loss.detach()) - float(ref_depth.detach())) < 2e-6 * max(1.0, abs(float(ref_depth.detach())))
    assert abs(float(normals_loss.detach()) - float(ref_normals.detach())) < 2e-5 * max(1.0, abs(float(ref_normals.detach())))
    (0.7 * depth_loss + 1.3 * normals_loss).backward()
    got = d_pred.grad[:, 0].cpu().numpy()
    scale = np.abs(ref_grad).max() + 1e-30
    err = np.abs(got - ref_grad)
    assert err.max() < 2e-3 * scale and np.sqrt((err ** 2).mean()) < 2e-4 * scale, (err.max(), scale)
    # the normals half equals the stand-alone kernels given the same mask, bit for bit; the L1 gradient is exact
    mask = ((dev(gt) >= lo).float() * (dev(gt) <= hi).float())[:, None]
    d2 = dev(pred)[:, None].requires_grad_(True)
    alone = ops.normals_loss(dev(gt)[:, None], d2, dev(k), mask)
    assert torch.equal(alone.detach(), normals_loss.detach())
    alone.backward()
    d3 = dev(pred)[:, None].requires_grad_(True)
    ops.supervised_losses(dev(gt)[:, None], d3, dev(k), lo, hi)[1].backward()
    assert torch.equal(d3.grad, d2.grad)
    d4 = dev(pred)[:, None].requires_grad_(True)
    ops.supervised_losses(dev(gt)[:, None], d4, dev(k), lo, hi)[0].backward()
    expect = torch.sign(d4.detach() - dev(gt)[:, None]) * mask / mask.sum()
    assert torch.allclose(d4.grad, expect, rtol=1e-6, atol=0)


def _ill_conditioned_neighbourhood(gt, k):
    """Pixels whose GT normal is a round-off direction in ANY float32 evaluation (parallel gradients next to holes), and
    the pixels within one step of them (whose gradient gathers from them)."""
    _, bound, dead = O.depth_to_normals_conditioning(gt[:, None], k)
    ill = (P.STENCIL_ROUNDINGS * 2.0 ** -24 * bound > P.NORMAL_TOL) & ~dead
    near = ill.copy()
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            near |= np.roll(np.roll(ill, dy, axis=1), dx, axis=2)
    return ill, near


@pytest.mark.parametrize("shape", [(320, 480), (64, 96), (37, 131), (16, 128), (9, 260)])
def test_losses_on_ground_truth_with_zero_depth_holes(shape):
    """The normal HAMMER case (trainer.py:1242-1243, :1298-1309): depth_gt == 0 marks invalid pixels, next to valid ones.
    Forward and backward of the stand-alone normals loss (range mask passed as a tensor) and of the supervised block
    (mask derived from the staged GT tile) against float64 torch autograd.  Where the GT normal is a pure round-off
    direction (exactly parallel gradients: a handful of pixels, counted) the loss may differ by at most 2 / sum(mask) per
    such pixel and the gradient is compared outside their 3 x 3 neighbourhoods; everything else is held to the usual bounds."""
    h, w = shape
    gt, _, _, k = holey_depth(11, 3, h, w)
    assert 0.1 < (gt == 0).mean() < 0.6
    vv, uu = np.mgrid[0:h, 0:w].astype(np.float32)
    smooth = np.where(gt > 0, gt, 0.7).astype(np.float32)
    pred = (smooth * (1.0 + 0.05 * np.sin(uu / 23.0 + np.arange(3)[:, None, None]) * np.cos(vv / 17.0))).astype(np.float32)
    lo, hi = 0.1, 2.0
    mask = ((gt >= lo) & (gt <= hi)).astype(np.float32)
    ill, near = _ill_conditioned_neighbourhood(gt, k)
    n_ill = int((ill & (mask > 0)).sum())
    assert n_ill <= max(32, 2e-4 * mask.size), n_ill
    t64 = lambda a: torch.from_numpy(a.astype(np.float64))
    g64, m64 = t64(gt)[:, None], t64(mask)[:, None]
    p64 = t64(pred)[:, None].requires_grad_(True)
    ref_normals = O.normals_loss_torch(g64, p64, t64(k), m64)
    ref_depth = ((g64 - p64).abs() * m64).sum() / m64.sum()
    (1.3 * ref_normals + 0.7 * ref_depth).backward()
    ref_grad = p64.grad[:, 0].numpy()
    slack = 2.0 * n_ill / mask.sum()

    def check_grad(got, ref):
        scale = np.abs(ref).max() + 1e-30
        err = np.abs(got - ref) / scale
        assert np.isfinite(got).all()
        assert err[~near].max() < 2e-3 and np.sqrt((err[~near] ** 2).mean()) < 2e-4, (float(err[~near].max()), scale)
        assert np.abs(got[near]).max(initial=0.0) <= 4.0 * scale + 1e-30           # bounded next to the round-off normals too

    # stand-alone normals loss, mask tensor
    d1 = dev(pred)[:, None].requires_grad_(True)
    loss = ops.normals_loss(dev(gt)[:, None], d1, dev(k), dev(mask)[:, None])
    assert abs(float(loss.detach()) - float(ref_normals.detach())) <= 2e-5 * abs(float(ref_normals.detach())) + slack
    loss.backward()
    p2 = t64(pred)[:, None].requires_grad_(True)
    O.normals_loss_torch(g64, p2, t64(k), m64).backward()
    check_grad(d1.grad[:, 0].cpu().numpy().astype(np.float64), p2.grad[:, 0].numpy())
    # supervised block: range mask from the GT tile, L1 + normals
    d2 = dev(pred)[:, None].requires_grad_(True)
    depth_loss, normals_loss = ops.supervised_losses(dev(gt)[:, None], d2, dev(k), lo, hi)
    assert torch.equal(normals_loss.detach(), loss.detach())                     # same kernel arithmetic, same mask
    assert abs(float(depth_loss.detach()) - float(ref_depth.detach())) < 2e-6 * max(1.0, abs(float(ref_depth.detach())))
    (1.3 * normals_loss + 0.7 * depth_loss).backward()
    check_grad(d2.grad[:, 0].cpu().numpy().astype(np.float64), ref_grad)
    # invalid GT pixels get no L1 gradient and pull no normals gradient through their own mask
    only_l1 = dev(pred)[:, None].requires_grad_(True)
    ops.supervised_losses(dev(gt)[:, None], only_l1, dev(k), lo, hi)[0].backward()
    assert float(only_l1.grad[:, 0][dev(gt) == 0].abs().max()) == 0.0


def _misaligned(t):
    """The same values at an address that is 4 mod 16 bytes: forces the scalar (hand-staged, unpacked) kernels."""
    flat = torch.empty(t.numel() + 1, dtype=t.dtype, device=t.device)
    view = flat[1:].view(t.shape)
    view.copy_(t)
    assert view.data_ptr() % 16 == 4
    return view


@pytest.mark.parametrize("shape", [(320, 480), (64, 96), (36, 132), (16, 128), (2, 260), (45, 4)])
def test_packed_loss_kernels_are_bit_identical_to_the_scalar_kernels(shape):
    """Rows that are 16-byte multiples take the packed-FP32 kernels (GT and prediction in the two lanes of FFMA2 / FADD2 /
    FMUL2, interleaved shared tile); every lane runs the scalar kernels' operation sequence, so loss, sums and gradients must
    be BIT-IDENTICAL to the scalar kernels (reached here through 16-byte-misaligned views of the same data).  GT with holes."""
    h, w = shape
    gt, _, _, k = holey_depth(21, 3, h, w)
    vv, uu = np.mgrid[0:h, 0:w].astype(np.float32)
    pred = (np.where(gt > 0, gt, 0.7) * (1.0 + 0.05 * np.sin(uu / 23.0 + np.arange(3)[:, None, None]) * np.cos(vv / 17.0))).astype(np.float32)
    rng = np.random.default_rng(5)
    mask = (rng.random(gt.shape) < 0.8).astype(np.float32)
    g, pr, m, kk = dev(gt)[:, None], dev(pred)[:, None], dev(mask)[:, None], dev(k)
    results = []
    for variant in ("packed", "scalar"):
        gg = g if variant == "packed" else _misaligned(g)
        d1 = pr.clone().requires_grad_(True)
        loss = ops.normals_loss(gg, d1, kk, m)
        (1.7 * loss).backward()
        d2 = pr.clone().requires_grad_(True)
        dl, nl = ops.supervised_losses(gg, d2, kk, 0.1, 1.6)
        (0.7 * dl + 1.3 * nl).backward()
        results.append((loss.detach(), d1.grad, dl.detach(), nl.detach(), d2.grad))
    for a, b in zip(*results):
        assert torch.equal(a, b)
    assert float(results[0][1].abs().max()) > 0


def test_normals_loss_matches_composition_of_public_ops_and_is_deterministic():
    gt, pred, mask, k = _loss_case((64, 96))
    args = (dev(gt)[:, None], dev(pred)[:, None], dev(k), dev(mask)[:, None])
    loss = ops.normals_loss(*args)
    n_gt, n_pred = ops.depth_to_normals(args[0], args[2]), ops.depth_to_normals(args[1], args[2])
    cos = torch.nn.functional.cosine_similarity(n_gt.double(), n_pred.double(), dim=1).unsqueeze(1)
    ref = ((2 - cos) * args[3]).sum() / args[3].sum()
    assert abs(float(loss) - float(ref)) < 1e-6
    # (the forward kernel evaluates the stencil in a cancellation-free form, the public op in the reference's operation
    # order: a float64 re-evaluation of the cosine from the public op's float32 normals agrees to 1e-6)
    for misalign in (False, True):
        d = _misaligned(args[0]) if misalign else args[0]
        assert torch.equal(ops.depth_to_normals(d, args[2]), n_gt)
    assert torch.equal(loss, ops.normals_loss(*args))
    from polcue.compat import trainer as c_trainer
    k44 = torch.eye(4, device="cuda")[None].repeat(3, 1, 1)
    k44[:, :3, :3] = args[2]
    assert torch.equal(c_trainer.compute_supervised_normals_losses(args[0], args[1], k44, args[3]), loss)


def test_normals_loss_gradient_by_finite_differences():
    gt, pred, mask, k = _loss_case((24, 40), batch=1)
    args = (dev(gt)[:, None], dev(k), dev(mask)[:, None])
    d_pred = dev(pred)[:, None].requires_grad_(True)
    ops.normals_loss(args[0], d_pred, args[1], args[2]).backward()
    grad = d_pred.grad[0, 0].cpu().numpy()
    t64 = lambda a: torch.from_numpy(a.astype(np.float64))
    base = lambda p: float(O.normals_loss_torch(t64(gt)[:, None], t64(p)[:, None], t64(k), t64(mask)[:, None]))
    for (y, x) in ((0, 0), (0, 17), (23, 39), (12, 20), (11, 0), (23, 5)):      # corners, borders, interior
        e = np.zeros_like(pred, dtype=np.float64)
        e[0, y, x] = 1e-5
        fd = (base(pred.astype(np.float64) + e) - base(pred.astype(np.float64) - e)) / 2e-5
        assert abs(grad[y, x] - fd) < 2e-3 * np.abs(grad).max() + 1e-7, ((y, x), grad[y, x], fd)


def test_depth_to_normals_properties():
    k = dev(synth.scaled_intrinsics(64, 96)[None].astype(np.float32))
    flat = torch.full((1, 1, 64, 96), 0.8, device="cuda")
    n = ops.depth_to_normals(flat, k)
    assert torch.allclose(n[:, 2], torch.ones_like(n[:, 2]), atol=1e-6) and float(n[:, :2].abs().max()) < 1e-5
    # invariance to depth scale
    gt = dev(np.where(synth.gen_depth_batch(0, 1, 64, 96)[0] > 0, 0.9, 0.9).astype(np.float32))[:, None]
    v = torch.arange(64, device="cuda", dtype=torch.float32)[None, None, :, None]
    ramp = gt + 0.002 * v
    a, b2 = ops.depth_to_normals(ramp, k), ops.depth_to_normals(3.0 * ramp, k)
    assert P.angular_error(a.cpu().numpy(), b2.cpu().numpy(), axis=1).max() < 2e-4
    with pytest.raises(NotImplementedError):
        ops.depth_to_normals(ramp.requires_grad_(True), k)
    with pytest.raises(ValueError):
        ops.depth_to_normals(ramp.detach()[0], k)


# ------------------------------------------------------------------------------------------------
def test_depth_errors_vs_golden(golden):
    gt, pred = dev(golden["met_gt"]), dev(golden["met_pred"])
    got = np.array([float(v) for v in c_layers.compute_depth_errors(gt, pred)])
    assert np.allclose(got, golden["met_torch"], rtol=2e-5)
    assert got[4] == golden["met_torch"][4] and got[5] == golden["met_torch"][5] and got[6] == golden["met_torch"][6]
    got_np = np.array(c_layers.compute_depth_errors_numpy(golden["met_gt"], golden["met_pred"]), dtype=np.float64)
    assert np.allclose(got_np, golden["met_numpy"], rtol=2e-5)


@pytest.mark.parametrize("count", [1, 3, 31, 257, 4099, 1_843_200, 5_000_003])
def test_depth_error_sums_sizes(count):
    rng = np.random.default_rng(count)
    gt = (rng.random(count) * 1.9 + 0.1).astype(np.float32)
    pred = np.clip(gt * (1 + 0.3 * (rng.random(count) - 0.5)), 0.1, 2.0).astype(np.float32)
    sums, metrics = ops.depth_error_sums(dev(gt), dev(pred))
    ref = O.depth_error_sums(gt, pred)
    s = sums.cpu().numpy()
    assert np.array_equal(s[:4], ref[:4])                      # integer counts: bit-exact
    assert np.allclose(s[4:], ref[4:], rtol=3e-6)
    assert np.allclose(metrics.cpu().numpy(), O.compute_depth_errors(gt, pred), rtol=3e-6)
    # deterministic: same bits on a second launch (fixed reduction order), also from an unaligned view
    sums2, _ = ops.depth_error_sums(dev(gt), dev(pred))
    assert torch.equal(sums, sums2)
    if count > 8:
        s3, _ = ops.depth_error_sums(dev(np.concatenate(([1.0], gt)).astype(np.float32))[1:], dev(np.concatenate(([1.0], pred)).astype(np.float32))[1:])
        assert np.array_equal(s3.cpu().numpy()[:4], ref[:4]) and np.allclose(s3.cpu().numpy()[4:], ref[4:], rtol=3e-6)


def test_threshold_counts_follow_float32_division_semantics():
    """a1/a2/a3 decide borderline pixels exactly like the reference's float32 quotients (layers.py:542-545)."""
    rng = np.random.default_rng(1)
    n = 2_000_000
    lo = (rng.random(n) * 1.9 + 0.1).astype(np.float32)
    thr = rng.choice(np.array([1.25, 1.5625, 1.953125]), n)
    hi = (lo.astype(np.float64) * thr).astype(np.float32)
    hi = (hi.view(np.int32) + rng.integers(-3, 4, n).astype(np.int32)).view(np.float32)   # within 3 ulp of a threshold
    swap = rng.random(n) < 0.5
    gt, pred = np.where(swap, hi, lo), np.where(swap, lo, hi)
    ratio = np.maximum(gt / pred, pred / gt)                    # float32, as numpy/torch evaluate it
    want = [(ratio < np.float32(1.25 ** k)).sum() for k in (1, 2, 3)]
    sums, _ = ops.depth_error_sums(dev(gt), dev(pred))
    assert [int(v) for v in sums[1:4].cpu().numpy()] == [int(v) for v in want]
    bsums, _ = ops.depth_errors_per_image(dev(gt.reshape(4, -1)), dev(pred.reshape(4, -1)), 0.05, 5.0)
    assert [int(v) for v in bsums[:, 1:4].sum(0).cpu().numpy()] == [int(v) for v in want]


def test_depth_errors_empty_is_nan():
    sums, metrics = ops.depth_error_sums(torch.empty(0, device="cuda"), torch.empty(0, device="cuda"))
    assert float(sums[0]) == 0.0 and torch.isnan(metrics).all()


@pytest.mark.parametrize("shape", [(320, 480), (64, 96), (33, 47)])
def test_depth_errors_per_image(golden, shape):
    h, w = shape
    gt, pred, inst, _ = synth.gen_depth_batch(0, 5, h, w)
    for inst_id in (None, 40, 180, 7):
        sums, metrics = ops.depth_errors_per_image(dev(gt), dev(pred), 0.1, 2.0, dev(inst) if inst_id is not None else None, inst_id)
        rows, _ = O.depth_errors_per_image(gt, pred, np.float32(0.1), np.float32(2.0), inst, inst_id)
        got = metrics.cpu().numpy().astype(np.float64)
        assert np.array_equal(np.isnan(got), np.isnan(rows))
        assert np.allclose(got, rows, rtol=5e-6, equal_nan=True)
        for b in range(5):
            m = (gt[b] > np.float32(0.1)) & (gt[b] < np.float32(2.0))
            if inst_id is not None:
                m &= inst[b] == inst_id
            assert float(sums[b, 0]) == m.sum()
    if shape == (64, 96):
        _, metrics = ops.depth_errors_per_image(dev(gt[:3]), dev(pred[:3]), 0.1, 2.0)
        assert np.allclose(metrics.cpu().numpy(), golden["met_img_rows"], rtol=2e-5)


@pytest.mark.parametrize("shape", [(320, 480), (33, 47), (1, 5), (2, 2)])
def test_median_scaling_of_the_self_supervised_configurations(shape):
    """trainer.py:1413-1414: the masked medians are exact (radix select), even and odd counts, ties, empty masks."""
    h, w = shape
    gt, pred, inst, _ = synth.gen_depth_batch(2, 6, h, w)
    pred = (pred * np.float32(1.7)).astype(np.float32)                 # a scale-ambiguous prediction, as a self-supervised model gives
    pred[1] = np.round(pred[1], 1)                                     # many equal values around the median
    gt[2, : max(1, h // 2)] = 0                                        # changes the count's parity
    gt[3] = 0                                                          # empty mask
    inst[4] = 40
    for inst_id in (None, 40):
        med, scale = ops.masked_median_scale(dev(gt), dev(pred), 0.1, 2.0, dev(inst) if inst_id is not None else None, inst_id)
        med, scale = med.cpu().numpy(), scale.cpu().numpy()
        for b in range(6):
            m = (gt[b] > np.float32(0.1)) & (gt[b] < np.float32(2.0))
            if inst_id is not None:
                m &= inst[b] == inst_id
            if m.sum() == 0:
                assert np.isnan(med[b]).all() and np.isnan(scale[b])
                continue
            assert med[b, 0] == np.median(gt[b][m]) and med[b, 1] == np.median(pred[b][m]), (b, m.sum())      # bit-exact
            assert scale[b] == np.median(gt[b][m]) / np.median(pred[b][m])
        _, metrics = ops.depth_errors_per_image(dev(gt), dev(pred), 0.1, 2.0, dev(inst) if inst_id is not None else None, inst_id,
                                                median_scaling=True)
        rows, _ = O.depth_errors_per_image(gt, pred, 0.1, 2.0, inst, inst_id, median_scaling=True)
        got = metrics.cpu().numpy().astype(np.float64)
        assert np.array_equal(np.isnan(got), np.isnan(rows))
        assert np.allclose(got, rows, rtol=1e-5, equal_nan=True)
    # without scaling the scaled entry point is the plain one, bit for bit
    a = ops.depth_errors_per_image(dev(gt), dev(pred), 0.1, 2.0)[0]
    L = _lib.lib()
    import ctypes as C
    sums = torch.empty((6, 8), dtype=torch.float64, device="cuda")
    assert L.polcue_depth_errors_images_f32(dev(gt).data_ptr(), dev(pred).data_ptr(), None, 6, h * w, C.c_float(0.1), C.c_float(2.0), 0,
                                            sums.data_ptr(), None, C.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    assert torch.equal(torch.nan_to_num(a), torch.nan_to_num(sums))


@pytest.mark.parametrize("supervised", [True, False])
def test_evaluation_loop_mirror(supervised, capsys):
    """compute_depth_losses_from_list called the way Trainer.test / Evaluation.test call it (lists of CPU batches, one call
    per material group incl. evaluation.py's id RANGE "objects"), against a numpy transcription of the reference loop."""
    import types
    from polcue.compat import trainer as c_tr
    h, w, bsz = 64, 96, 3
    opt = types.SimpleNamespace(min_depth=0.1, max_depth=2.0, height=h, width=w, batch_size=bsz, depth_supervision=supervised,
                                train_stereo_only=False)
    me = types.SimpleNamespace(opt=opt, depth_metric_names=list(c_tr.DEPTH_METRIC_NAMES))
    gts, preds, masks = [], [], []
    for k in range(2):
        gt, pred, inst, _ = synth.gen_depth_batch(10 * k, bsz, h, w)
        pred = (pred * np.float32(1.0 if supervised else 1.6) + np.float32(0.0)).astype(np.float32)
        pred[0, :4] = 5.0                                              # beyond max_depth: exercises the batch-level clamp
        gts.append(torch.from_numpy(gt)[:, None]); preds.append(torch.from_numpy(pred)[:, None])
        masks.append(torch.from_numpy(inst.astype(np.int32))[:, None])
    for obj in ("all", "bottle", "wall", "objects"):
        losses = {}
        got = c_tr.compute_depth_losses_from_list(me, gts, preds, losses, masks, object=obj)
        rows = []
        inst_id = None if obj == "all" else c_tr.OBJECT_IDS[obj]
        for k in range(2):
            r, _ = O.depth_errors_per_image(gts[k][:, 0].numpy(), preds[k][:, 0].numpy(), 0.1, 2.0, masks[k][:, 0].numpy(), inst_id,
                                            median_scaling=not supervised, clamp_first=True)
            rows.append(r)
        ref = np.concatenate(rows).mean(0)
        assert np.allclose(got, ref, rtol=1e-5, equal_nan=True), (obj, got, ref)
        assert set(losses) == set(c_tr.DEPTH_METRIC_NAMES) and losses["de/abs_rel"].shape == ()
    assert "abs_rel |" in capsys.readouterr().out
    if supervised:      # the id-range group equals the sum of its materials' accumulators from the single-pass kernel
        gt, pred, inst = (dev(t[0][:, 0].numpy()) for t in (gts, preds, masks))
        inst = inst.to(torch.uint8)
        levels = list(range(20, 161, 20))
        sums, _ = ops.depth_errors_groups(gt, pred.clamp(0.1, 2.0), inst, 0.1, 2.0, levels)
        union = ops.metrics_from_sums(sums.sum(1)).cpu().numpy()
        _, direct = ops.depth_errors_per_image(gt, pred, 0.1, 2.0, inst, (20, 160), clamp_first=True)
        assert np.allclose(union, direct.cpu().numpy(), rtol=1e-6, equal_nan=True)


@pytest.mark.parametrize("shape", [(3, 2, 64, 96), (2, 9, 33, 47), (5, 11, 320, 480), (1, 1, 1, 1)])
def test_channel_stats_and_xolp_statistics(shape):
    rng = np.random.default_rng(sum(shape))
    x = (rng.normal(0.3, 0.8, shape)).astype(np.float32)
    got = ops.channel_stats(dev(x)).cpu().numpy()
    x64 = x.astype(np.float64)
    # float32 inside a tile of 1024 elements (the canonical order shared with the fused kernel's by-product), float64 above
    assert np.allclose(got[:, 0], x64.sum(axis=(0, 2, 3)), rtol=0, atol=2e-7 * np.abs(x64).sum(axis=(0, 2, 3)).max() + 1e-12)
    assert np.allclose(got[:, 1], (x64 ** 2).sum(axis=(0, 2, 3)), rtol=2e-7)
    assert torch.equal(ops.channel_stats(dev(x)), ops.channel_stats(dev(x)))          # fixed reduction order
    if shape[1] == 2:
        st = ops.xolp_statistics(dev(x))
        ref = O.xolp_statistics(x[:, 0], x[:, 1])
        for key, val in ref.items():
            assert abs(st[key] - val) < 1e-6 * max(1.0, abs(val)), key


@pytest.mark.parametrize("case", [("P", 3, 128, 192, 1.5), ("U", 2, 64, 264, 1.5), ("P", 5, 250, 328, 1.33), ("U", 1, 2048, 2448, 1.5),
                                  ("U", 2, 34, 66, 1.5), ("U", 2, 60, 144, 1.8)])
def test_statistics_by_product_of_the_fused_kernel_equals_the_statistics_pass_bit_for_bit(case):
    """per-plane sums (and the squares of DoLP / AoLP) come out of the fused launch itself.  They equal
    polcue_channel_stats_f32 of the stored outputs BIT FOR BIT (one canonical summation order, polcue_device.cuh), are
    identical from run to run although tiles are handed out dynamically, and match float64 numpy sums.  The last two cases
    take the internal fallback (Ws = 33: no 4-pixel groups; n = 1.8: steep table) and must give the same."""
    kind, b, h, w, n = case
    mosaic = dev(synth.gen_batch(kind, 17, b, h, w))
    out = ops.fused_mosaic(mosaic, n, want_stats=True)
    plain = ops.fused_mosaic(mosaic, n)
    assert torch.equal(out["xolp"], plain["xolp"]) and torch.equal(out["normals"], plain["normals"])
    xs, ns = ops.stats_from_stats13(out["stats13"])
    assert torch.equal(xs, ops.channel_stats(out["xolp"]))
    assert torch.equal(ns, ops.channel_stats(out["normals"])[:, 0])
    for _ in range(3):
        again = ops.fused_mosaic(mosaic, n, want_stats=True, out={})
        assert torch.equal(again["stats13"], out["stats13"])
    x64, n64 = out["xolp"].double(), out["normals"].double()
    ref = torch.cat((x64.sum(dim=(0, 2, 3)), n64.sum(dim=(0, 2, 3)), (x64 ** 2).sum(dim=(0, 2, 3))))
    scale = torch.cat((x64.abs().sum(dim=(0, 2, 3)), n64.abs().sum(dim=(0, 2, 3)), (x64 ** 2).sum(dim=(0, 2, 3))))
    assert float(((out["stats13"] - ref).abs() / scale).max()) < 2e-7
    st = ops.xolp_statistics(out["xolp"], reduce_over_ranks=False, stats13=out["stats13"])
    assert st == ops.xolp_statistics(out["xolp"], reduce_over_ranks=False)


def test_xolp_statistics_of_fused_output_match_the_reference_script():
    mosaic = synth.gen_batch("P", 0, 3, 128, 192)
    out = ops.fused_mosaic(dev(mosaic), 1.5)
    st = ops.xolp_statistics(out["xolp"])
    rho, phi = [], []
    for b in range(3):
        _, r, p = O.iun_and_xolp_closed(O.stack_quadrants(mosaic[b]))
        rho.append(r)
        phi.append(p)
    ref = O.xolp_statistics(np.stack(rho), np.stack(phi))
    for key, val in ref.items():
        assert abs(st[key] - val) < 2e-6, (key, st[key], val)


def test_all_mask_groups_in_one_launch_match_separate_launches_and_oracle():
    gt, pred, inst, _ = synth.gen_depth_batch(0, 6, 96, 128)
    groups = [None] + list(synth.MATERIAL_LEVELS)
    before = _lib.launch_count()
    sums, metrics = ops.depth_errors_groups(dev(gt), dev(pred), dev(inst), 0.1, 2.0, groups)
    assert _lib.launch_count() == before + 1 and sums.shape == (6, 11, 8) and metrics.shape == (6, 11, 7)
    for gi, level in enumerate(groups):
        s1, m1 = ops.depth_errors_per_image(dev(gt), dev(pred), 0.1, 2.0, dev(inst) if level is not None else None, level)
        assert torch.equal(s1[:, :4], sums[:, gi, :4])                                   # pixel counts: exact
        assert torch.allclose(s1[:, 4:], sums[:, gi, 4:], rtol=1e-6, atol=1e-12)         # float sums: other summation order
        assert torch.allclose(m1, metrics[:, gi], rtol=2e-6, equal_nan=True)
        rows, _ = O.depth_errors_per_image(gt, pred, 0.1, 2.0, inst if level is not None else None, level)
        assert np.allclose(metrics[:, gi].cpu().numpy(), rows, rtol=5e-6, equal_nan=True)
    only_all, _ = ops.depth_errors_groups(dev(gt), dev(pred), None, 0.1, 2.0, [None])
    # "all" is assembled from the per-material partial sums, so a different group set means a different (fixed) order
    assert torch.equal(only_all[:, 0, :4], sums[:, 0, :4]) and torch.allclose(only_all[:, 0, 4:], sums[:, 0, 4:], rtol=1e-6, atol=0)
    again, _ = ops.depth_errors_groups(dev(gt), dev(pred), dev(inst), 0.1, 2.0, groups)
    assert torch.equal(again, sums)                                                      # bitwise reproducible
    again_all, _ = ops.depth_errors_groups(dev(gt), dev(pred), None, 0.1, 2.0, [None])
    assert torch.equal(again_all, only_all)
    odd = ops.depth_errors_groups(dev(gt[:, :95, :127].copy()), dev(pred[:, :95, :127].copy()), dev(inst[:, :95, :127].copy()), 0.1, 2.0,
                                  [40, None, 200])[1].cpu().numpy()                     # scalar path, other group order
    for gi, level in enumerate([40, None, 200]):
        rows, _ = O.depth_errors_per_image(gt[:, :95, :127], pred[:, :95, :127], 0.1, 2.0, inst[:, :95, :127] if level else None, level)
        assert np.allclose(odd[:, gi], rows, rtol=5e-6, equal_nan=True)


def _guarded(shape, dtype, fill):
    """A tensor view with 4 KB canary bands on both sides (compute-sanitizer is closed on this pool)."""
    n = int(np.prod(shape))
    item = torch.empty((), dtype=dtype).element_size()
    pad = 4096 // item
    buf = torch.full((n + 2 * pad,), fill, dtype=dtype, device="cuda")
    return buf, buf[pad:pad + n].view(shape), pad


def _bands_intact(buf, pad, fill):
    return bool((buf[:pad] == fill).all()) and bool((buf[-pad:] == fill).all())


def test_eval_pass_equals_the_separate_entry_points_and_replays_from_a_cuda_graph():
    """polcue_eval_pass_f32 (BASELINE configs[4]): GT normals, every mask group's per-image metrics and the accumulators of
    the mean over images in three launches -- bit-equal to the separate entry points, equal to the oracle's mean over
    images, NaN rows (empty masks) poisoning the mean as np.array(errors).mean(0) does, and capturable in a CUDA graph."""
    gt, pred, inst, k = synth.gen_depth_batch(40, 7, 64, 96)
    inst[3] = 60                                     # image 3 holds a single material: every other group's mask is empty there
    groups = [None] + list(synth.MATERIAL_LEVELS)
    g, q, i, kk = dev(gt), dev(pred), dev(inst), dev(k)
    out = ops.eval_pass(g, q, i, kk, 0.1, 2.0, groups)
    assert torch.equal(out["normals"], ops.depth_to_normals(g[:, None], kk))
    sums, metrics = ops.depth_errors_groups(g, q, i, 0.1, 2.0, groups)
    assert torch.equal(out["sums"], sums) and torch.equal(out["metrics"].nan_to_num(-1.0), metrics.nan_to_num(-1.0))
    acc = out["mean_acc"].cpu().numpy()
    assert acc[0] == 7
    rows = metrics.double().cpu().numpy()                          # [B, G, 7]
    expect = rows.sum(axis=0).reshape(-1)                          # NaN where an image has an empty mask for the group
    assert np.array_equal(np.isnan(acc[1:]), np.isnan(expect))
    assert np.allclose(acc[1:], expect, rtol=1e-12, equal_nan=True)
    for gi, level in enumerate(groups):
        _, mean = O.depth_errors_per_image(gt, pred, 0.1, 2.0, inst if level is not None else None, level)
        got = acc[1 + 7 * gi: 8 + 7 * gi] / acc[0]
        assert np.array_equal(np.isnan(got), np.isnan(mean)) and np.allclose(got, mean, rtol=5e-6, equal_nan=True), level
    # a fixed set of buffers: capture once, replay, same bits
    bufs = ops.eval_pass(g, q, i, kk, 0.1, 2.0, groups)
    graph, side = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            ops.eval_pass(g, q, i, kk, 0.1, 2.0, groups, out=bufs)
    torch.cuda.current_stream().wait_stream(side)
    bufs["mean_acc"].zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert np.array_equal(bufs["mean_acc"].cpu().numpy(), acc, equal_nan=True)
    # no normals requested, bad arguments
    lean = ops.eval_pass(g, q, i, None, 0.1, 2.0, groups, want_normals=False)
    assert lean["normals"] is None and torch.equal(lean["mean_acc"].nan_to_num(-1.0), out["mean_acc"].nan_to_num(-1.0))
    with pytest.raises(ValueError):
        ops.eval_pass(g, q[:, :32], i, kk, 0.1, 2.0, groups)


def test_trig_mode_is_a_property_of_the_table_handle_and_host_buffers_are_pinned():
    """No process-global state changes numerics: `ops.trig` only selects which table handle a call receives (the mode lives in
    the handle, polcue_lut_set_trig), so two threads / two blocks cannot disturb each other; polcue_host_alloc_on hands out
    pinned memory that the host entry points (and torch) treat as such, and rejects bad arguments."""
    import ctypes as C
    mosaic = dev(synth.gen_u_mosaic(3, 64, 96))[None]
    with ops.trig("poly"):
        poly = ops.fused_mosaic(mosaic, 1.5)["normals"].clone()
        with ops.trig("mufu"):
            inner = ops.fused_mosaic(mosaic, 1.5)["normals"].clone()
        assert torch.equal(ops.fused_mosaic(mosaic, 1.5)["normals"], poly)          # restored after the inner block
    mufu = ops.fused_mosaic(mosaic, 1.5)["normals"]
    assert torch.equal(inner, mufu) and not torch.equal(poly, mufu) and float((poly - mufu).abs().max()) < 2e-6
    assert ops.lut_for(1.5, mosaic.device, "poly").value != ops.lut_for(1.5, mosaic.device, "mufu").value
    L = _lib.lib()
    assert L.polcue_lut_set_trig(None, 1) == _lib.EINVAL
    t = ops.host_empty((3, 5, 7), torch.float32)
    assert t.is_pinned() and t.shape == (3, 5, 7) and float(t.abs().sum()) == 0.0
    view = t[1]
    del t
    view.fill_(2.0)                                                                   # the block lives as long as any view of it
    assert float(view.sum()) == 70.0
    ptr = C.c_void_p()
    assert L.polcue_host_alloc_on(C.byref(ptr), 0, -1) == _lib.EINVAL and L.polcue_host_alloc_on(None, 16, -1) == _lib.EINVAL
    assert L.polcue_host_free(C.c_void_p(12345)) == _lib.EINVAL and L.polcue_host_free(None) == 0
    with pytest.raises(ValueError):                                                  # a stale host buffer of another geometry is refused
        ops.fused_mosaic_host(mosaic.cpu(), 1.5, out={"xolp": torch.empty((1, 2, 16, 24))})


@pytest.mark.parametrize("shape", [(2, 2), (6, 10), (34, 66), (66, 128), (130, 250), (128, 192)])
def test_kernels_never_write_outside_their_outputs(shape):
    import ctypes as C
    h, w = shape
    hs, ws = h // 2, w // 2
    L = _lib.lib()
    b = 3
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    mosaic = dev(np.random.default_rng(1).integers(0, 256, (b, h, w), dtype=np.uint8))
    lut = ops.lut_for(1.5, mosaic.device)
    xb, xolp, xp = _guarded((b, 2, hs, ws), torch.float32, -7.0)
    nb, nrm, npad = _guarded((b, 9, hs, ws), torch.float32, -7.0)
    ib, iun, ipad = _guarded((b, hs, ws), torch.float32, -7.0)
    pb, planes, ppad = _guarded((b, 4, hs, ws), torch.uint8, 201)
    assert L.polcue_fused_mosaic_u8(mosaic.data_ptr(), b, h, w, lut, planes.data_ptr(), iun.data_ptr(), xolp.data_ptr(),
                                    nrm.data_ptr(), stream) == 0
    torch.cuda.synchronize()
    assert _bands_intact(xb, xp, -7.0) and _bands_intact(nb, npad, -7.0) and _bands_intact(ib, ipad, -7.0) and _bands_intact(pb, ppad, 201)
    assert not bool((nrm == -7.0).any()) and not bool((xolp == -7.0).any())          # and every element was written
    # get_normals, XOLP from planes, split, stencil and the loss gradient through the same guard
    nb2, nrm2, npad2 = _guarded((b, 9, hs, ws), torch.float32, -7.0)
    assert L.polcue_normals_from_xolp_f32(xolp.data_ptr(), b, hs, ws, lut, nrm2.data_ptr(), stream) == 0
    xb2, xolp2, xp2 = _guarded((b, 2, hs, ws), torch.float32, -7.0)
    pl = [planes[:, k].contiguous() for k in range(4)]
    assert L.polcue_xolp_planes_u8(*(q.data_ptr() for q in pl), b, hs, ws, None, xolp2.data_ptr(), stream) == 0
    quads = [_guarded((b, hs, ws), torch.uint8, 201) for _ in range(4)]
    assert L.polcue_split_pol(mosaic.data_ptr(), b, h, w, 1, *(q[1].data_ptr() for q in quads), stream) == 0
    gt, pred, _, k = synth.gen_depth_batch(0, b, hs, ws)
    d_gt, d_pred, d_k = dev(np.where(gt > 0, gt, 0.7).astype(np.float32)), dev(pred), dev(k)
    sb, sn, spad = _guarded((b, 3, hs, ws), torch.float32, -7.0)
    assert L.polcue_depth_to_normals_f32(d_gt.data_ptr(), d_k.data_ptr(), b, hs, ws, sn.data_ptr(), stream) == 0
    gb, grad, gpad = _guarded((b, 1, hs, ws), torch.float32, -7.0)
    mask = torch.ones_like(d_gt)
    ws_buf = torch.zeros(int(L.polcue_normals_loss_workspace_bytes()), dtype=torch.uint8, device="cuda")
    sums, one = torch.empty(2, dtype=torch.float64, device="cuda"), torch.ones((), device="cuda")
    assert L.polcue_normals_loss_fwd_f32(d_gt.data_ptr(), d_pred.data_ptr(), d_k.data_ptr(), mask.data_ptr(), b, hs, ws,
                                         ws_buf.data_ptr(), sums.data_ptr(), None, stream) == 0
    assert L.polcue_normals_loss_bwd_f32(d_gt.data_ptr(), d_pred.data_ptr(), d_k.data_ptr(), mask.data_ptr(), b, hs, ws,
                                         sums.data_ptr(), one.data_ptr(), grad.data_ptr(), stream) == 0
    torch.cuda.synchronize()
    assert _bands_intact(nb2, npad2, -7.0) and _bands_intact(xb2, xp2, -7.0) and _bands_intact(sb, spad, -7.0)
    assert _bands_intact(gb, gpad, -7.0) and all(_bands_intact(q[0], q[2], 201) for q in quads)
    assert torch.allclose(nrm2, nrm, atol=3e-6) and torch.equal(xolp2, xolp)   # sincos(phi) vs the algebraic half angle
    for got, ref in zip((q[1] for q in quads), O.split_pol(mosaic[0].cpu().numpy())):
        assert np.array_equal(got[0].cpu().numpy(), ref)
    assert not bool((sn == -7.0).any()) and not bool((grad == -7.0).any())


def test_cuda_graph_capture_and_replay_of_the_loader_path():
    """The launch-latency-bound training-resolution path (two kernels) captured in a CUDA graph and replayed."""
    planes_a = [dev(p)[None].repeat(4, 1, 1) for p in synth.gen_p_planes(1, 64, 96)]
    planes_b = [dev(p)[None].repeat(4, 1, 1) for p in synth.gen_p_planes(2, 64, 96)]
    static = [p.clone() for p in planes_a]
    mosaic_static = dev(synth.gen_batch("P", 0, 2, 64, 96))
    out = {"xolp": torch.empty((2, 2, 32, 48), device="cuda"), "normals": torch.empty((2, 9, 32, 48), device="cuda")}
    ops.lut_for(1.5, static[0].device)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):                         # warm-up on the side stream, as torch's graph recipe asks
        ops.get_normals(ops.xolp_from_planes(*static)[1], 1.5)
        ops.fused_mosaic(mosaic_static, 1.5, out=out)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        _, xolp = ops.xolp_from_planes(*static)
        normals = ops.get_normals(xolp, 1.5)
        ops.fused_mosaic(mosaic_static, 1.5, out=out)
    for src in (planes_b, planes_a):
        for dst, s_ in zip(static, src):
            dst.copy_(s_)
        mosaic_static.copy_(dev(synth.gen_batch("P", 5 if src is planes_b else 9, 2, 64, 96)))
        graph.replay()
        torch.cuda.synchronize()
        ref = ops.get_normals(ops.xolp_from_planes(*src)[1], 1.5)
        assert torch.equal(normals, ref)
        assert torch.equal(out["normals"], ops.fused_mosaic(mosaic_static, 1.5)["normals"])


def test_launch_counter_counts_kernels():
    before = _lib.launch_count()
    ops.fused_mosaic(dev(synth.gen_u_mosaic(0, 32, 48))[None], 1.5)
    assert _lib.launch_count() == before + 1


# ------------------------------------------------------------------------------------------------
# loader front end: Pillow-exact Lanczos resize, then XOLP + normals
# ------------------------------------------------------------------------------------------------
def test_lanczos_resize_bit_exact_vs_reference_outputs(resize_golden):
    g = resize_golden
    for i, (ih, iw, oh, ow) in enumerate(g["case_shapes"]):
        img = g[f"in_{i}"]
        got = ops.lanczos_resize(dev(img), (int(oh), int(ow)))
        assert np.array_equal(got.cpu().numpy(), g[f"out_{i}"]), (i, ih, iw, oh, ow)
        both = ops.lanczos_resize(dev(np.stack((img, img))), (int(oh), int(ow)), flip=[False, True]).cpu().numpy()
        assert np.array_equal(both[0], g[f"out_{i}"]) and np.array_equal(both[1], g[f"outflip_{i}"]), i


def test_lanczos_resize_hammer_geometry_vs_reference_digest(resize_golden):
    import hashlib
    planes = synth.gen_p_planes(4242, 832, 1088)
    small = ops.lanczos_resize(dev(np.stack(planes)), (320, 480), flip=[False, True, False, True]).cpu().numpy()
    for k in range(4):
        assert np.array_equal(small[k][::8, ::8], resize_golden[f"hammer_sample_{k}"])
        assert hashlib.sha256(small[k].tobytes()).hexdigest() == resize_golden["hammer_sha256"][k]


@pytest.fixture
def byte_form(request):
    _lib.lib().polcue_debug_resize_force_bytes(1 if request.param else 0)
    yield request.param
    _lib.lib().polcue_debug_resize_force_bytes(0)


@pytest.mark.parametrize("byte_form", [False, True], indirect=True)
def test_lanczos_resize_random_geometries_vs_oracle(byte_form):
    """Every tap-count class (<= 16, <= 32, generic), upscaling, identity axes, odd widths (scalar paths), saturating inputs;
    the horizontal pass in its dp4a form and in its byte-load form."""
    rng = np.random.default_rng(11)
    shapes = [(40, 52, 17, 23), (90, 130, 9, 10), (256, 300, 8, 9), (12, 10, 30, 41), (31, 64, 31, 16), (64, 31, 16, 31), (1, 1, 1, 1),
              (3, 200, 3, 7), (200, 3, 7, 3), (17, 16, 16, 16),
              # bulk-copy (TMA) tiles whose rows are not 16-byte multiples: word-aligned rows, unaligned rows, 17..32 taps, short last tile
              (64, 72, 20, 30), (64, 70, 20, 30), (96, 1224, 40, 480), (70, 72, 20, 30), (70, 70, 20, 30), (48, 1248, 19, 640), (40, 2000, 11, 992),
              (40, 2000, 11, 1000)]
    shapes += [tuple(int(v) for v in rng.integers(1, 140, 4)) for _ in range(12)]
    for ih, iw, oh, ow in shapes:
        imgs = rng.integers(0, 256, (3, ih, iw), dtype=np.uint8)
        imgs[1] = rng.choice([0, 255], (ih, iw))
        flips = [False, True, True]
        got = ops.lanczos_resize(dev(imgs), (oh, ow), flip=flips).cpu().numpy()
        for k in range(3):
            src = np.ascontiguousarray(imgs[k][:, ::-1]) if flips[k] else imgs[k]
            assert np.array_equal(got[k], O.resize_lanczos_u8(src, (oh, ow))), (ih, iw, oh, ow, k)


@pytest.mark.parametrize("geometry", [(104, 136, 40, 60), (61, 83, 23, 31), (208, 272, 80, 120)])
def test_loader_front_end_vs_oracle(geometry):
    ih, iw, oh, ow = geometry
    b = 3
    planes = [np.stack([synth.gen_p_planes(50 + f, ih, iw)[k] for f in range(b)]) for k in range(4)]
    flips = [False, True, False]
    out = ops.loader_front_end(*(dev(p) for p in planes), (oh, ow), n=1.5, flip=flips, want_iun=True, normalize_xolp=ops.XOLP_MEAN_STD)
    # ShallowEncoder.normalizeInput(x, 'XOLP') (pre_encoders.py:79) on the float32 tensor, bit for bit
    assert torch.equal(out["xolp_norm"], (out["xolp"] - 0.08693199701957657) / 0.44430732785457433)
    assert out["planes"].shape == (b, 4, oh, ow) and out["xolp"].shape == (b, 2, oh, ow) and out["normals"].shape == (b, 9, oh, ow)
    for f in range(b):
        small, xolp, normals = O.loader_front_end(*(p[f] for p in planes), (oh, ow), n=1.5, flip=flips[f])
        assert np.array_equal(out["planes"][f].cpu().numpy(), small), "resized planes are not bit-exact"
        P.assert_dolp_close(out["xolp"][f, 0].cpu().numpy(), xolp[0])
        P.assert_aolp_close(out["xolp"][f, 1].cpu().numpy(), xolp[1])
        P.assert_normals_close(out["normals"][f].cpu().numpy().reshape(3, 3, oh, ow), normals.reshape(3, 3, oh, ow), axis=1)
    # the same numbers as the two-step path through the public pieces
    small = ops.lanczos_resize(dev(np.stack(planes, axis=1)), (oh, ow), flip=np.repeat(flips, 4).tolist())
    assert torch.equal(small, out["planes"])
    two = ops.fused_planes(*(small[:, k].contiguous() for k in range(4)), want_iun=True)
    for key in ("xolp", "normals", "iun"):
        assert torch.equal(two[key], out[key]), key


def test_indoor_dataset_mirrors(resize_golden):
    """ResizePol and get_xolp used exactly as indoor_dataset.py:335-349, :354 use the originals."""
    from PIL import Image
    from polcue.compat import indoor_dataset as c_ds
    g = resize_golden
    ih, iw, oh, ow = (int(v) for v in g["case_shapes"][0])
    resize_pol = c_ds.ResizePol((oh, ow))
    out = resize_pol(Image.fromarray(g["in_0"]).convert("L"))
    assert out.mode == "L" and out.size == (ow, oh) and np.array_equal(np.asarray(out), g["out_0"])
    assert np.array_equal(resize_pol(g["in_0"]), g["out_0"])
    with pytest.raises(ValueError):
        resize_pol(Image.fromarray(g["in_0"]).convert("RGB"))
    planes = synth.gen_p_planes(8, 2 * oh, 2 * ow)
    inputs = {}
    for name, k in (("pol00", 0), ("pol01", 1), ("pol10", 2), ("pol11", 3)):       # 0, 45, 90, 135 deg
        inputs[(name + "_gray", 0, 0)] = resize_pol(Image.fromarray(planes[k], "L"))
    c_ds.get_xolp(None, inputs, 0)
    xolp = inputs[("xolp", 0, 0)]
    assert xolp.shape == (2, oh, ow) and xolp.dtype == torch.float64 and not xolp.is_cuda
    small = np.stack([np.asarray(inputs[(name + "_gray", 0, 0)]) for name in ("pol00", "pol01", "pol10", "pol11")], axis=2)
    _, rho, phi = O.iun_and_xolp_closed(small)
    P.assert_dolp_close(xolp[0].numpy(), rho)
    P.assert_aolp_close(xolp[1].numpy(), phi)
    # batched form, loader argument order (pol00, pol10, pol01, pol11)
    full = [dev(p[None]) for p in planes]
    both = c_ds.polarization_inputs(full[0], full[2], full[1], full[3], (oh, ow))
    assert np.array_equal(both["planes"][0].cpu().numpy(), small.transpose(2, 0, 1))
    assert torch.allclose(both["xolp"][0].cpu().double(), xolp, atol=0, rtol=0)


def test_loader_front_end_host_entry_point_matches_device_entry_point():
    rng = np.random.default_rng(12)
    b, ih, iw, oh, ow = 11, 104, 144, 40, 60                          # 11 samples in chunks of 4: 3 ring slots + a short last chunk
    full = [torch.from_numpy(rng.integers(0, 256, (b, ih, iw), dtype=np.uint8)) for _ in range(4)]
    flips = [bool(v) for v in rng.integers(0, 2, b)]
    host = ops.loader_front_end_host(*(t.pin_memory() for t in full), (oh, ow), flip=flips, normalize_xolp=ops.XOLP_MEAN_STD, chunk_samples=4)
    devo = ops.loader_front_end(*(t.cuda() for t in full), (oh, ow), flip=flips, normalize_xolp=ops.XOLP_MEAN_STD)
    for key in ("planes", "xolp", "xolp_norm", "normals"):
        assert not host[key].is_cuda and torch.equal(host[key], devo[key].cpu()), key
    again = ops.loader_front_end_host(*full, (oh, ow), flip=flips, want_planes=False, want_normals=False)      # pageable, cached ring
    assert set(again) == {"xolp"} and torch.equal(again["xolp"], host["xolp"])
    with pytest.raises(TypeError):
        ops.loader_front_end_host(*(t.cuda() for t in full), (oh, ow))


def test_loader_front_end_rejects_bad_arguments():
    a = torch.zeros((2, 8, 8), dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):
        ops.loader_front_end(a, a, a, a[:1], (4, 4))
    with pytest.raises(ValueError):
        ops.loader_front_end(a, a, a, a, (4, 4), flip=[True])
    with pytest.raises(TypeError):
        ops.lanczos_resize(a.cpu(), (4, 4))
    with pytest.raises(_lib.PolcueError):
        ops.lanczos_resize(a, (0, 4))


@pytest.mark.parametrize("geometry", [(104, 144, 40, 60), (61, 83, 23, 31), (96, 16, 8, 16), (12, 10, 30, 41), (90, 130, 9, 10)])
def test_resize_kernels_never_write_outside_their_outputs(geometry):
    """Canary bands around the workspace, the resized planes and every float output of the front end (TMA-pipelined,
    manually staged, identity-axis, upscaling and generic-tap-count geometries), through the C ABI."""
    import ctypes as C
    ih, iw, oh, ow = geometry
    L = _lib.lib()
    b = 3
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rng = np.random.default_rng(4)
    full = [dev(rng.integers(0, 256, (b, ih, iw), dtype=np.uint8)) for _ in range(4)]
    flips = dev(np.array([0, 1, 0], np.uint8))
    plan = ops.resize_plan((ih, iw), (oh, ow), full[0].device)
    lut = ops.lut_for(1.5, full[0].device)
    need = int(L.polcue_resize_workspace_bytes(plan, 4 * b))
    wb, wsp, wpad = _guarded((need,), torch.uint8, 201)
    pb, planes, ppad = _guarded((b, 4, oh, ow), torch.uint8, 201)
    xb, xolp, xp = _guarded((b, 2, oh, ow), torch.float32, -7.0)
    qb, xnorm, qp = _guarded((b, 2, oh, ow), torch.float32, -7.0)
    nb, nrm, npad = _guarded((b, 9, oh, ow), torch.float32, -7.0)
    ib, iun, ipad = _guarded((b, oh, ow), torch.float32, -7.0)
    mean_std = (C.c_float * 2)(*ops.XOLP_MEAN_STD)
    assert L.polcue_loader_front_end_u8(plan, *(f.data_ptr() for f in full), b, flips.data_ptr(), lut, wsp.data_ptr(), planes.data_ptr(),
                                        iun.data_ptr(), xolp.data_ptr(), nrm.data_ptr(), mean_std, xnorm.data_ptr(), stream) == 0
    torch.cuda.synchronize()
    assert _bands_intact(wb, wpad, 201) and _bands_intact(pb, ppad, 201) and _bands_intact(xb, xp, -7.0) and _bands_intact(qb, qp, -7.0)
    assert _bands_intact(nb, npad, -7.0) and _bands_intact(ib, ipad, -7.0)
    assert not bool((nrm == -7.0).any()) and not bool((xolp == -7.0).any()) and not bool((xnorm == -7.0).any())
    for f in range(b):
        small, _, _ = O.loader_front_end(*(p[f].cpu().numpy() for p in full), (oh, ow), flip=bool(flips[f]))
        assert np.array_equal(planes[f].cpu().numpy(), small)
    # single-tensor entry point with its own guards
    ob, outp, opad = _guarded((b, oh, ow), torch.uint8, 201)
    wb2, wsp2, wpad2 = _guarded((int(L.polcue_resize_workspace_bytes(plan, b)),), torch.uint8, 201)
    assert L.polcue_resize_lanczos_u8(plan, full[1].data_ptr(), b, None, wsp2.data_ptr(), outp.data_ptr(), stream) == 0
    torch.cuda.synchronize()
    assert _bands_intact(ob, opad, 201) and _bands_intact(wb2, wpad2, 201)
    assert np.array_equal(outp[1].cpu().numpy(), O.resize_lanczos_u8(full[1][1].cpu().numpy(), (oh, ow)))
    # rejected arguments launch nothing
    assert L.polcue_resize_lanczos_u8(plan, None, b, None, wsp2.data_ptr(), outp.data_ptr(), stream) == -22
    assert L.polcue_loader_front_end_u8(plan, *(f.data_ptr() for f in full), b, None, lut, wsp.data_ptr(), planes.data_ptr(), None,
                                        xolp.data_ptr(), nrm.data_ptr(), None, xnorm.data_ptr(), stream) == -22
