"""CPU: the oracle restatements against outputs of the reference itself (tests/golden)."""
import os

import numpy as np
import pytest

from oracle import polcue_oracle as O
import parity as P

ANG = O.CANONICAL_ANGLES


def test_split_pol_bit_exact(golden):
    for tag in ("gray", "bgr"):
        q = O.split_pol(golden[f"split_{tag}_in"])
        for name, arr in zip(("im00", "im10", "im01", "im11"), q):
            assert np.array_equal(arr, golden[f"split_{tag}_{name}"])
    with pytest.raises(ValueError):
        O.split_pol(np.zeros((5, 4), np.uint8))


def test_xolp_lstsq_restatement_matches_reference_exactly(golden):
    for tag in ("u", "p"):
        iun, rho, phi = O.iun_and_xolp_lstsq(golden[f"xolp_{tag}_in"], ANG)
        # same LAPACK call on the same bytes: agree to round-off, ties may still flip phi by pi
        assert np.allclose(iun, golden[f"xolp_{tag}_iun"], rtol=1e-13, atol=1e-12)
        assert np.allclose(rho, golden[f"xolp_{tag}_rho"], rtol=1e-12, atol=1e-14)
        _, degenerate = O.xolp_tie_masks(golden[f"xolp_{tag}_in"])
        P.assert_aolp_close(phi, golden[f"xolp_{tag}_phi"], exclude=degenerate, tol=1e-12)


def test_xolp_closed_form_vs_reference(golden):
    for tag in ("u", "p"):
        st = golden[f"xolp_{tag}_in"]
        iun, rho, phi = O.iun_and_xolp_closed(st)
        sign_tie, degenerate = O.xolp_tie_masks(st)
        P.assert_dolp_close(iun, golden[f"xolp_{tag}_iun"], "Iun")
        P.assert_dolp_close(rho, golden[f"xolp_{tag}_rho"], "rho")
        P.assert_aolp_close(phi, golden[f"xolp_{tag}_phi"], exclude=degenerate, tol=1e-12)
        # away from ties the closed form equals lstsq to ~1e-14 without the mod-pi wrap
        clean = ~(sign_tie | degenerate)
        assert np.abs(phi - golden[f"xolp_{tag}_phi"])[clean].max() < 1e-12


def test_xolp_known_answers(golden):
    iun, rho, phi = O.iun_and_xolp_closed(golden["kat_in"][None])
    sign_tie, degenerate = O.xolp_tie_masks(golden["kat_in"][None])
    P.assert_dolp_close(iun[0], golden["kat_iun"])
    P.assert_dolp_close(rho[0], golden["kat_rho"])
    P.assert_aolp_close(phi[0], golden["kat_phi"], exclude=degenerate[0], tol=1e-12)
    # SURVEY 4 table
    table = {(10, 20, 30, 40): (25.0, 0.565685424949238, -1.1780972450961722),
             (0, 255, 0, 0): (63.75, 2.0, 0.7853981633974484),
             (0, 0, 255, 0): (63.75, 2.0, np.pi / 2),
             (100, 100, 0, 0): (50.0, 1.4142135623730951, 0.39269908169872436),
             (0, 0, 0, 0): (0.0, 0.0, 0.0)}
    for px, (a, b, c) in table.items():
        i, r, p = O.iun_and_xolp_closed(np.array(px, np.uint8).reshape(1, 1, 4))
        assert abs(i[0, 0] - a) < 1e-12 and abs(r[0, 0] - b) < 1e-12 and abs(p[0, 0] - c) < 1e-12


def test_xolp_noncanonical_angles_and_float_input(golden):
    st = golden["xolp_p_in"]
    iun, rho, phi = O.iun_and_xolp_closed(st, golden["xolp_ang2"])
    assert np.allclose(iun, golden["xolp_ang2_iun"], rtol=1e-12)
    assert np.allclose(rho, golden["xolp_ang2_rho"], rtol=1e-11, atol=1e-13)
    P.assert_aolp_close(phi, golden["xolp_ang2_phi"], tol=1e-10)
    iun, rho, phi = O.iun_and_xolp_closed(golden["xolp_f32_in"])
    _, degenerate = O.xolp_tie_masks(golden["xolp_f32_in"])  # integer view only flags exact ties
    P.assert_dolp_close(rho, golden["xolp_f32_rho"])
    s1 = golden["xolp_f32_in"][..., 0] - golden["xolp_f32_in"][..., 2]
    s2 = golden["xolp_f32_in"][..., 1] - golden["xolp_f32_in"][..., 3]
    P.assert_aolp_close(phi, golden["xolp_f32_phi"], exclude=(s1 == 0) & (s2 == 0), tol=1e-10)


@pytest.mark.parametrize("n", [1.3, 1.5, 1.8])
def test_table_inversions_match_scipy_reference(golden, n):
    rq = golden["tab_rho"]
    assert np.allclose(O.rho_diffuse(rq, n), golden[f"tab_d_{n}"], rtol=1e-13, atol=1e-13)
    t1, t2 = O.rho_spec(rq, n)
    assert np.allclose(t1, golden[f"tab_s1_{n}"], rtol=1e-12, atol=1e-12)
    assert np.allclose(t2, golden[f"tab_s2_{n}"], rtol=1e-12, atol=1e-12)


def test_table_anchors_survey():
    anchors = {0.0: (0.0, 0.0, 1.570796327), 0.01: (0.410434783, 0.086477641, 1.566324189),
               0.3: (1.472298774, 0.460030611, 1.436485518), 0.5: (1.692111847, 0.590509261, 1.345596980),
               1.0: (2.217812434, 0.981710534, 0.982727665), 1.2: (2.428092719, 9.416327864, -27.260250003),
               2.0: (3.269213607, 43.154787130, -140.232127007)}
    for rho, (d, s1, s2) in anchors.items():
        q = np.array([rho], np.float32)          # the survey fed float32 tensors (trainer.py:510 `.float()`)
        assert abs(O.rho_diffuse(q, 1.5)[0] - d) < 1e-8
        a, b = O.rho_spec(q, 1.5)
        assert abs(a[0] - s1) < 1e-8 and abs(b[0] - s2) < 1e-8
    _, _, _, imax = O.fresnel_tables(1.5)
    assert imax == 625


def test_restatement_equals_scipy_interp1d():
    scipy_interp = pytest.importorskip("scipy.interpolate")
    rq = np.random.default_rng(5).uniform(-0.1, 2.1, 5000)
    for name, (xk, yk) in O.sorted_knots(1.5).items():
        f = scipy_interp.interp1d(xk, yk, fill_value="extrapolate")
        assert np.array_equal(f(rq), O.interp_linear_extrap(xk, yk, rq)), name


def test_get_normals_vs_reference(golden):
    for tag in ("u", "p"):
        got = O.get_normals(golden[f"getn_{tag}_x32"], 1.5)
        ref = golden[f"getn_{tag}_n"]
        assert np.abs(got - ref).max() < 5e-7        # torch f32 cos/sin vs numpy f32 cos/sin: 1 ulp
        P.assert_normals_close(got.reshape(1, 3, 3, *got.shape[2:]), ref.reshape(1, 3, 3, *ref.shape[2:]), axis=2)
    # SURVEY 4 anchor pixel
    ref0 = [0.5870980739, 0.8079686681, 0.0500249307, -0.3794740927, 0.2757389967, 0.8831576455,
            -0.7996517416, 0.5810546047, 0.1514353819]
    assert np.abs(O.get_normals(golden["getn_u_x32"], 1.5)[0, :, 0, 0] - ref0).max() < 1e-6


def test_float64_chain_vs_reference(golden):
    for tag in ("u", "p"):
        st = golden[f"xolp_{tag}_in"]
        _, rho, phi = O.iun_and_xolp_closed(st)
        sign_tie, degenerate = O.xolp_tie_masks(st)
        th_d = O.rho_diffuse(rho, 1.5)
        th_1, th_2 = O.rho_spec(rho, 1.5)
        P.assert_normals_close(O.calc_normals_hw3(phi, th_d), golden[f"chain_{tag}_nd"], axis=2,
                               twin_ok=sign_tie, skip=degenerate & (rho > 0), tol=1e-9)
        P.assert_normals_close(O.calc_normals_hw3(phi + np.pi / 2, th_1), golden[f"chain_{tag}_n1"], axis=2,
                               twin_ok=sign_tie, tol=1e-9)
        n2 = O.calc_normals_hw3(phi + np.pi / 2, th_2)
        P.assert_normals_close(n2, golden[f"chain_{tag}_n2"], axis=2, twin_ok=sign_tie, skip=degenerate, tol=1e-9)
        assert np.abs(np.abs(n2[..., 2]) - np.abs(golden[f"chain_{tag}_n2"][..., 2]))[degenerate].max(initial=0) < 1e-9


def test_ppp_channel_variant_vs_reference(golden):
    rho, phi, iun = O.polarisation_image_channel(golden["ppp_images"], ANG, golden["ppp_mask"])
    assert np.array_equal(np.isnan(rho), np.isnan(golden["ppp_rho"]))
    assert np.allclose(rho, golden["ppp_rho"], rtol=1e-14, atol=0, equal_nan=True)
    assert np.allclose(phi, golden["ppp_phi"], rtol=1e-14, atol=1e-15)
    assert np.array_equal(iun, golden["ppp_iun"])
    th_d = O.rho_diffuse(rho, 1.5)
    th_1, th_2 = O.rho_spec(rho, 1.5)
    for key, ph, th in (("ppp_nd", phi, th_d), ("ppp_n1", phi + np.pi / 2, th_1), ("ppp_n2", phi + np.pi / 2, th_2)):
        assert np.allclose(O.calc_normals_channel(ph, th, golden["ppp_mask"]), golden[key], rtol=1e-12, atol=1e-12,
                           equal_nan=True)


def test_depth_errors_vs_reference(golden):
    got = np.array(O.compute_depth_errors(golden["met_gt"], golden["met_pred"]))
    assert np.allclose(got, golden["met_torch"], rtol=2e-6)
    assert np.allclose(got, golden["met_numpy"], rtol=2e-6)
    assert np.allclose(got, [0.048497424, 0.003272478, 0.065363288, 0.056583799, 1, 1, 1], rtol=2e-6)
    sums = O.depth_error_sums(golden["met_gt"], golden["met_pred"])
    n = sums[0]
    rebuilt = [sums[6] / n, sums[7] / n, np.sqrt(sums[4] / n), np.sqrt(sums[5] / n), sums[1] / n, sums[2] / n, sums[3] / n]
    assert np.allclose(rebuilt, got, rtol=1e-13)


def test_depth_errors_per_image_vs_reference(golden):
    from polcue import synth
    gt, pred, inst, _ = synth.gen_depth_batch(0, 3, 64, 96)
    rows, mean = O.depth_errors_per_image(gt, pred, 0.1, 2.0)
    assert np.allclose(rows, golden["met_img_rows"], rtol=3e-6)
    assert np.allclose(mean, golden["met_img_rows"].mean(0), rtol=3e-6)
    rows_m, _ = O.depth_errors_per_image(gt, pred, 0.1, 2.0, inst, 40)
    assert rows_m.shape == (3, 7)
    assert all(np.isnan(v) for v in O.compute_depth_errors(np.zeros(0), np.zeros(0)))


# ------------------------------------------------------------------------------------------------
# loader front end: Pillow's 8-bit Lanczos resize
# ------------------------------------------------------------------------------------------------
def test_lanczos_resize_restatement_is_bit_exact_with_the_reference_call_chain(resize_golden):
    import hashlib
    from polcue import synth
    g = resize_golden
    for i, (ih, iw, oh, ow) in enumerate(g["case_shapes"]):
        img = g[f"in_{i}"]
        assert img.shape == (ih, iw)
        assert np.array_equal(O.resize_lanczos_u8(img, (oh, ow)), g[f"out_{i}"]), i
        assert np.array_equal(O.resize_lanczos_u8(np.ascontiguousarray(img[:, ::-1]), (oh, ow)), g[f"outflip_{i}"]), i
    planes = synth.gen_p_planes(4242, 832, 1088)       # HAMMER geometry, inputs regenerated from the seed
    for k in range(4):
        assert hashlib.sha256(planes[k].tobytes()).hexdigest() == g["hammer_in_sha256"][k]
        src = np.ascontiguousarray(planes[k][:, ::-1]) if k & 1 else planes[k]
        small = O.resize_lanczos_u8(src, (320, 480))
        assert np.array_equal(small[::8, ::8], g[f"hammer_sample_{k}"])
        assert hashlib.sha256(small.tobytes()).hexdigest() == g["hammer_sha256"][k]


def test_lanczos_resize_restatement_vs_installed_pillow():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(5)
    for _ in range(25):
        ih, iw, oh, ow = (int(v) for v in rng.integers(1, 120, 4))
        img = rng.integers(0, 256, (ih, iw), dtype=np.uint8)
        if rng.random() < 0.3:
            img[:] = rng.choice([0, 255], img.shape)      # ringing: exercises both clamps of clip8
        ref = np.asarray(Image.fromarray(img, "L").resize((ow, oh), Image.LANCZOS))
        assert np.array_equal(O.resize_lanczos_u8(img, (oh, ow)), ref), (ih, iw, oh, ow)


def test_loader_front_end_oracle_composition():
    from polcue import synth
    planes = synth.gen_p_planes(3, 52, 68)
    small, xolp, normals = O.loader_front_end(planes[0], planes[1], planes[2], planes[3], (20, 30), flip=True)
    assert small.shape == (4, 20, 30) and xolp.shape == (2, 20, 30) and normals.shape == (9, 20, 30)
    assert np.array_equal(small[2], O.resize_lanczos_u8(np.ascontiguousarray(planes[2][:, ::-1]), (20, 30)))
    assert np.allclose(np.linalg.norm(normals.reshape(3, 3, 20, 30), axis=1), 1.0)


KORNIA_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kornia_outputs.npz")


@pytest.mark.skipif(not os.path.exists(KORNIA_GOLDEN), reason="tests/golden/kornia_outputs.npz absent: run tools/pin_kornia.py where kornia 0.5.11 is installed")
def test_depth_to_normals_oracle_against_kornia_outputs():
    """Pins a10 (kornia.geometry.depth.depth_to_normals, trainer.py:1305-1306) on outputs of kornia itself when the fixture
    exists: float64 oracle == kornia float64 to 1e-12 wherever the cross product does not vanish, zero vectors where kornia
    gives zero vectors."""
    data = np.load(KORNIA_GOLDEN)
    names = sorted({k.split("/")[0] for k in data.files if "/" in k})
    assert names
    for name in names:
        depth, km, ref = data[name + "/depth"], data[name + "/K"], data[name + "/normals_f64"]
        got = O.depth_to_normals(depth, km)
        _, bound, dead = O.depth_to_normals_conditioning(depth, km)
        well = np.isfinite(bound) & (bound < 1e6) & ~dead
        assert np.abs(got - ref).transpose(0, 2, 3, 1)[well].max(initial=0.0) < 1e-12 * max(1.0, float(bound[well].max(initial=1.0))), name
        assert (ref.transpose(0, 2, 3, 1)[dead] == 0).all() and (got.transpose(0, 2, 3, 1)[dead] == 0).all(), name


def test_six_functional_form_of_the_normals_loss_equals_the_reference_formulation():
    """The loss kernels evaluate the stencil through six window functionals (G, V, A, B, Cu, Cv; csrc/normals_loss.cu) instead of
    the reference's gradients of (fx Z, fy Z, Z).  The float64 restatement of that form -- forward, adjoints, and the gather
    with the replicate-padding folds, tools/probes/bwd_terms_proto.py -- must agree with float64 autograd of the reference
    formulation (oracle.normals_loss_torch, manydepth/trainer.py:1298-1309) to rounding on every shape, borders and 2-pixel
    images included."""
    import importlib.util
    import sys
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "probes", "bwd_terms_proto.py")
    spec = importlib.util.spec_from_file_location("bwd_terms_proto", path)
    proto = importlib.util.module_from_spec(spec)
    sys.modules["bwd_terms_proto"] = proto
    spec.loader.exec_module(proto)
    for seed, (h, w) in enumerate(((7, 9), (2, 2), (12, 16), (3, 4), (2, 9), (9, 2), (5, 33))):
        d_loss, d_grad = proto.compare_with_autograd(h, w, seed)
        assert d_loss < 1e-13 and d_grad < 1e-12, ((h, w), d_loss, d_grad)
