"""CPU-only checks: the C ABI surface, the host-side table construction, argument validation that needs no GPU,
and the world_size-2 sharding / reduction logic over gloo."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import polcue_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from polcue import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.lib()


def test_library_exports_every_declared_symbol(L):
    header = open(os.path.join(ROOT, "include", "polcue.h")).read()
    declared = set(re.findall(r"POLCUE_API[^;(]*?\b(polcue_\w+)\s*\(", header))
    assert len(declared) >= 28
    from polcue import _lib
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.polcue_version()
    assert L.polcue_error_string(-22).startswith(b"invalid argument")


def test_header_is_plain_c_and_a_c_program_can_bind_the_library(tmp_path):
    """include/polcue.h must compile as C (no C++-isms, no CUDA or torch types) and a C program must link against it."""
    from polcue import _lib
    src = tmp_path / "abi.c"
    src.write_text(
        '#include "polcue.h"\n#include <stdio.h>\n'
        "int main(void) {\n"
        "  polcue_lut* lut = 0;\n"
        "  int rc = polcue_lut_host_build(1.5, &lut);\n"
        "  float rho[3] = {0.0f, 0.3f, 2.0f}, theta[3];\n"
        "  if (rc != POLCUE_OK) return 1;\n"
        "  if (polcue_lut_eval_host(lut, 0, rho, 3, theta) != POLCUE_OK) return 2;\n"
        '  printf("%s %d %.6f %.6f\\n", polcue_version(), polcue_lut_cells(lut, 0), theta[1], theta[2]);\n'
        "  if (polcue_fused_mosaic_u8(0, 1, 4, 4, lut, 0, 0, 0, 0, 0) != POLCUE_EINVAL) return 3;\n"
        "  polcue_lut_destroy(lut);\n  return 0;\n}\n")
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    build = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                            "-L", libdir, "-l:libpolcue.so", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0, (run.returncode, run.stdout, run.stderr)
    fields = run.stdout.split()
    assert abs(float(fields[-2]) - 1.472298774) < 1e-5 and abs(float(fields[-1]) - 3.269213607) < 1e-5


def test_library_is_built_for_sm_100a_only():
    from polcue import _lib
    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def _host_lut(L, n):
    h = C.c_void_p()
    rc = L.polcue_lut_host_build(float(n), C.byref(h))
    return rc, h


@pytest.mark.parametrize("n", [1.3, 1.33, 1.5, 1.8, 2.5])
def test_cell_tables_reproduce_the_reference_interpolant(L, n):
    rc, h = _host_lut(L, n)
    assert rc == 0
    knots = O.sorted_knots(n)
    rng = np.random.default_rng(int(n * 100))
    queries = np.concatenate((rng.uniform(0, 1, 200_000) ** 2, rng.uniform(0, 2.2, 50_000), rng.uniform(-0.2, 0, 1000),
                              [0.0, 1.0, 0.5, 0.49999997, 2.0])).astype(np.float32)
    for t, name in enumerate(("diffuse", "spec1", "spec2")):
        xk, yk = knots[name]
        # same knots as the oracle / scipy hold (libm vs numpy sin/cos: 1 ulp)
        count = L.polcue_lut_knots(h, t, None, None, 0)
        assert count == len(xk)
        x, y = np.empty(count), np.empty(count)
        assert L.polcue_lut_knots(h, t, x.ctypes.data, y.ctypes.data, count) == count
        assert np.allclose(x, xk, rtol=1e-12, atol=1e-30) and np.allclose(y, yk, rtol=1e-15)
        # knot positions themselves (float32-rounded) and random queries
        q = np.concatenate((queries, xk.astype(np.float32), (0.5 * (xk[1:] + xk[:-1])).astype(np.float32)))
        theta = np.empty(q.size, np.float32)
        assert L.polcue_lut_eval_host(h, t, q.ctypes.data, q.size, theta.ctypes.data) == 0
        q64 = q.astype(np.float64)
        ref = O.interp_linear_extrap(xk, yk, q64)
        hi = np.clip(np.searchsorted(xk, q64), 1, len(xk) - 1)
        slope = np.abs((yk[hi] - yk[hi - 1]) / (xk[hi] - xk[hi - 1]))
        tol = 4e-7 * (1 + np.abs(ref)) + slope * 1.3e-7 * np.maximum(np.abs(q64), np.abs(1 - q64))
        bad = np.abs(theta - ref) > tol
        assert not bad.any(), (name, q[bad][:5], theta[bad][:5], ref[bad][:5])
    cells = [L.polcue_lut_cells(h, t) for t in range(3)]
    assert all(c > 0 for c in cells) and sum(cells) * 16 <= 200 * 1024
    L.polcue_lut_destroy(h)


def test_steep_end_segments_are_evaluated_in_float64(L):
    """n = 1.8: the knot next to the specular peak lies 1.5e-7 below it, so the last segment of the second branch has
    slope -10797 and theta reaches -1e4 rad at rho = 2; float32 arithmetic is 2e-3 rad off there (found by the
    randomized GPU sweep).  The library flags such tables and evaluates them in float64."""
    rc, h = _host_lut(L, 1.8)
    xys = np.empty(3)
    assert [L.polcue_lut_steep(h, t, xys.ctypes.data) for t in (0, 1)] == [0, 0]
    assert L.polcue_lut_steep(h, 2, xys.ctypes.data) == 1
    xk, yk = O.sorted_knots(1.8)["spec2"]
    assert xys[0] == xk[-2] and xys[1] == yk[-2] and abs(xys[2] + 10796.578) < 1e-2
    q = np.concatenate((np.linspace(0.99999, 1.00001, 401), np.linspace(1.0, 2.2, 1000))).astype(np.float32)
    theta = np.empty(q.size, np.float32)
    assert L.polcue_lut_eval_host(h, 2, q.ctypes.data, q.size, theta.ctypes.data) == 0
    ref = O.interp_linear_extrap(xk, yk, q.astype(np.float64))
    assert np.abs(ref).max() > 1e4
    assert (np.abs(theta - ref) <= 6.0e-8 * np.abs(ref) + 7e-6).all()     # one float32 rounding of theta (+ 66 * ulp(rho) below the steep segment)
    L.polcue_lut_destroy(h)
    for n in (1.2, 1.33, 1.5, 2.4):
        rc, h = _host_lut(L, n)
        assert [L.polcue_lut_steep(h, t, None) for t in range(3)] == [0, 0, 0]
        L.polcue_lut_destroy(h)


@pytest.mark.parametrize("geometry", [(832, 1088, 320, 480), (1024, 1224, 320, 480), (33, 47, 66, 94), (50, 70, 50, 30), (7, 9, 3, 2),
                                      (2048, 2448, 192, 640), (5, 5, 5, 5)])
def test_resize_plan_weights_equal_the_pillow_restatement(L, geometry):
    """The library's Lanczos weights (host C++, libm) against the oracle's numpy restatement of Resample.c, integer for integer."""
    ih, iw, oh, ow = geometry
    h = C.c_void_p()
    assert L.polcue_resize_plan_host_build(ih, iw, oh, ow, C.byref(h)) == 0
    for axis, (n_in, n_out) in enumerate(((iw, ow), (ih, oh))):
        ksize = L.polcue_resize_plan_coeffs(h, axis, None, None, 0)
        bounds = np.empty((n_out, 2), np.int32)
        kk = np.empty((n_out, ksize), np.int32)
        assert L.polcue_resize_plan_coeffs(h, axis, bounds.ctypes.data, kk.ctypes.data, kk.size) == ksize
        if n_in == n_out:        # Pillow skips the pass; the library runs it with single-tap identity weights
            assert ksize == 1 and (kk == 1 << 22).all() and np.array_equal(bounds[:, 0], np.arange(n_out)) and (bounds[:, 1] == 1).all()
            continue
        ks, b_ref, kk_ref = O.lanczos_coeffs_8bpc(n_in, n_out)
        assert ks == ksize and np.array_equal(bounds, b_ref) and np.array_equal(kk, kk_ref)
        assert (np.abs(kk.sum(axis=1) - (1 << 22)) <= ksize).all()          # weights sum to one in fixed point
    assert L.polcue_resize_workspace_bytes(h, 8) == (8 * ih + 32) * ow
    L.polcue_resize_plan_destroy(h)
    assert L.polcue_resize_plan_host_build(0, 4, 4, 4, C.byref(h)) == -22


def test_tables_for_random_refractive_indices(L):
    """Any n the builder accepts: float32 cells + float64 steep segments reproduce scipy's interpolant to float32 rounding
    of theta (+ 2e-6), in-table, at the peak and in both extrapolation ranges.  (A 78-index sweep found n = 1.5548 with an
    end slope of -53 006.)"""
    rng = np.random.default_rng(7)
    for n in np.concatenate((rng.uniform(1.02, 3.0, 10), [1.5548, 1.8])):
        rc, h = _host_lut(L, n)
        assert rc == 0, n
        knots = O.sorted_knots(float(n))
        q = np.concatenate((rng.uniform(0, 1, 8000) ** 2, rng.uniform(0, 2.2, 3000), 1 - rng.uniform(0, 1, 3000) ** 3 * 1e-3,
                            [0.0, 1.0, 2.0])).astype(np.float32)
        for t, name in enumerate(("diffuse", "spec1", "spec2")):
            xk, yk = knots[name]
            theta = np.empty(q.size, np.float32)
            assert L.polcue_lut_eval_host(h, t, q.ctypes.data, q.size, theta.ctypes.data) == 0
            ref = O.interp_linear_extrap(xk, yk, q.astype(np.float64))
            assert (np.abs(theta - ref) <= 1.2e-7 * np.abs(ref) + 2e-5).all(), (n, name)
        L.polcue_lut_destroy(h)


def test_table_anchor_values(L):
    rc, h = _host_lut(L, 1.5)
    anchors = {0.0: (0.0, 0.0, 1.570796327), 0.01: (0.410434783, 0.086477641, 1.566324189), 0.3: (1.472298774, 0.460030611, 1.436485518),
               0.5: (1.692111847, 0.590509261, 1.345596980), 1.0: (2.217812434, 0.981710534, 0.982727665),
               1.2: (2.428092719, 9.416327864, -27.260250003), 2.0: (3.269213607, 43.154787130, -140.232127007)}
    q = np.array(list(anchors), np.float32)
    for t in range(3):
        theta = np.empty(q.size, np.float32)
        L.polcue_lut_eval_host(h, t, q.ctypes.data, q.size, theta.ctypes.data)
        ref = np.array([anchors[k][t] for k in anchors])
        assert np.allclose(theta, ref, rtol=3e-6, atol=1e-6), (t, theta, ref)
    L.polcue_lut_destroy(h)


def test_unrepresentable_refractive_indices_are_rejected(L):
    for n in (1.0, 0.5, float("nan"), 1.0001):
        rc, h = _host_lut(L, n)
        assert rc == -34 and not h.value
    assert L.polcue_lut_host_build(1.5, None) == -22


def test_argument_validation_needs_no_gpu(L):
    assert L.polcue_split_pol(None, 1, 4, 4, 1, None, None, None, None, None) == -22
    assert L.polcue_fused_mosaic_u8(None, 1, 4, 4, None, None, None, None, None, None) == -22
    assert L.polcue_depth_to_normals_f32(None, None, 1, 4, 4, None, None) == -22
    assert L.polcue_depth_errors_f32(None, None, 4, None, None, None, None) == -22
    assert L.polcue_depth_errors_workspace_bytes() >= 64 + 148 * 8 * 8 * 8


def test_ops_refuse_cpu_tensors_and_missing_library():
    from polcue import ops
    with pytest.raises(TypeError):
        ops.fused_mosaic(torch.zeros((1, 4, 4), dtype=torch.uint8))
    with pytest.raises(TypeError):
        ops.get_normals(torch.zeros((1, 2, 4, 4)))
    code = ("import sys; sys.path.insert(0, %r); import polcue._lib as L; L.LIB_PATH = '/nonexistent/libpolcue.so'\n"
            "try:\n    L.lib()\nexcept RuntimeError as e:\n    print('RAISED', 'no CPU fallback' in str(e))\n") % os.path.join(
        ROOT, "supervised-depth-estimation-from-polarized-images_b200")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True).stdout
    assert "RAISED True" in out


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "polcue_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_shadow_serves_the_reference_module_names():
    """`polcue.compat.shadow()` serves `polarisation.*` and `ppp_code.physical_normals_channels` under the reference's names,
    and refuses when the reference's own modules are already imported (that case is `install()`'s)."""
    pkg = os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200")
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import polcue.compat as c; served = c.shadow()\n"
            "from polarisation.xolp import Iun_and_xolp\n"
            "from polarisation.pol_split_and_save import split_pol\n"
            "from polarisation.xolp_and_normals import rho_spec, rho_diffuse, calc_normals\n"
            "from ppp_code.physical_normals_channels import PolarisationImage_channel, calc_normals_channel\n"
            "import polarisation, polcue.compat.xolp as cx\n"
            "print('OK', Iun_and_xolp is cx.Iun_and_xolp, split_pol.__module__, len(served), polarisation.xolp is cx)\n") % pkg
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "OK True polcue.compat.pol_split_and_save 4 True" in out.stdout, out.stderr[-1500:]
    if os.path.isdir("/root/reference/polarisation"):
        code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, '/root/reference')\n"
                "import polarisation.xolp, polcue.compat as c\n"
                "try:\n    c.shadow()\nexcept RuntimeError as e:\n    print('REFUSED', 'install()' in str(e))\n") % pkg
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
        assert "REFUSED True" in out.stdout, out.stderr[-1500:]


@pytest.mark.skipif(not os.path.isdir("/root/reference/manydepth"), reason="reference checkout not present on this box")
def test_install_patches_the_imported_reference_modules():
    """`polcue.compat.install()` swaps the hot-path functions of the real reference modules in place (no GPU needed)."""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, '/root/reference')\n"
            "import manydepth.normals_vec as nv, manydepth.layers as ly, manydepth.networks.pre_encoders as pe, polarisation.xolp as px\n"
            "import polcue.compat as c, polcue.compat.normals_vec as cnv, polcue.compat.layers as cly\n"
            "done = c.install()\n"
            "assert nv.rho_diffuse is cnv.rho_diffuse and pe.rho_spec is cnv.rho_spec and ly.compute_depth_errors is cly.compute_depth_errors\n"
            "assert pe.ShallowNormalsEncoder.get_normals.__module__ == 'polcue.compat.normals_vec'\n"
            "assert px.Iun_and_xolp.__module__ == 'polcue.compat.xolp'\n"
            "print('PATCHED', len(done))\n") % os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "PATCHED" in out.stdout and int(out.stdout.split()[-1]) >= 10, out.stderr[-1500:]


def test_install_patches_the_trainer_and_evaluation_classes():
    """manydepth.trainer / manydepth.evaluation cannot be imported here (kornia, skimage ... are absent), so stand-in
    modules with the reference's class and method names check the patch list of `install()`."""
    import types
    import polcue.compat as c
    from polcue.compat import trainer as c_tr
    fakes = {}
    for mod_name, cls_name, methods in (("manydepth.trainer", "Trainer", ("compute_supervised_normals_losses", "compute_depth_losses_from_list")),
                                        ("manydepth.evaluation", "Evaluation", ("compute_depth_losses_from_list",))):
        mod = types.ModuleType(mod_name)
        cls = type(cls_name, (), {m: (lambda self, *a, **k: "reference") for m in methods})
        setattr(mod, cls_name, cls)
        mod.compute_depth_errors = lambda gt, pred: "reference"
        fakes[mod_name] = mod
    loader = types.ModuleType("manydepth.datasets.indoor_dataset")
    loader.Iun_and_xolp = lambda images, angles: "reference"
    fakes["manydepth.datasets.indoor_dataset"] = loader
    saved = {k: sys.modules.get(k) for k in fakes}
    sys.modules.update(fakes)
    try:
        done = c.install()
        # the loader-side function runs in forked DataLoader workers: left alone unless asked for
        assert loader.Iun_and_xolp(None, None) == "reference" and not any("indoor_dataset" in d for d in done)
        assert "manydepth.datasets.indoor_dataset.Iun_and_xolp" in c.install(patch_loader=True)
        assert loader.Iun_and_xolp.__module__ == "polcue.compat.xolp"
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert "manydepth.trainer.Trainer.compute_depth_losses_from_list" in done
    assert "manydepth.evaluation.Evaluation.compute_depth_losses_from_list" in done
    assert "manydepth.trainer.Trainer.compute_supervised_normals_losses" in done and "manydepth.trainer.compute_depth_errors" in done
    assert fakes["manydepth.trainer"].Trainer.compute_depth_losses_from_list is c_tr.compute_depth_losses_from_list
    assert fakes["manydepth.evaluation"].Evaluation.compute_depth_losses_from_list is c_tr.compute_depth_losses_from_list


def test_synthetic_generators_are_seeded_per_frame():
    from polcue import synth
    a = synth.gen_batch("P", 3, 2, 32, 48)
    assert np.array_equal(a[1], synth.gen_p_mosaic(4, 32, 48))
    assert np.array_equal(synth.gen_u_mosaic(7, 32, 48), synth.gen_u_mosaic(7, 32, 48))
    st = O.stack_quadrants(synth.gen_p_mosaic(0, 64, 96))
    _, rho, _ = O.iun_and_xolp_closed(st)
    assert 0.0 <= rho.min() and rho.max() < 0.8        # physically plausible DoLP
    gt, pred, inst, k = synth.gen_depth_sample(0, 32, 48)
    assert gt.dtype == np.float32 and set(np.unique(inst)) <= set(range(0, 201, 20)) and k.shape == (3, 3)
    assert 0.05 < (gt == 0).mean() < 0.15


def test_shard_ranges_partition_exactly():
    from polcue.dist import shard_range
    for total in (0, 1, 7, 120, 10_000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import numpy as np, torch
from polcue import dist as D, synth
from oracle import polcue_oracle as O
rank, _, world = D.init("gloo")
n_images = 7
gt, pred, inst, _ = synth.gen_depth_batch(0, n_images, 32, 48)
lo, hi = D.shard_range(n_images, rank, world)
# each rank finalises its own images (here with the oracle standing in for the kernel: host logic only)
rows, _ = O.depth_errors_per_image(gt[lo:hi], pred[lo:hi], 0.1, 2.0)
mean = D.mean_over_images(torch.from_numpy(rows))
pooled = torch.zeros(8, dtype=torch.float64)
for b in range(lo, hi):
    m = (gt[b] > 0.1) & (gt[b] < 2.0)
    pooled += torch.from_numpy(O.depth_error_sums(gt[b][m], np.clip(pred[b][m], 0.1, 2.0)))
D.all_reduce_sums(pooled)
t = D.max_over_ranks(float(rank + 1), "cpu")
if rank == 0:
    np.save(sys.argv[3], np.concatenate((mean.numpy(), pooled.numpy(), [t])))
D.barrier()
"""


def test_two_rank_gloo_reduction_matches_single_process(tmp_path):
    from polcue import synth
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    out = tmp_path / "out.npy"
    pkg = os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", str(script), pkg, ROOT, str(out)]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    got = np.load(out)
    gt, pred, _, _ = synth.gen_depth_batch(0, 7, 32, 48)
    rows, mean = O.depth_errors_per_image(gt, pred, 0.1, 2.0)
    assert np.allclose(got[:7], mean, rtol=1e-12)
    pooled = np.zeros(8)
    for b in range(7):
        m = (gt[b] > 0.1) & (gt[b] < 2.0)
        pooled += O.depth_error_sums(gt[b][m], np.clip(pred[b][m], 0.1, 2.0))
    assert np.allclose(got[7:15], pooled, rtol=1e-12)
    assert got[15] == 2.0


def test_bench_reference_arm_prints_one_contract_line_with_blas_pinned():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the CUDA arm): one JSON line on stdout with the
    contract's keys, the same workload label as the CUDA arm, the requested frames / steps honoured, and worker
    processes whose BLAS really runs single-threaded (round 1 pinned the variables too late and measured 3x low)."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--frames", "2", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-1500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
                "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["steps"] == 2 and line["config"]["frames_per_step"] == 2 and line["gpu_launches"] == 0
    assert line["config"]["workload"].startswith("cfg2:") and line["unit"] == "Mpix/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 2
    assert "observed with threadpoolctl: 1;" in line["cpu_baseline"]["sample"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_bench_side_modules_import_and_encoder_standins_have_the_reference_geometry():
    """bench.py's cfg3 / cfg4 / cfg5 blocks live in tools/workloads.py; the encoder stand-ins it times must keep the layer
    geometry of pre_encoders.py:49-97 / resnet_encoder.py:783-822 (2 / 9 -> 64 channels at 1/8 resolution, resnet18 stem +
    layer1 + layer2 -> 128 channels at 1/8)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import workloads                                   # noqa: F401  (imports polcue.ops; no GPU call at import time)
    from encoder_standins import PolarEncoders
    enc = PolarEncoders().eval()
    with torch.no_grad():
        fx, fn, fr = enc(torch.rand(2, 2, 64, 96), torch.rand(2, 9, 64, 96), torch.rand(2, 3, 64, 96))
    assert fx.shape == (2, 64, 8, 12) and fn.shape == (2, 64, 8, 12) and fr.shape == (2, 128, 8, 12)
    convs = [m for m in enc.xolp_branch.modules() if isinstance(m, torch.nn.Conv2d)]
    assert [c.kernel_size[0] for c in convs] == [7, 3, 3, 5, 3, 3, 5, 3, 3] and convs[0].stride == (2, 2)
