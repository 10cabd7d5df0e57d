"""Parity protocol shared by the CPU (oracle vs golden) and GPU (kernel vs oracle) tests.

Tolerances are BASELINE.json's north_star: Stokes/DoLP 1e-5 relative (+1e-7 absolute floor),
AoLP 1e-4 rad wrap-aware (mod pi), normals 1e-3 rad angular error, split bit-exact.
"""
import numpy as np

DOLP_RTOL, DOLP_ATOL = 1e-5, 1e-7
AOLP_TOL = 1e-4
NORMAL_TOL = 1e-3


def assert_dolp_close(got, ref, what="rho"):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    err = np.abs(got - ref)
    lim = DOLP_RTOL * np.abs(ref) + DOLP_ATOL
    bad = ~(err <= lim)
    assert not bad.any(), f"{what}: {bad.sum()} px out of tolerance, worst {np.nanmax(err - lim):.3e}"


def aolp_error(got, ref):
    d = np.abs(np.asarray(got, np.float64) - np.asarray(ref, np.float64)) % np.pi
    return np.minimum(d, np.pi - d)


def assert_aolp_close(got, ref, exclude=None, tol=AOLP_TOL):
    """A NaN (or inf) AoLP outside `exclude` fails: ~(err <= tol) is true for NaN errors."""
    err = aolp_error(got, ref)
    if exclude is not None:
        err = np.where(exclude, 0.0, err)
    bad = ~(err <= tol)
    assert not bad.any(), (f"AoLP: {int(bad.sum())} px out of tolerance (NaN counted), first at "
                           f"{np.unravel_index(np.argmax(bad), bad.shape)}, worst finite {np.nanmax(err):.3e} rad")


def angular_error(a, b, axis):
    """Angle between two unit-ish vector fields along `axis` (robust near 0 via atan2)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    cr = np.linalg.norm(np.cross(a, b, axis=axis), axis=axis)
    dt = (a * b).sum(axis=axis)
    return np.arctan2(cr, dt)


def assert_normals_close(got, ref, axis, twin_ok=None, skip=None, tol=NORMAL_TOL, what="normals"):
    """twin_ok: mask of pixels where the (-Nx, -Ny, Nz) twin of `ref` is also accepted."""
    err = angular_error(got, ref, axis)
    if twin_ok is not None:
        sign = np.array([-1.0, -1.0, 1.0]).reshape([3 if i == (axis % np.ndim(ref)) else 1 for i in range(np.ndim(ref))])
        err = np.where(twin_ok, np.minimum(err, angular_error(got, np.asarray(ref) * sign, axis)), err)
    if skip is not None:
        err = np.where(skip, 0.0, err)
    bad = ~(err <= tol)                     # NaN normals (bad table index, 0 * inf, unwritten tail) count as failures
    assert not bad.any(), (f"{what}: {int(bad.sum())} px out of tolerance (NaN counted), first at "
                           f"{np.unravel_index(np.argmax(bad), bad.shape)}, worst finite {np.nanmax(err):.3e} rad")
    return float(err.max()) if err.size else 0.0
