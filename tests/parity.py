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


STENCIL_ROUNDINGS = 4.0          # float32 evaluations of the stencil stay below this many unit round-offs of the model


def assert_stencil_normals_close(got, depth, camera_matrix, oracle, what="depth_to_normals", max_ill_fraction=2e-3):
    """Parity of a float32 depth->normals result with the float64 oracle on depth maps WITH zero-depth holes.

    Every pixel is checked, none is skipped:
      * where both Sobel gradients are exactly zero (all eight neighbours invalid: hole interiors) the result must be the
        exact zero vector, as F.normalize gives it in any arithmetic;
      * elsewhere the angular error must be <= max(1e-3 rad, STENCIL_ROUNDINGS * 2^-24 * bound), `bound` being the
        oracle's float64 conditioning model (oracle.depth_to_normals_conditioning), and the normal must have unit length;
      * the pixels whose model bound exceeds 1e-3 rad (the cross product cancels: exactly or nearly parallel gradients,
        e.g. one isolated valid neighbour; float32 cannot reach 1e-3 there, in the reference's float32 torch ops either)
        are counted and must stay below `max_ill_fraction` of the image; they must still be finite and of length <= 1.
    Returns (worst error over the well-conditioned pixels, number of ill-conditioned pixels)."""
    got = np.asarray(got, np.float64)
    assert np.isfinite(got).all(), f"{what}: non-finite normals"
    ref = oracle.depth_to_normals(depth, camera_matrix)
    _, bound, dead = oracle.depth_to_normals_conditioning(depth, camera_matrix)
    assert (got[np.broadcast_to(dead[:, None], got.shape)] == 0).all(), f"{what}: non-zero normal inside a zero-depth hole"
    lim = np.maximum(NORMAL_TOL, STENCIL_ROUNDINGS * 2.0 ** -24 * bound)
    ill = (lim > NORMAL_TOL) & ~dead
    err = np.where(dead | ~np.isfinite(lim), 0.0, angular_error(got, ref, axis=1))
    bad = ~(err <= lim)
    assert not bad.any(), (f"{what}: {int(bad.sum())} px beyond max(1e-3, model bound), worst {np.nanmax(err - lim):.3e} rad over, first at "
                           f"{np.unravel_index(np.argmax(bad), bad.shape)}")
    assert ill.mean() <= max_ill_fraction, f"{what}: {int(ill.sum())} ill-conditioned px ({ill.mean():.2e} of the image)"
    length = np.sqrt((got * got).sum(axis=1))
    assert (length < 1 + 1e-5).all() and ((np.abs(length - 1) < 1e-5) | dead | ill).all(), f"{what}: normals are not unit length"
    well = ~ill & ~dead
    return (float(err[well].max()) if well.any() else 0.0), int(ill.sum())
