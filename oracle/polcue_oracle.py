"""CPU oracle for the polarization hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The shipped package (``polcue``) never
does; it fails loudly when the CUDA library is missing.

Every function restates, in float64 numpy, one reference function of the path and cites the
reference file:line it follows (paths relative to the reference repository root).

Pinning status
--------------
* quadrant split, XOLP, Fresnel tables, physics normals, ``get_normals``, depth metrics:
  PINNED -- the reference has no tests or golden vectors of its own, so these restatements are
  checked against outputs of the reference functions themselves, generated in the authoring
  container by ``tests/golden/make_golden.py`` and committed as ``tests/golden/*.npz``.
* ``ppp_code/physical_normals_channels.py`` (needs matplotlib/TkAgg to import): the numpy
  bodies were executed with a stub ``matplotlib`` by the same script  ->  PINNED.
* loader front end (``indoor_dataset.py:115,335-349``: ``transforms.Resize(..., Image.ANTIALIAS)`` of the four
  8-bit polarizer images): the arithmetic lives in Pillow (pinned 6.2.1, ``environment.yml:14``; installed 12.2),
  ``src/libImaging/Resample.c``.  PINNED by execution: ``tests/test_oracle_golden.py`` runs the installed Pillow on the
  same inputs, and ``tests/golden/resize_golden.npz`` holds Pillow outputs made by ``tests/golden/make_golden.py``.
* ``depth_to_normals``: the algorithm lives in kornia 0.5.11 (``environment.yml:41``), which is
  neither vendored in the reference nor installed here.  PARITY UNPINNED: the restatement
  follows kornia's published algorithm (depth_to_3d -> spatial_gradient(sobel, normalized,
  replicate pad) -> cross -> normalize) and is anchored only on the reference call sites
  ``manydepth/trainer.py:1305-1306`` and on self-consistency properties.
"""
from __future__ import annotations

import numpy as np

CANONICAL_ANGLES = np.array([0.0, 45.0, 90.0, 135.0]) * np.pi / 180.0
N_TABLE = 1000  # knots of every Fresnel table (normals_vec.py:13,27)


# --------------------------------------------------------------------------------------
# a1 / a2 -- quadrant split and stack order
# --------------------------------------------------------------------------------------
def split_pol(img):
    """Quadrant split, return order (im00, im10, im01, im11).

    Follows polarisation/pol_split_and_save.py:10-27 (np.split on axis 1 then axis 0; H and W
    must be even, np.split raises ValueError otherwise).  The reference's two ``print`` calls
    are not reproduced.
    """
    img = np.asarray(img)
    h, w = img.shape[0], img.shape[1]
    if h % 2 or w % 2:
        raise ValueError("array split does not result in an equal division")
    hh, hw = h // 2, w // 2
    return img[:hh, :hw], img[hh:, :hw], img[:hh, hw:], img[hh:, hw:]


def stack_quadrants(img):
    """H x W mosaic -> Hs x Ws x 4 stack in angle order (0, 45, 90, 135) = (TL, TR, BL, BR).

    Follows manydepth/datasets/indoor_dataset.py:435-439 and polarisation/xolp_and_normals.py:107-109.
    """
    im00, im10, im01, im11 = split_pol(img)
    return np.stack((im00, im01, im10, im11), axis=2)


# --------------------------------------------------------------------------------------
# a3 -- XOLP
# --------------------------------------------------------------------------------------
def iun_and_xolp_lstsq(images, angles):
    """Faithful restatement of polarisation/xolp.py:8-34 (LAPACK least squares per pixel).

    Returns (Iun, rho, phi), float64, each H x W.  Inherits the reference's round-off
    behaviour on tie pixels: use `iun_and_xolp_closed` for the deterministic form.
    """
    images = np.asarray(images)
    h, w = images.shape[0], images.shape[1]
    samples = images.reshape(h * w, 4)
    design = np.stack((np.ones(4), np.cos(2 * angles), np.sin(2 * angles)), axis=1)  # xolp.py:15-18
    coef = np.linalg.lstsq(design, samples.T, rcond=None)[0].T                        # xolp.py:20-21
    amp = np.sqrt(coef[:, 1] ** 2 + coef[:, 2] ** 2)
    i_max = coef[:, 0] + amp                                                          # xolp.py:22
    i_min = coef[:, 0] - amp                                                          # xolp.py:23
    iun = (i_max + i_min) / 2                                                         # xolp.py:24
    with np.errstate(divide="ignore", invalid="ignore"):
        rho = np.true_divide(i_max - i_min, i_max + i_min)                            # xolp.py:27
        rho[rho == np.inf] = 0                                                        # xolp.py:28
        rho = np.nan_to_num(rho)                                                      # xolp.py:29
    phi = 0.5 * np.arctan2(coef[:, 2], coef[:, 1])                                    # xolp.py:30
    return iun.reshape(h, w), rho.reshape(h, w), phi.reshape(h, w)


def pinv_design(angles):
    """3x4 pseudo-inverse of the design matrix of xolp.py:15-18 (what lstsq applies)."""
    angles = np.asarray(angles, dtype=np.float64)
    design = np.stack((np.ones(4), np.cos(2 * angles), np.sin(2 * angles)), axis=1)
    return np.linalg.pinv(design)


def iun_and_xolp_closed(images, angles=None):
    """Deterministic closed form of xolp.py:8-34.

    For the canonical angles: x0 = sum/4, x1 = (I0-I90)/2, x2 = (I45-I135)/2 exactly (the
    normal equations of xolp.py:15-21 with cos(90deg)=0 taken exactly), so no tie-pixel
    round-off.  For other angles: x = pinv(A) @ I.
    """
    images = np.asarray(images, dtype=np.float64)
    if angles is None or np.allclose(angles, CANONICAL_ANGLES, rtol=0, atol=1e-12):
        x0 = images.sum(axis=-1) / 4.0
        x1 = (images[..., 0] - images[..., 2]) / 2.0
        x2 = (images[..., 1] - images[..., 3]) / 2.0
    else:
        p = pinv_design(angles)
        x0, x1, x2 = (images @ p[k] for k in range(3))
    amp = np.sqrt(x1 * x1 + x2 * x2)
    with np.errstate(divide="ignore", invalid="ignore"):
        rho = amp / x0
        rho[rho == np.inf] = 0
        rho = np.nan_to_num(rho)
    phi = 0.5 * np.arctan2(x2, x1)
    return x0, rho, phi


def xolp_tie_masks(images):
    """Masks of the pixels where the reference's lstsq output is decided by round-off.

    sign_tie : I45 == I135 and I0 < I90   -> phi = +-pi/2 at random
    degenerate: I0 == I90 and I45 == I135 -> phi arbitrary, rho ~ 1e-16
    """
    im = np.asarray(images).astype(np.int64)
    s1 = im[..., 0] - im[..., 2]
    s2 = im[..., 1] - im[..., 3]
    return (s2 == 0) & (s1 < 0), (s1 == 0) & (s2 == 0)


# --------------------------------------------------------------------------------------
# a4 / a5 -- Fresnel tables and the piecewise-linear inverse
# --------------------------------------------------------------------------------------
def fresnel_tables(n):
    """theta grid, diffuse DoLP table, specular DoLP table and its argmax.

    Follows manydepth/normals_vec.py:13-19 (diffuse) and :27-40 (specular); identical formulas in
    polarisation/xolp_and_normals.py:47-59,75-81 and ppp_code/physical_normals_channels.py:40-46,52-64.
    """
    theta = np.linspace(0, np.pi / 2, N_TABLE)
    s = np.sin(theta)
    c = np.cos(theta)
    rho_d = ((n - 1 / n) ** 2 * s ** 2) / (
        2 + 2 * n ** 2 - (n + 1 / n) ** 2 * s ** 2 + 4 * c * np.sqrt(n ** 2 - s ** 2))
    rho_s = (2 * s ** 2 * c * np.sqrt(n ** 2 - s ** 2)) / (
        n ** 2 - s ** 2 - n ** 2 * s ** 2 + 2 * s ** 4)
    return theta, rho_d, rho_s, int(np.argmax(rho_s))


def sorted_knots(n):
    """The three (x ascending, y) knot arrays exactly as scipy's interp1d holds them.

    interp1d(assume_sorted=False) stably argsorts x (mergesort) before use; branch 2 of the
    specular table (normals_vec.py:45-47) is descending in rho and is therefore reversed.
    Returns dict name -> (x, y) for 'diffuse', 'spec1', 'spec2'.
    """
    theta, rho_d, rho_s, imax = fresnel_tables(n)
    out = {}
    for name, x, y in (("diffuse", rho_d, theta),
                       ("spec1", rho_s[:imax], theta[:imax]),
                       ("spec2", rho_s[imax:], theta[imax:])):
        order = np.argsort(x, kind="mergesort")
        out[name] = (x[order], y[order])
    return out


def interp_linear_extrap(xk, yk, xq):
    """scipy.interpolate.interp1d(kind='linear', fill_value='extrapolate') on sorted knots.

    Restates scipy 1.18.1 ``interp1d._call_linear``: searchsorted(side='left'), clip the index
    to [1, len-1], then  ((x-x_lo)/(x_hi-x_lo))*y_hi + ((x_hi-x)/(x_hi-x_lo))*y_lo ; with
    'extrapolate' the same end segments serve out-of-range queries.
    """
    xq = np.asarray(xq, dtype=np.float64)
    flat = xq.ravel()
    hi = np.clip(np.searchsorted(xk, flat), 1, len(xk) - 1)
    lo = hi - 1
    x_lo, x_hi, y_lo, y_hi = xk[lo], xk[hi], yk[lo], yk[hi]
    width = x_hi - x_lo
    val = ((flat - x_lo) / width) * y_hi + ((x_hi - flat) / width) * y_lo
    return val.reshape(xq.shape)


def rho_diffuse(rho, n):
    """Zenith angle from diffuse DoLP; normals_vec.py:11-22 / xolp_and_normals.py:69-83."""
    xk, yk = sorted_knots(n)["diffuse"]
    return interp_linear_extrap(xk, yk, rho)


def rho_spec(rho, n):
    """Two zenith candidates from specular DoLP; normals_vec.py:25-50 / xolp_and_normals.py:41-67."""
    knots = sorted_knots(n)
    return (interp_linear_extrap(*knots["spec1"], rho),
            interp_linear_extrap(*knots["spec2"], rho))


# --------------------------------------------------------------------------------------
# a6 / a7 -- normals
# --------------------------------------------------------------------------------------
def calc_normals_hw3(phi, theta):
    """numpy layout H x W x 3; polarisation/xolp_and_normals.py:85-98."""
    return np.stack((np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta), np.cos(theta)), axis=-1)


def calc_normals_b3hw(phi, theta):
    """torch layout B x 3 x H x W; manydepth/normals_vec.py:53-60."""
    return np.stack((np.cos(phi) * np.sin(theta), np.sin(phi) * np.sin(theta), np.cos(theta)), axis=1)


def get_normals(x, n=1.5):
    """ShallowNormalsEncoder.get_normals, manydepth/networks/pre_encoders.py:99-113.

    x: B x 2 x H x W (rho, phi).  If x is float32 the `phi + pi/2` sum is rounded to float32
    exactly as torch does for a float32 tensor plus a python scalar; everything downstream is
    float64 (theta is float64, so torch promotes).  Returns B x 9 x H x W float64.
    """
    x = np.asarray(x)
    rho, phi = x[:, 0], x[:, 1]
    # torch evaluates cos/sin of phi in phi's own dtype before promoting (normals_vec.py:56-57)
    if phi.dtype == np.float32:
        phi_s = (phi + np.float32(np.pi / 2)).astype(np.float32)
    else:
        phi_s = phi + np.pi / 2
    cphi, sphi = np.cos(phi).astype(np.float64), np.sin(phi).astype(np.float64)
    cphi_s, sphi_s = np.cos(phi_s).astype(np.float64), np.sin(phi_s).astype(np.float64)
    th_d = rho_diffuse(rho, n)
    th_1, th_2 = rho_spec(rho, n)
    chans = []
    for cp, sp, th in ((cphi, sphi, th_d), (cphi_s, sphi_s, th_1), (cphi_s, sphi_s, th_2)):
        chans += [cp * np.sin(th), sp * np.sin(th), np.cos(th)]
    return np.stack(chans, axis=1)


# --------------------------------------------------------------------------------------
# a8 -- masked direct-Stokes "channel" variant
# --------------------------------------------------------------------------------------
def polarisation_image_channel(images, angles, mask):
    """ppp_code/physical_normals_channels.py:15-36.  `angles` is ignored there too.

    s0 = I0 + I90 (NOT the 4-sample mean), no nan scrub (0/0 stays NaN where mask is true),
    outputs zero outside mask.  Return order (rho, phi, Iun).
    """
    images = np.asarray(images, dtype=np.float64)
    mask = np.asarray(mask, dtype=bool)
    s0 = images[..., 0] + images[..., 2]
    s1 = images[..., 0] - images[..., 2]
    s2 = images[..., 1] - images[..., 3]
    with np.errstate(divide="ignore", invalid="ignore"):
        rho = np.sqrt(s1 ** 2 + s2 ** 2) / s0
    phi = 0.5 * np.arctan2(s2, s1)
    iun = s0 / 2
    z = np.zeros(mask.shape)
    return np.where(mask, rho, z), np.where(mask, phi, z), np.where(mask, iun, z)


def calc_normals_channel(phi, theta, mask):
    """ppp_code/physical_normals_channels.py:75-83: H x W x 3, zero outside mask."""
    nrm = calc_normals_hw3(phi, theta)
    return np.where(np.asarray(mask, dtype=bool)[..., None], nrm, 0.0)


# --------------------------------------------------------------------------------------
# a10 -- depth -> normals (kornia 0.5.11 algorithm; PARITY UNPINNED, see module docstring)
# --------------------------------------------------------------------------------------
def depth_to_normals(depth, camera_matrix, dtype=np.float64):
    """kornia.geometry.depth.depth_to_normals as called at manydepth/trainer.py:1305-1306.

    depth: B x 1 x H x W, camera_matrix: B x 3 x 3 (pixel units).  Steps (kornia 0.5.11):
      xyz  = ((u-cx)/fx * Z, (v-cy)/fy * Z, Z), u in [0,W-1], v in [0,H-1]       (depth_to_3d)
      grad = cross-correlation of replicate-padded xyz with sobel_x/8 and sobel_y/8  (spatial_gradient)
      n    = cross(d xyz/du, d xyz/dv);  n / max(||n||, 1e-12)                      (F.normalize)
    Returns B x 3 x H x W.
    """
    depth = np.asarray(depth, dtype=dtype)
    km = np.asarray(camera_matrix, dtype=dtype)
    b, _, h, w = depth.shape
    u = np.arange(w, dtype=dtype)[None, None, :]
    v = np.arange(h, dtype=dtype)[None, :, None]
    fx, fy = km[:, 0, 0][:, None, None], km[:, 1, 1][:, None, None]
    cx, cy = km[:, 0, 2][:, None, None], km[:, 1, 2][:, None, None]
    z = depth[:, 0]
    xyz = np.stack((((u - cx) / fx) * z, ((v - cy) / fy) * z, z), axis=1)  # B x 3 x H x W
    pad = np.pad(xyz, ((0, 0), (0, 0), (1, 1), (1, 1)), mode="edge")

    def win(dy, dx):
        return pad[:, :, 1 + dy:1 + dy + h, 1 + dx:1 + dx + w]

    eighth = dtype(1) / dtype(8)
    gu = ((win(-1, 1) - win(-1, -1)) + 2 * (win(0, 1) - win(0, -1)) + (win(1, 1) - win(1, -1))) * eighth
    gv = ((win(1, -1) - win(-1, -1)) + 2 * (win(1, 0) - win(-1, 0)) + (win(1, 1) - win(-1, 1))) * eighth
    nrm = np.cross(gu, gv, axis=1)
    length = np.sqrt((nrm * nrm).sum(axis=1, keepdims=True))
    return nrm / np.maximum(length, dtype(1e-12))


def depth_to_normals_conditioning(depth, camera_matrix):
    """Float64 forward-error model of a float32 evaluation of `depth_to_normals` (any association order), per pixel.

    Returns (n, bound, dead): n = the UN-normalised cross product gu x gv (B x 3 x H x W); dead (B x H x W) = both
    gradients are exactly the zero vector (all eight neighbours invalid: the normal is the exact zero vector in ANY
    arithmetic); bound (B x H x W) = the first-order angular error, in radians per unit round-off, of the normalised cross
    product when every input of the Sobel sums carries a relative error of one unit round-off (inf where n vanishes
    although the gradients do not, i.e. they are exactly parallel -- one isolated valid neighbour -- where the result is
    0 or a round-off direction depending on FMA contraction, in the reference's CUDA ops as well):
        |d gu| <= A_u = sum |w_u| |xyz| / 8,   |d gv| <= A_v,
        |d n|  <= |A_u| |gv| + |gu| |A_v| + |gu| |gv|,        bound = |d n| / |n|.
    A float32 implementation with k roundings per term stays below k * 2^-24 * bound; tests use it to tell the pixels
    where 1e-3 rad is attainable in float32 (nearly all) from those where the cross product cancels (gradients nearly
    parallel next to zero-depth holes), and to bound the error there as well.  Test infrastructure only.
    """
    depth = np.asarray(depth, dtype=np.float64)
    km = np.asarray(camera_matrix, dtype=np.float64)
    b, _, h, w = depth.shape
    u = np.arange(w, dtype=np.float64)[None, None, :]
    v = np.arange(h, dtype=np.float64)[None, :, None]
    fx, fy = km[:, 0, 0][:, None, None], km[:, 1, 1][:, None, None]
    cx, cy = km[:, 0, 2][:, None, None], km[:, 1, 2][:, None, None]
    z = depth[:, 0]
    xyz = np.stack((((u - cx) / fx) * z, ((v - cy) / fy) * z, z), axis=1)
    pad = np.pad(xyz, ((0, 0), (0, 0), (1, 1), (1, 1)), mode="edge")
    apad = np.abs(pad)

    def win(a, dy, dx):
        return a[:, :, 1 + dy:1 + dy + h, 1 + dx:1 + dx + w]

    gu = ((win(pad, -1, 1) - win(pad, -1, -1)) + 2 * (win(pad, 0, 1) - win(pad, 0, -1)) + (win(pad, 1, 1) - win(pad, 1, -1))) / 8
    gv = ((win(pad, 1, -1) - win(pad, -1, -1)) + 2 * (win(pad, 1, 0) - win(pad, -1, 0)) + (win(pad, 1, 1) - win(pad, -1, 1))) / 8
    au = (win(apad, -1, 1) + win(apad, -1, -1) + 2 * (win(apad, 0, 1) + win(apad, 0, -1)) + win(apad, 1, 1) + win(apad, 1, -1)) / 8
    av = (win(apad, 1, -1) + win(apad, -1, -1) + 2 * (win(apad, 1, 0) + win(apad, -1, 0)) + win(apad, 1, 1) + win(apad, -1, 1)) / 8
    n = np.cross(gu, gv, axis=1)
    length = lambda a: np.sqrt((a * a).sum(axis=1))
    dn = length(au) * length(gv) + length(gu) * length(av) + length(gu) * length(gv)
    with np.errstate(divide="ignore", invalid="ignore"):
        bound = np.where(length(n) > 0, dn / length(n), np.inf)
    dead = (gu == 0).all(axis=1) & (gv == 0).all(axis=1)
    return n, bound, dead


def normals_loss_torch(depth_gt, depth_pred, camera_matrix, mask):
    """Trainer.compute_supervised_normals_losses, manydepth/trainer.py:1298-1309, as a float64 torch graph (CPU) so that
    autograd supplies the reference gradient w.r.t. depth_pred.

    depth_to_normals follows kornia 0.5.11 (pad replicate + cross-correlation with sobel/8, cross, F.normalize);
    the similarity follows torch 1.7.1's F.cosine_similarity, w12 * rsqrt(clamp_min(w1 * w2, eps^2)) with eps = 1e-8
    (environment.yml pins torch 1.7.1).  PARITY UNPINNED for the kornia part, see the module docstring.
    Inputs: torch tensors B x 1 x H x W (camera_matrix B x 3 x 3).  Returns the zero-dim loss tensor.
    """
    import torch
    import torch.nn.functional as F

    def d2n(depth):
        b, _, h, w = depth.shape
        u = torch.arange(w, dtype=depth.dtype)[None, None, :].expand(b, h, w)
        v = torch.arange(h, dtype=depth.dtype)[None, :, None].expand(b, h, w)
        km = camera_matrix.to(depth.dtype)
        fx, fy = km[:, 0, 0, None, None], km[:, 1, 1, None, None]
        cx, cy = km[:, 0, 2, None, None], km[:, 1, 2, None, None]
        z = depth[:, 0]
        xyz = torch.stack(((u - cx) / fx * z, (v - cy) / fy * z, z), 1)
        kx = torch.tensor([[-1., 0., 1.], [-2., 0., 2.], [-1., 0., 1.]], dtype=depth.dtype) / 8
        ker = torch.stack((kx, kx.t()))[:, None]
        g = F.conv2d(F.pad(xyz.reshape(b * 3, 1, h, w), (1, 1, 1, 1), mode="replicate"), ker).view(b, 3, 2, h, w)
        return F.normalize(torch.cross(g[:, :, 0], g[:, :, 1], dim=1), dim=1, p=2, eps=1e-12)

    n_gt, n_pred = d2n(depth_gt), d2n(depth_pred)
    w12 = (n_gt * n_pred).sum(1)
    w1, w2 = (n_gt * n_gt).sum(1), (n_pred * n_pred).sum(1)
    cos = (w12 * torch.rsqrt(torch.clamp_min(w1 * w2, 1e-16))).unsqueeze(1)
    return ((2.0 - cos) * mask).sum() / mask.sum()


# --------------------------------------------------------------------------------------
# a9 -- depth error metrics
# --------------------------------------------------------------------------------------
METRIC_NAMES = ("abs_rel", "sq_rel", "rmse", "rmse_log", "a1", "a2", "a3")


def _threshold_ratio(gt, pred):
    """max(gt/pred, pred/gt) evaluated in the inputs' OWN dtype, as layers.py:542 / :562 do: for float32 inputs the
    quotients are float32-rounded before they are compared with 1.25**k, which decides borderline pixels."""
    gt, pred = np.asarray(gt).ravel(), np.asarray(pred).ravel()
    if gt.dtype != np.float32 or pred.dtype != np.float32:
        gt, pred = gt.astype(np.float64), pred.astype(np.float64)
    with np.errstate(all="ignore"):
        return np.maximum(gt / pred, pred / gt)


def compute_depth_errors(gt, pred):
    """manydepth/layers.py:539-577; order (abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3).

    The threshold ratio keeps the inputs' dtype (see `_threshold_ratio`); every mean is accumulated in float64.
    Empty inputs give NaN for every entry, like numpy/torch means of empty arrays.
    """
    ratio = _threshold_ratio(gt, pred)
    gt = np.asarray(gt, dtype=np.float64).ravel()
    pred = np.asarray(pred, dtype=np.float64).ravel()
    with np.errstate(all="ignore"):
        if gt.size == 0:
            return (np.nan,) * 7
        a1 = (ratio < 1.25).mean()
        a2 = (ratio < 1.25 ** 2).mean()
        a3 = (ratio < 1.25 ** 3).mean()
        diff = gt - pred
        rmse = np.sqrt((diff ** 2).mean())
        rmse_log = np.sqrt(((np.log(gt) - np.log(pred)) ** 2).mean())
        abs_rel = (np.abs(diff) / gt).mean()
        sq_rel = (diff ** 2 / gt).mean()
    return abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3


def depth_error_sums(gt, pred):
    """The eight additive accumulators behind `compute_depth_errors`:
    (count, n[t<1.25], n[t<1.25^2], n[t<1.25^3], sum d^2, sum dlog^2, sum |d|/gt, sum d^2/gt)."""
    ratio = _threshold_ratio(gt, pred)
    gt = np.asarray(gt, dtype=np.float64).ravel()
    pred = np.asarray(pred, dtype=np.float64).ravel()
    with np.errstate(all="ignore"):
        diff = gt - pred
        return np.array([gt.size, (ratio < 1.25).sum(), (ratio < 1.25 ** 2).sum(), (ratio < 1.25 ** 3).sum(),
                         (diff ** 2).sum(), ((np.log(gt) - np.log(pred)) ** 2).sum(),
                         (np.abs(diff) / gt).sum(), (diff ** 2 / gt).sum()], dtype=np.float64)


def depth_errors_per_image(gt, pred, min_depth, max_depth, inst=None, inst_id=None, median_scaling=False, clamp_first=False):
    """Per-image masked metrics of Trainer.compute_depth_losses_from_list, manydepth/trainer.py:1376-1428.
    `median_scaling` = the branch of :1413-1414 (`not depth_supervision and not train_stereo_only`):
    depth_pred *= np.median(depth_gt) / np.median(depth_pred) on the masked arrays, before the clamp.
    `inst_id` may be an inclusive (lo, hi) range (evaluation.py:259-262, "objects" = 20..160); `clamp_first` = the
    batch-level torch.clamp of trainer.py:1368-1370 / evaluation.py:223-225 ahead of everything else.

    gt, pred: B x H x W float; inst: B x H x W instance-id map or None; inst_id: material level
    (20, 40, ... 200; trainer.py:1389-1411) or None for object == "all".
    mask = (gt > min) & (gt < max) [& inst == inst_id]; pred clamped to [min, max]; metrics per
    image.  Returns (B x 7 per-image metrics, mean over images).  The reference's try/except
    quirk (re-appending the previous image's errors) is not reproduced: an empty mask yields NaN.
    """
    gt, pred = np.asarray(gt), np.asarray(pred)       # dtype preserved: see `_threshold_ratio`
    if gt.dtype == np.float32:
        min_depth, max_depth = np.float32(min_depth), np.float32(max_depth)
    rows = []
    for b in range(gt.shape[0]):
        m = (gt[b] > min_depth) & (gt[b] < max_depth)
        if inst is not None and inst_id is not None:
            lo, hi = (inst_id if isinstance(inst_id, (tuple, list)) else (inst_id, inst_id))
            m &= (np.asarray(inst[b]) >= lo) & (np.asarray(inst[b]) <= hi)
        g, q = gt[b][m], (np.clip(pred[b], min_depth, max_depth) if clamp_first else pred[b])[m]
        if median_scaling:
            with np.errstate(all="ignore"):
                q = q * (np.median(g) / np.median(q)) if g.size else q
        rows.append(compute_depth_errors(g, np.clip(q, min_depth, max_depth)))
    rows = np.array(rows, dtype=np.float64).reshape(gt.shape[0], 7)
    return rows, rows.mean(axis=0)


# --------------------------------------------------------------------------------------
# XOLP dataset statistics
# --------------------------------------------------------------------------------------
def xolp_statistics(dolp, aolp):
    """polarisation/xolp_mean_and_std_dev.py:26-32: mean and population std of all DoLP / AoLP values of a set of
    frames, and the two 'XOLP' averages the script prints.  dolp, aolp: N x H x W."""
    dolp, aolp = np.asarray(dolp, dtype=np.float64), np.asarray(aolp, dtype=np.float64)
    dm, ds = dolp.mean(axis=(0, 1, 2)), dolp.std(axis=(0, 1, 2))
    am, asd = aolp.mean(axis=(0, 1, 2)), aolp.std(axis=(0, 1, 2))
    return {"dolp_mean": dm, "dolp_std": ds, "aolp_mean": am, "aolp_std": asd,
            "xolp_mean": 0.5 * (dm + am), "xolp_std": 0.5 * (ds + asd)}


# --------------------------------------------------------------------------------------
# loader front end: PIL Lanczos resize of 8-bit images (manydepth/datasets/indoor_dataset.py:77,115,335-349)
# --------------------------------------------------------------------------------------
PIL_PRECISION_BITS = 32 - 8 - 2   # Resample.c: PRECISION_BITS


def _lanczos(x):
    """Resample.c lanczos_filter: sinc(x) sinc(x/3) on [-3, 3), 0 elsewhere."""
    def sinc(v):
        v = np.asarray(v, dtype=np.float64)
        out = np.ones_like(v)
        nz = v != 0.0
        out[nz] = np.sin(v[nz] * np.pi) / (v[nz] * np.pi)
        return out
    x = np.asarray(x, dtype=np.float64)
    return np.where((x >= -3.0) & (x < 3.0), sinc(x) * sinc(x / 3.0), 0.0)


def lanczos_coeffs_8bpc(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the whole-image box (0, in_size).
    Returns (ksize, bounds[out_size, 2] = (xmin, count), kk[out_size, ksize] int32)."""
    scale = float(np.float32(in_size) - np.float32(0)) / out_size       # box edges are C floats
    filterscale = max(scale, 1.0)
    support = 3.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = _lanczos((np.arange(xmax) + xmin - center + 0.5) * ss)
        ww = 0.0
        for v in w:                      # sequential sum, as the C loop
            ww += v
        if ww != 0.0:
            w = w / ww
        fixed = np.where(w < 0, -0.5 + w * (1 << PIL_PRECISION_BITS), 0.5 + w * (1 << PIL_PRECISION_BITS))
        kk[xx, :xmax] = np.trunc(fixed).astype(np.int32)
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _clip8(acc):
    """Resample.c clip8: lookup of (acc >> PRECISION_BITS) clamped to [0, 255] (arithmetic shift = floor)."""
    return np.clip(acc >> PIL_PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_lanczos_u8(img, out_hw):
    """PIL ``Image.resize((W, H), Image.ANTIALIAS)`` of a single-band 8-bit image (ImagingResampleInner: horizontal
    pass over the rows the vertical pass needs, 8-bit intermediate, then the vertical pass).  img: H x W uint8."""
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim == 2
    h_out, w_out = out_hw
    h_in, w_in = img.shape
    cur = img
    half = 1 << (PIL_PRECISION_BITS - 1)
    if w_out != w_in:
        _, bounds, kk = lanczos_coeffs_8bpc(w_in, w_out)
        tmp = np.empty((h_in, w_out), np.uint8)
        for xx in range(w_out):
            x0, cnt = bounds[xx]
            acc = half + (cur[:, x0:x0 + cnt].astype(np.int64) * kk[xx, :cnt].astype(np.int64)).sum(axis=1)
            tmp[:, xx] = _clip8(acc)
        cur = tmp
    if h_out != h_in:
        _, bounds, kk = lanczos_coeffs_8bpc(h_in, h_out)
        out = np.empty((h_out, cur.shape[1]), np.uint8)
        for yy in range(h_out):
            y0, cnt = bounds[yy]
            acc = half + (cur[y0:y0 + cnt].astype(np.int64) * kk[yy, :cnt, None].astype(np.int64)).sum(axis=0)
            out[yy] = _clip8(acc)
        cur = out
    return cur if cur is not img else img.copy()


def loader_front_end(pol00, pol01, pol10, pol11, out_hw, n=1.5, flip=False):
    """indoor_dataset.py:335-349 + get_xolp (:430-442) + get_normals (pre_encoders.py:99-113) for one sample:
    optional left-right flip (hammer_dataset.py:72-73), Lanczos resize of the four gray images, stack in the order
    (im00, im01, im10, im11) = (0, 45, 90, 135 deg), closed-form XOLP, three normal candidates.
    Returns (planes u8 [4, H, W] in angle order, xolp f64 [2, H, W], normals f64 [9, H, W])."""
    planes = []
    for im in (pol00, pol01, pol10, pol11):
        im = np.asarray(im)
        if flip:
            im = im[:, ::-1]
        planes.append(resize_lanczos_u8(np.ascontiguousarray(im), out_hw))
    stack = np.stack(planes, axis=2)
    _, rho, phi = iun_and_xolp_closed(stack)
    xolp = np.stack((rho, phi))
    return np.stack(planes), xolp, get_normals(xolp[None], n)[0]


# --------------------------------------------------------------------------------------
# whole chain (polarisation/xolp_and_normals.py:101-129) -- used for the CPU baseline timing
# --------------------------------------------------------------------------------------
def frame_chain_reference(mosaic, n=1.5, angles=CANONICAL_ANGLES):
    """split -> stack -> lstsq XOLP -> three table inversions -> three normal maps, one frame.

    This is the reference's own operation sequence (xolp_and_normals.py:107-121) with scipy's
    interp1d replaced by its restatement above; it is what `cpu_baseline` times.
    Returns (iun, rho, phi, normals[9, Hs, Ws]).
    """
    stack = stack_quadrants(mosaic)
    iun, rho, phi = iun_and_xolp_lstsq(stack, angles)
    th_d = rho_diffuse(rho, n)
    th_1, th_2 = rho_spec(rho, n)
    nd = calc_normals_hw3(phi, th_d)
    n1 = calc_normals_hw3(phi + np.pi / 2, th_1)
    n2 = calc_normals_hw3(phi + np.pi / 2, th_2)
    normals = np.concatenate((nd, n1, n2), axis=2).transpose(2, 0, 1)
    return iun, rho, phi, normals
