"""Tensor-level operators: torch CUDA tensors in, torch CUDA tensors out, through the C ABI.

PyTorch is plumbing here (device memory, streams); every arithmetic step runs in libpolcue.so.
Each function validates dtype / contiguity / device, allocates its outputs on the input's device,
enqueues on the current CUDA stream and raises on a non-zero return code.  No fallbacks.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np
import torch

from . import _lib

CANONICAL_ANGLES = np.array([0.0, 45.0, 90.0, 135.0]) * np.pi / 180.0
METRIC_NAMES = ("abs_rel", "sq_rel", "rmse", "rmse_log", "a1", "a2", "a3")

_lut_cache = {}
_lut_lock = threading.Lock()


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _need_cuda(t, name, dtype=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name} must be a CUDA tensor (polcue has no CPU path)")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must have dtype {dtype}, got {t.dtype}")
    return t.contiguous()


_trig = threading.local()


class trig:
    """``with ops.trig("poly"):`` -- the zenith-angle sincos variant of the calls issued by this thread inside the block:
    "mufu" (MUFU sin/cos, 3.6e-7 abs, the default) or "poly" (polynomial, 1.4e-7 abs).  It only selects WHICH table handle
    `lut_for` hands to the library (the mode is a property of the handle, polcue_lut_set_trig); nothing process-global."""

    def __init__(self, mode):
        if mode not in ("mufu", "poly"):
            raise ValueError("trig mode must be 'mufu' or 'poly'")
        self.mode = mode

    def __enter__(self):
        self.saved = getattr(_trig, "mode", "mufu")
        _trig.mode = self.mode
        return self

    def __exit__(self, *exc):
        _trig.mode = self.saved
        return False


def set_default_trig(mode):
    """The sincos variant of every later call of THIS thread outside a `with ops.trig(..)` block ("mufu" or "poly")."""
    if mode not in ("mufu", "poly"):
        raise ValueError("trig mode must be 'mufu' or 'poly'")
    _trig.mode = mode


def lut_for(n, device, mode=None):
    """Zenith-angle tables for refractive index `n` on `device` (cached; built once per (n, device, trig mode))."""
    device = torch.device(device)
    mode = mode or getattr(_trig, "mode", "mufu")
    key = (float(n), device.index if device.index is not None else torch.cuda.current_device(), mode)
    with _lut_lock:
        h = _lut_cache.get(key)
        if h is None:
            with torch.cuda.device(key[1]):
                out = C.c_void_p()
                _lib.check(_lib.lib().polcue_lut_create(float(n), C.byref(out)), f"polcue_lut_create(n={n})")
                _lib.check(_lib.lib().polcue_lut_set_trig(out, 1 if mode == "mufu" else 0), "polcue_lut_set_trig")
            h = _lut_cache[key] = out
    return h


# ------------------------------------------------------------------------------------------
# pinned host buffers on the GPU's NUMA node (for the *_host entry points)
# ------------------------------------------------------------------------------------------
class _HostBlock:
    """Owns one polcue_host_alloc_on block; freed when the last tensor viewing it is gone."""

    def __init__(self, nbytes, device_index):
        self.ptr = C.c_void_p()
        _lib.check(_lib.lib().polcue_host_alloc_on(C.byref(self.ptr), max(int(nbytes), 1), int(device_index)), "polcue_host_alloc_on")
        self.nbytes = int(nbytes)

    def __del__(self):
        try:
            if self.ptr:
                _lib.lib().polcue_host_free(self.ptr)
        except Exception:
            pass


def host_empty(shape, dtype=torch.float32, device=None):
    """Zero-filled pinned CPU tensor whose pages live on the NUMA node of `device` (default: the
    current CUDA device) -- the buffers to hand to `fused_mosaic_host` / `loader_front_end_host`."""
    index = torch.cuda.current_device() if device is None else torch.device(device).index
    numel = 1
    for d in shape:
        numel *= int(d)
    nbytes = numel * torch.empty((), dtype=dtype).element_size()
    block = _HostBlock(nbytes, index if index is not None else torch.cuda.current_device())
    raw = (C.c_uint8 * max(nbytes, 1)).from_address(block.ptr.value)
    raw._owner = block               # torch.frombuffer keeps `raw` alive as long as the storage lives (views included)
    return torch.frombuffer(raw, dtype=dtype, count=numel).reshape(tuple(int(d) for d in shape))


def _host_out(out, key, shape, dtype):
    """A caller-supplied host output buffer, validated (shape, dtype, contiguity, CPU), or a new pinned one."""
    t = out.get(key)
    if t is None:
        t = out[key] = host_empty(shape, dtype)
    elif (not isinstance(t, torch.Tensor) or t.is_cuda or tuple(t.shape) != tuple(shape) or t.dtype != dtype
          or not t.is_contiguous()):
        raise ValueError(f"preallocated host `{key}` must be a contiguous CPU {dtype} tensor of shape {tuple(shape)}")
    return t


# ------------------------------------------------------------------------------------------
# quadrant split
# ------------------------------------------------------------------------------------------
def _split(img, b, h, w, tail):
    if h % 2 or w % 2:
        raise ValueError("array split does not result in an equal division")   # numpy's message (np.split)
    px_bytes = img.element_size()
    for t in tail:
        px_bytes *= t
    lead = (b,) if img.dim() == 2 + len(tail) + 1 else ()
    outs = [torch.empty(lead + (h // 2, w // 2) + tuple(tail), dtype=img.dtype, device=img.device) for _ in range(4)]
    with torch.cuda.device(img.device):
        _lib.check(_lib.lib().polcue_split_pol(_ptr(img), b, h, w, px_bytes, *(_ptr(o) for o in outs), _stream(img)),
                   "polcue_split_pol")
    return tuple(outs)


def split_pol(img):
    """One image H x W (x C...) -> (im00, im10, im01, im11), as pol_split_and_save.py:10-27 (axes 0 and 1)."""
    img = _need_cuda(img, "img")
    if img.dim() < 2:
        raise ValueError("img must have at least two dimensions")
    return _split(img, 1, img.shape[0], img.shape[1], tuple(img.shape[2:]))


def split_pol_batch(imgs):
    """B x H x W (x C...) -> four B x H/2 x W/2 (x C...) tensors in the order of `split_pol`."""
    imgs = _need_cuda(imgs, "imgs")
    if imgs.dim() < 3:
        raise ValueError("imgs must be B x H x W (x C)")
    return _split(imgs, imgs.shape[0], imgs.shape[1], imgs.shape[2], tuple(imgs.shape[3:]))


# ------------------------------------------------------------------------------------------
# fused pipeline
# ------------------------------------------------------------------------------------------
def fused_mosaic(mosaic, n=1.5, want_iun=False, want_planes=False, want_normals=True, out=None, superpixel=None, want_stats=False):
    """B x H x W uint8 mosaics -> dict(xolp [B,2,Hs,Ws], normals [B,9,Hs,Ws], iun, planes).

    `out` may hold preallocated tensors under the same keys (steady-state loops reuse them).
    `want_stats`: adds `stats13` (float64 [13] on device: sum rho, sum phi, the sums of the nine normal channels, sum rho^2,
    sum phi^2 over the batch) as a by-product of the same launch -- the dataset statistics of xolp_mean_and_std_dev.py and
    the output checksums of the sequence benchmark without reading the outputs again; bit-identical to `channel_stats` of
    the stored tensors (`stats_from_stats13` splits it the same way).
    `superpixel`: None for the reference's quadrant-tiled images; for a raw interleaved 2x2 polarizer mosaic the four
    angle indices (0..3 = 0/45/90/135 deg) at positions (0,0), (0,1), (1,0), (1,1), e.g. (2, 1, 3, 0).
    """
    mosaic = _need_cuda(mosaic, "mosaic", torch.uint8)
    if mosaic.dim() == 2:
        mosaic = mosaic[None]
    b, h, w = mosaic.shape
    if h % 2 or w % 2:
        raise ValueError("mosaic height and width must be even")
    hs, ws = h // 2, w // 2
    out = dict(out or {})
    dev = mosaic.device

    def buf(key, shape, dtype):
        t = out.get(key)
        if t is None:
            t = out[key] = torch.empty(shape, dtype=dtype, device=dev)
        elif tuple(t.shape) != tuple(shape) or t.dtype != dtype or not t.is_contiguous() or t.device != dev:
            raise ValueError(f"preallocated `{key}` has the wrong shape/dtype/device")
        return t

    xolp = buf("xolp", (b, 2, hs, ws), torch.float32)
    normals = buf("normals", (b, 9, hs, ws), torch.float32) if want_normals else None
    iun = buf("iun", (b, hs, ws), torch.float32) if want_iun else None
    planes = buf("planes", (b, 4, hs, ws), torch.uint8) if want_planes else None
    lut = lut_for(n, dev) if want_normals else C.c_void_p(0)
    if want_stats and (superpixel is not None or not want_normals):
        raise ValueError("want_stats needs the quadrant layout and the normals output")
    with torch.cuda.device(dev):
        if want_stats:
            need = int(_lib.lib().polcue_fused_stats_workspace_bytes(b, h, w))
            key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
            ws_t = _fused_stat_workspaces.get(key)
            if ws_t is None or ws_t.numel() < need:
                ws_t = _fused_stat_workspaces[key] = torch.zeros(need, dtype=torch.uint8, device=dev)
            stats = out.get("stats13")
            if stats is None or stats.dtype != torch.float64 or tuple(stats.shape) != (13,) or stats.device != dev:
                stats = out["stats13"] = torch.empty(13, dtype=torch.float64, device=dev)
            _lib.check(_lib.lib().polcue_fused_mosaic_stats_u8(_ptr(mosaic), b, h, w, lut, _ptr(planes), _ptr(iun), _ptr(xolp),
                                                               _ptr(normals), _ptr(ws_t), _ptr(stats), _stream(mosaic)),
                       "polcue_fused_mosaic_stats_u8")
        elif superpixel is None:
            _lib.check(_lib.lib().polcue_fused_mosaic_u8(_ptr(mosaic), b, h, w, lut, _ptr(planes), _ptr(iun), _ptr(xolp),
                                                         _ptr(normals), _stream(mosaic)), "polcue_fused_mosaic_u8")
        else:
            if sorted(int(a) for a in superpixel) != [0, 1, 2, 3]:
                raise ValueError("superpixel must list each angle index 0..3 exactly once")
            arr = (C.c_int * 4)(*[int(a) for a in superpixel])
            _lib.check(_lib.lib().polcue_fused_superpixel_u8(_ptr(mosaic), b, h, w, arr, lut, _ptr(planes), _ptr(iun), _ptr(xolp),
                                                             _ptr(normals), _stream(mosaic)), "polcue_fused_superpixel_u8")
    return out


_fused_stat_workspaces = {}


def stats_from_stats13(stats13):
    """(xolp_stats [2, 2] = (sum, sum of squares) of rho and phi, normal_sums [9]) from the by-product vector."""
    return torch.stack((stats13[:2], stats13[11:13]), dim=1), stats13[2:11]


def fused_planes(i0, i45, i90, i135, n=1.5, want_iun=False, want_normals=True, out=None):
    """Four (B x) H x W uint8 planes (indoor_dataset.py:435-438) -> dict(xolp [B,2,H,W], normals [B,9,H,W], iun) in one launch:
    the loader's get_xolp plus the encoder's get_normals."""
    planes = [_need_cuda(p, "plane", torch.uint8) for p in (i0, i45, i90, i135)]
    shape = planes[0].shape
    if any(p.shape != shape or p.device != planes[0].device for p in planes) or len(shape) not in (2, 3):
        raise ValueError("planes must share one (B x) H x W shape and device")
    b = shape[0] if len(shape) == 3 else 1
    h, w = shape[-2:]
    dev = planes[0].device
    out = dict(out or {})

    def buf(key, shp):
        t = out.get(key)
        if t is None:
            t = out[key] = torch.empty(shp, dtype=torch.float32, device=dev)
        elif tuple(t.shape) != tuple(shp) or t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev:
            raise ValueError(f"preallocated `{key}` has the wrong shape/dtype/device")
        return t

    xolp = buf("xolp", (b, 2, h, w))
    normals = buf("normals", (b, 9, h, w)) if want_normals else None
    iun = buf("iun", (b, h, w)) if want_iun else None
    lut = lut_for(n, dev) if want_normals else C.c_void_p(0)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().polcue_fused_planes_u8(*(_ptr(p) for p in planes), b, h, w, lut, _ptr(iun), _ptr(xolp), _ptr(normals),
                                                     _stream(planes[0])), "polcue_fused_planes_u8")
    return out


# ------------------------------------------------------------------------------------------
# loader front end: Pillow-exact Lanczos resize (+ XOLP + normals)
# ------------------------------------------------------------------------------------------
_plan_cache = {}


def resize_plan(in_hw, out_hw, device):
    """Lanczos weights for one (in_h, in_w) -> (out_h, out_w) geometry on `device` (cached)."""
    device = torch.device(device)
    key = (int(in_hw[0]), int(in_hw[1]), int(out_hw[0]), int(out_hw[1]),
           device.index if device.index is not None else torch.cuda.current_device())
    with _lut_lock:
        h = _plan_cache.get(key)
        if h is None:
            with torch.cuda.device(key[4]):
                out = C.c_void_p()
                _lib.check(_lib.lib().polcue_resize_plan_create(*key[:4], C.byref(out)), f"polcue_resize_plan_create{key[:4]}")
            h = _plan_cache[key] = out
    return h


def _flip_flags(flip, count, device):
    if flip is None:
        return None
    if isinstance(flip, bool):
        flip = [flip] * count
    t = torch.as_tensor(flip).to(device=device, dtype=torch.uint8).contiguous()
    if t.numel() != count:
        raise ValueError(f"flip needs one flag per image ({count}), got {t.numel()}")
    return t


def lanczos_resize(images, out_hw, flip=None, workspace=None):
    """uint8 (N x ... x) H x W -> same leading dims x out_h x out_w, bit-exact with PIL ``Image.resize(.., ANTIALIAS)`` of an
    'L' image -- ``transforms.Resize((height, width), Image.ANTIALIAS)``, indoor_dataset.py:115.  `flip`: bool or one flag per
    image, mirrors left-right first (hammer_dataset.py:72-73)."""
    images = _need_cuda(images, "images", torch.uint8)
    if images.dim() < 2:
        raise ValueError("images must be (... x) H x W")
    lead, (h, w) = tuple(images.shape[:-2]), images.shape[-2:]
    count = 1
    for d in lead:
        count *= d
    plan = resize_plan((h, w), out_hw, images.device)
    flags = _flip_flags(flip, count, images.device)
    need = _lib.lib().polcue_resize_workspace_bytes(plan, count)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(max(need, 1), dtype=torch.uint8, device=images.device)
    out = torch.empty(lead + (int(out_hw[0]), int(out_hw[1])), dtype=torch.uint8, device=images.device)
    with torch.cuda.device(images.device):
        _lib.check(_lib.lib().polcue_resize_lanczos_u8(plan, _ptr(images), count, _ptr(flags), _ptr(workspace), _ptr(out),
                                                       _stream(images)), "polcue_resize_lanczos_u8")
    return out


XOLP_MEAN_STD = (0.08693199701957657, 0.44430732785457433)     # ShallowEncoder.normalizeInput, pre_encoders.py:79


def loader_front_end(i0, i45, i90, i135, out_hw, n=1.5, flip=None, want_iun=False, want_normals=True, out=None, normalize_xolp=None):
    """The polarization branch of ``__getitem__`` for a batch, on the GPU: four full-resolution uint8 images (B x H x W each;
    pol00, pol01, pol10, pol11 = 0, 45, 90, 135 deg) -> dict(planes u8 [B,4,h,w] (the resized images), xolp [B,2,h,w],
    normals [B,9,h,w], iun).  indoor_dataset.py:335-349 + get_xolp (:430-442) + get_normals (pre_encoders.py:99-113).
    `normalize_xolp`: (mean, std), e.g. XOLP_MEAN_STD, adds `xolp_norm` = ShallowEncoder.normalizeInput(xolp, 'XOLP')
    (pre_encoders.py:75-83), written by the same kernel."""
    planes = [_need_cuda(p, "image", torch.uint8) for p in (i0, i45, i90, i135)]
    shape = planes[0].shape
    if any(p.shape != shape or p.device != planes[0].device for p in planes) or len(shape) != 3:
        raise ValueError("images must share one B x H x W shape and device")
    b, h, w = shape
    oh, ow = int(out_hw[0]), int(out_hw[1])
    dev = planes[0].device
    out = dict(out or {})

    def buf(key, shp, dtype=torch.float32):
        t = out.get(key)
        if t is None:
            t = out[key] = torch.empty(shp, dtype=dtype, device=dev)
        elif tuple(t.shape) != tuple(shp) or t.dtype != dtype or not t.is_contiguous() or t.device != dev:
            raise ValueError(f"preallocated `{key}` has the wrong shape/dtype/device")
        return t

    plan = resize_plan((h, w), (oh, ow), dev)
    small = buf("planes", (b, 4, oh, ow), torch.uint8)
    xolp = buf("xolp", (b, 2, oh, ow))
    normals = buf("normals", (b, 9, oh, ow)) if want_normals else None
    iun = buf("iun", (b, oh, ow)) if want_iun else None
    need = _lib.lib().polcue_resize_workspace_bytes(plan, 4 * b)
    ws = out.get("workspace")
    if ws is None or ws.numel() < need or ws.device != dev:
        ws = out["workspace"] = torch.empty(max(need, 1), dtype=torch.uint8, device=dev)
    flags = _flip_flags(flip, b, dev)
    lut = lut_for(n, dev) if want_normals else C.c_void_p(0)
    xnorm = buf("xolp_norm", (b, 2, oh, ow)) if normalize_xolp is not None else None
    mean_std = (C.c_float * 2)(*[float(v) for v in normalize_xolp]) if normalize_xolp is not None else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().polcue_loader_front_end_u8(plan, *(_ptr(p) for p in planes), b, _ptr(flags), lut, _ptr(ws), _ptr(small),
                                                         _ptr(iun), _ptr(xolp), _ptr(normals), mean_std, _ptr(xnorm),
                                                         _stream(planes[0])),
                   "polcue_loader_front_end_u8")
    return out


def fused_mosaic_host(mosaic, n=1.5, want_iun=False, want_normals=True, out=None, chunk_frames=0, device=None):
    """Host (ideally pinned) uint8 mosaics -> host float32 outputs; copies are pipelined inside the library."""
    if not isinstance(mosaic, torch.Tensor) or mosaic.is_cuda or mosaic.dtype != torch.uint8:
        raise TypeError("mosaic must be a CPU uint8 tensor")
    mosaic = mosaic.contiguous()
    if mosaic.dim() == 2:
        mosaic = mosaic[None]
    b, h, w = mosaic.shape
    hs, ws = h // 2, w // 2
    out = dict(out or {})

    if h % 2 or w % 2:
        raise ValueError("mosaic height and width must be even")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(dev):
        xolp = _host_out(out, "xolp", (b, 2, hs, ws), torch.float32)
        normals = _host_out(out, "normals", (b, 9, hs, ws), torch.float32) if want_normals else None
        iun = _host_out(out, "iun", (b, hs, ws), torch.float32) if want_iun else None
    lut = lut_for(n, dev) if want_normals else C.c_void_p(0)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().polcue_fused_mosaic_u8_host(_ptr(mosaic), b, h, w, lut, _ptr(iun), _ptr(xolp), _ptr(normals),
                                                          int(chunk_frames)), "polcue_fused_mosaic_u8_host")
    return out


def loader_front_end_host(i0, i45, i90, i135, out_hw, n=1.5, flip=None, want_planes=True, want_normals=True, normalize_xolp=None,
                          out=None, chunk_samples=0, device=None):
    """`loader_front_end` for HOST tensors (ideally pinned): four B x H x W uint8 CPU stacks in, host outputs; the copies are
    pipelined against the kernels inside the library (`polcue_loader_front_end_u8_host`)."""
    imgs = []
    for t in (i0, i45, i90, i135):
        if not isinstance(t, torch.Tensor) or t.is_cuda or t.dtype != torch.uint8:
            raise TypeError("images must be CPU uint8 tensors")
        imgs.append(t.contiguous())
    shape = imgs[0].shape
    if any(t.shape != shape for t in imgs) or len(shape) != 3:
        raise ValueError("images must share one B x H x W shape")
    b, h, w = shape
    oh, ow = int(out_hw[0]), int(out_hw[1])
    out = dict(out or {})

    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(dev):
        xolp = _host_out(out, "xolp", (b, 2, oh, ow), torch.float32)
        planes = _host_out(out, "planes", (b, 4, oh, ow), torch.uint8) if want_planes else None
        normals = _host_out(out, "normals", (b, 9, oh, ow), torch.float32) if want_normals else None
        xnorm = _host_out(out, "xolp_norm", (b, 2, oh, ow), torch.float32) if normalize_xolp is not None else None
    mean_std = (C.c_float * 2)(*[float(v) for v in normalize_xolp]) if normalize_xolp is not None else None
    flags = None
    if flip is not None:
        flags = torch.as_tensor([flip] * b if isinstance(flip, bool) else flip).to(torch.uint8).contiguous()
        if flags.numel() != b:
            raise ValueError(f"flip needs one flag per sample ({b})")
    lut = lut_for(n, dev) if want_normals else C.c_void_p(0)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().polcue_loader_front_end_u8_host(h, w, oh, ow, *(_ptr(t) for t in imgs), b, _ptr(flags), lut, _ptr(planes),
                                                              _ptr(xolp), _ptr(normals), mean_std, _ptr(xnorm), int(chunk_samples)),
                   "polcue_loader_front_end_u8_host")
    return out


# ------------------------------------------------------------------------------------------
# XOLP
# ------------------------------------------------------------------------------------------
def pinv_for_angles(angles):
    """None for the canonical angles, else the 3x4 pseudo-inverse lstsq applies (xolp.py:15-21)."""
    angles = np.asarray(angles, dtype=np.float64).reshape(4)
    if np.allclose(angles, CANONICAL_ANGLES, rtol=0, atol=1e-12):
        return None
    design = np.stack((np.ones(4), np.cos(2 * angles), np.sin(2 * angles)), axis=1)
    return np.linalg.pinv(design)


def xolp_from_stack(stack, angles=None, want_iun=True):
    """(B x) H x W x 4 uint8 / float32 stack -> (iun [B,H,W] or None, xolp [B,2,H,W]); xolp.py:8-34."""
    stack = _need_cuda(stack, "stack")
    if stack.dtype not in (torch.uint8, torch.float32):
        stack = stack.float()
    if stack.dim() == 3:
        stack = stack[None]
    if stack.dim() != 4 or stack.shape[-1] != 4:
        raise ValueError("stack must be (B x) H x W x 4")
    b, h, w, _ = stack.shape
    pinv = None if angles is None else pinv_for_angles(angles)
    pinv_arg = None if pinv is None else (C.c_float * 12)(*[float(v) for v in pinv.reshape(-1)])
    iun = torch.empty((b, h, w), dtype=torch.float32, device=stack.device) if want_iun else None
    xolp = torch.empty((b, 2, h, w), dtype=torch.float32, device=stack.device)
    fn = _lib.lib().polcue_xolp_stack_u8 if stack.dtype == torch.uint8 else _lib.lib().polcue_xolp_stack_f32
    with torch.cuda.device(stack.device):
        _lib.check(fn(_ptr(stack), b, h, w, pinv_arg, _ptr(iun), _ptr(xolp), _stream(stack)), "polcue_xolp_stack")
    return iun, xolp


def xolp_from_planes(i0, i45, i90, i135, want_iun=False):
    """Four B x H x W uint8 planes (indoor_dataset.py:435-438) -> (iun or None, xolp [B,2,H,W])."""
    planes = [_need_cuda(p, "plane", torch.uint8) for p in (i0, i45, i90, i135)]
    shape = planes[0].shape
    if any(p.shape != shape for p in planes) or len(shape) not in (2, 3):
        raise ValueError("planes must share one (B x) H x W shape")
    b = shape[0] if len(shape) == 3 else 1
    h, w = shape[-2:]
    dev = planes[0].device
    iun = torch.empty((b, h, w), dtype=torch.float32, device=dev) if want_iun else None
    xolp = torch.empty((b, 2, h, w), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().polcue_xolp_planes_u8(*(_ptr(p) for p in planes), b, h, w, _ptr(iun), _ptr(xolp),
                                                    _stream(planes[0])), "polcue_xolp_planes_u8")
    return iun, xolp


# ------------------------------------------------------------------------------------------
# physics normals
# ------------------------------------------------------------------------------------------
def get_normals(x, n=1.5):
    """B x 2 x H x W (rho, phi) -> B x 9 x H x W float32; pre_encoders.py:99-113."""
    x = _need_cuda(x, "x")
    if x.dim() != 4 or x.shape[1] != 2:
        raise ValueError("x must be B x 2 x H x W")
    x = x.float().contiguous()
    b, _, h, w = x.shape
    out = torch.empty((b, 9, h, w), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().polcue_normals_from_xolp_f32(_ptr(x), b, h, w, lut_for(n, x.device), _ptr(out), _stream(x)),
                   "polcue_normals_from_xolp_f32")
    return out


def rho_diffuse(rho, n):
    rho = _need_cuda(rho, "rho").float().contiguous()
    theta = torch.empty_like(rho)
    with torch.cuda.device(rho.device):
        _lib.check(_lib.lib().polcue_rho_diffuse_f32(_ptr(rho), rho.numel(), lut_for(n, rho.device), _ptr(theta), _stream(rho)),
                   "polcue_rho_diffuse_f32")
    return theta


def rho_spec(rho, n):
    rho = _need_cuda(rho, "rho").float().contiguous()
    t1, t2 = torch.empty_like(rho), torch.empty_like(rho)
    with torch.cuda.device(rho.device):
        _lib.check(_lib.lib().polcue_rho_spec_f32(_ptr(rho), rho.numel(), lut_for(n, rho.device), _ptr(t1), _ptr(t2),
                                                  _stream(rho)), "polcue_rho_spec_f32")
    return t1, t2


def calc_normals(phi, theta):
    """(phi, theta) B x H x W -> B x 3 x H x W; normals_vec.py:53-60."""
    phi = _need_cuda(phi, "phi").float().contiguous()
    theta = _need_cuda(theta, "theta").float().contiguous()
    if phi.shape != theta.shape or phi.dim() < 2:
        raise ValueError("phi and theta must share a B x ... shape")
    b = phi.shape[0]
    hw = phi.numel() // b if b else 0
    out = torch.empty((b, 3) + tuple(phi.shape[1:]), dtype=torch.float32, device=phi.device)
    with torch.cuda.device(phi.device):
        _lib.check(_lib.lib().polcue_calc_normals_f32(_ptr(phi), _ptr(theta), b, hw, _ptr(out), _stream(phi)),
                   "polcue_calc_normals_f32")
    return out


def stokes_channel(stack, mask):
    """H x W x 4 float stack + H x W mask -> (rho, phi, iun); physical_normals_channels.py:15-36."""
    stack = _need_cuda(stack, "stack").float().contiguous()
    mask = _need_cuda(mask, "mask").to(torch.uint8).contiguous()
    h, w, _ = stack.shape
    rho, phi, iun = (torch.empty((h, w), dtype=torch.float32, device=stack.device) for _ in range(3))
    with torch.cuda.device(stack.device):
        _lib.check(_lib.lib().polcue_stokes_channel_f32(_ptr(stack), _ptr(mask), h, w, _ptr(rho), _ptr(phi), _ptr(iun),
                                                        _stream(stack)), "polcue_stokes_channel_f32")
    return rho, phi, iun


def calc_normals_channel(phi, theta, mask, phi_offset=0.0):
    """H x W x 3, zero outside mask; physical_normals_channels.py:75-83 (phi_offset: the caller's phi + pi/2)."""
    phi = _need_cuda(phi, "phi").float().contiguous()
    theta = _need_cuda(theta, "theta").float().contiguous()
    mask = _need_cuda(mask, "mask").to(torch.uint8).contiguous()
    out = torch.empty(tuple(phi.shape) + (3,), dtype=torch.float32, device=phi.device)
    with torch.cuda.device(phi.device):
        _lib.check(_lib.lib().polcue_calc_normals_channel_f32(_ptr(phi), _ptr(theta), _ptr(mask), phi.numel(), float(phi_offset),
                                                              _ptr(out), _stream(phi)), "polcue_calc_normals_channel_f32")
    return out


# ------------------------------------------------------------------------------------------
# depth -> normals
# ------------------------------------------------------------------------------------------
def depth_to_normals(depth, camera_matrix):
    """depth B x 1 x H x W, camera_matrix B x 3 x 3 -> B x 3 x H x W (kornia 0.5.11 semantics, forward only)."""
    depth = _need_cuda(depth, "depth")
    if depth.requires_grad:
        raise NotImplementedError(
            "polcue.depth_to_normals is forward-only (the GT branch of trainer.py:1305); the predicted-depth branch "
            "needs autograd and is listed as the next step in DESIGN.md")
    if depth.dim() != 4 or depth.shape[1] != 1:
        raise ValueError("depth must be B x 1 x H x W")
    camera_matrix = _need_cuda(camera_matrix, "camera_matrix")
    if camera_matrix.shape != (depth.shape[0], 3, 3):
        raise ValueError("camera_matrix must be B x 3 x 3")
    depth = depth.float().contiguous()
    k = camera_matrix.float().contiguous()
    b, _, h, w = depth.shape
    out = torch.empty((b, 3, h, w), dtype=torch.float32, device=depth.device)
    with torch.cuda.device(depth.device):
        _lib.check(_lib.lib().polcue_depth_to_normals_f32(_ptr(depth), _ptr(k), b, h, w, _ptr(out), _stream(depth)),
                   "polcue_depth_to_normals_f32")
    return out


# ------------------------------------------------------------------------------------------
# depth metrics
# ------------------------------------------------------------------------------------------
_workspaces = {}


def _workspace(device):
    """Per-(device, stream) scratch for the flat reduction (ticket word + per-CTA partials)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        nbytes = int(_lib.lib().polcue_depth_errors_workspace_bytes())
        ws = _workspaces[key] = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    return ws


def depth_error_sums(gt, pred, want_metrics=True):
    """Flat (already masked) gt / pred -> (sums8 float64 [8], metrics float32 [7] or None), on device."""
    gt = _need_cuda(gt, "gt").float().contiguous().reshape(-1)
    pred = _need_cuda(pred, "pred").float().contiguous().reshape(-1)
    if gt.numel() != pred.numel():
        raise ValueError("gt and pred must have the same number of elements")
    sums = torch.empty(8, dtype=torch.float64, device=gt.device)
    metrics = torch.empty(7, dtype=torch.float32, device=gt.device) if want_metrics else None
    with torch.cuda.device(gt.device):
        _lib.check(_lib.lib().polcue_depth_errors_f32(_ptr(gt), _ptr(pred), gt.numel(), _ptr(_workspace(gt.device)), _ptr(sums),
                                                      _ptr(metrics), _stream(gt)), "polcue_depth_errors_f32")
    return sums, metrics


def compute_depth_errors(gt, pred):
    """manydepth/layers.py:539-557: 7 zero-dim float32 tensors (abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3)."""
    _, metrics = depth_error_sums(gt, pred)
    return tuple(metrics.unbind(0))


def _inst_range(inst_id):
    """None -> no filter; int -> that id; (lo, hi) -> the id range (evaluation.py:259-262 "objects" = (20, 160))."""
    if inst_id is None:
        return None
    if isinstance(inst_id, (tuple, list)):
        lo, hi = int(inst_id[0]), int(inst_id[1])
    else:
        lo = hi = int(inst_id)
    return lo, hi


def masked_median_scale(gt, pred, min_depth, max_depth, inst=None, inst_id=None, clamp_first=False):
    """Per image: (medians [B,2] = np.median(gt[mask]), np.median(pred[mask]); scale [B] = their float32 ratio), the factor of
    the reference's median scaling (trainer.py:1413-1414).  Same mask as `depth_errors_per_image`."""
    gt = _need_cuda(gt, "gt").float().contiguous()
    pred = _need_cuda(pred, "pred").float().contiguous()
    b = gt.shape[0]
    px = gt.numel() // b if b else 0
    if pred.numel() != gt.numel():
        raise ValueError("gt and pred must have the same shape")
    rng = _inst_range(inst_id) if inst is not None else None
    use_inst = rng is not None
    if use_inst:
        inst = _need_cuda(inst, "inst", torch.uint8)
    lo, hi = rng if use_inst else (0, 0)
    medians = torch.empty((b, 2), dtype=torch.float32, device=gt.device)
    scale = torch.empty((b,), dtype=torch.float32, device=gt.device)
    with torch.cuda.device(gt.device):
        _lib.check(_lib.lib().polcue_masked_median_scale_f32(_ptr(gt), _ptr(pred), _ptr(inst) if use_inst else C.c_void_p(0), b, px,
                                                             float(min_depth), float(max_depth), lo, hi, int(bool(clamp_first)),
                                                             _ptr(medians), _ptr(scale), _stream(gt)), "polcue_masked_median_scale_f32")
    return medians, scale


def depth_errors_per_image(gt, pred, min_depth, max_depth, inst=None, inst_id=None, median_scaling=False, clamp_first=False):
    """Per-image masked metrics (trainer.py:1376-1428): gt/pred B x H x W (or B x 1 x H x W), inst uint8 or None.
    `inst_id`: None, one material id, or an inclusive (lo, hi) id range (evaluation.py's "objects" group = (20, 160)).
    `median_scaling`: multiply each image's prediction by median(gt[mask]) / median(pred[mask]) first (trainer.py:1413-1414,
    the configurations without depth supervision); `clamp_first`: clamp the prediction before that too (trainer.py:1368-1370).

    Returns (sums [B,8] float64, metrics [B,7] float32) on device.
    """
    gt = _need_cuda(gt, "gt").float().contiguous()
    pred = _need_cuda(pred, "pred").float().contiguous()
    b = gt.shape[0]
    px = gt.numel() // b if b else 0
    if pred.numel() != gt.numel():
        raise ValueError("gt and pred must have the same shape")
    rng = _inst_range(inst_id) if inst is not None else None
    use_inst = rng is not None
    if use_inst:
        inst = _need_cuda(inst, "inst", torch.uint8)
        if inst.numel() != gt.numel():
            raise ValueError("inst must have the same shape as gt")
    lo, hi = rng if use_inst else (0, 0)
    sums = torch.empty((b, 8), dtype=torch.float64, device=gt.device)
    metrics = torch.empty((b, 7), dtype=torch.float32, device=gt.device)
    scale = (masked_median_scale(gt, pred, min_depth, max_depth, inst if use_inst else None, inst_id, clamp_first)[1]
             if median_scaling else None)
    with torch.cuda.device(gt.device):
        _lib.check(_lib.lib().polcue_depth_errors_images_scaled_f32(_ptr(gt), _ptr(pred), _ptr(inst) if use_inst else C.c_void_p(0), b,
                                                                    px, float(min_depth), float(max_depth), lo, hi,
                                                                    int(bool(clamp_first)), _ptr(scale), _ptr(sums), _ptr(metrics),
                                                                    _stream(gt)), "polcue_depth_errors_images_scaled_f32")
    return sums, metrics


def depth_errors_groups(gt, pred, inst, min_depth, max_depth, group_ids):
    """Every mask group of the reference's evaluation loop in one launch.  group_ids: iterable of instance ids
    (20 ... 200, trainer.py:1389-1411) with None / -1 meaning object == "all".
    Returns (sums [B, G, 8] float64, metrics [B, G, 7] float32) on device.  The accumulators are additive, so a union of
    materials (evaluation.py's "objects" = ids 20..160) is `metrics_from_sums(sums[:, those groups].sum(1))`."""
    gt = _need_cuda(gt, "gt").float().contiguous()
    pred = _need_cuda(pred, "pred").float().contiguous()
    ids = [(-1 if g is None else int(g)) for g in group_ids]
    if not 1 <= len(ids) <= 16:
        raise ValueError("between 1 and 16 groups per launch")
    if any(g >= 0 for g in ids):
        inst = _need_cuda(inst, "inst", torch.uint8)
        if inst.numel() != gt.numel():
            raise ValueError("inst must have the same shape as gt")
    else:
        inst = None
    if pred.numel() != gt.numel():
        raise ValueError("gt and pred must have the same shape")
    b = gt.shape[0]
    px = gt.numel() // b if b else 0
    sums = torch.empty((b, len(ids), 8), dtype=torch.float64, device=gt.device)
    metrics = torch.empty((b, len(ids), 7), dtype=torch.float32, device=gt.device)
    arr = (C.c_int * len(ids))(*ids)
    with torch.cuda.device(gt.device):
        _lib.check(_lib.lib().polcue_depth_errors_groups_f32(_ptr(gt), _ptr(pred), _ptr(inst), b, px, float(min_depth), float(max_depth),
                                                             arr, len(ids), _ptr(sums), _ptr(metrics), _stream(gt)),
                   "polcue_depth_errors_groups_f32")
    return sums, metrics


def eval_pass(gt, pred, inst, camera_matrix, min_depth, max_depth, group_ids, want_normals=True, out=None, peer=None, _prepare_only=False):
    """One evaluation pass over this rank's images with no host work between the launches (BASELINE configs[4]): the GT
    depth->normals stencil, every mask group's per-image metrics, and `mean_acc` = {n_images, sum over images of the
    per-image metrics} (float64 [1 + G * 7]) -- all-reduce it and divide for the reference's mean over images
    (trainer.py:1426).  gt / pred: B x H x W float32, inst: uint8, camera_matrix: B x 3 x 3.
    `out` may hold the tensors of a previous call (a fixed set of buffers makes the pass CUDA-graph capturable).
    `peer` (a `polcue.dist.PeerExchange`): the last kernel also sums `mean_acc` over all ranks through NVLink peer memory
    into `mean_acc_all` -- the collective of a sharded evaluation costs no extra launch; every rank must make the call.
    Returns dict(normals [B,3,H,W] or None, sums [B,G,8], metrics [B,G,7], mean_acc [1 + 7 G] (, mean_acc_all [1 + 7 G]))."""
    gt = _need_cuda(gt, "gt", torch.float32)
    pred = _need_cuda(pred, "pred", torch.float32)
    if gt.dim() != 3 or pred.shape != gt.shape:
        raise ValueError("gt and pred must share one B x H x W shape")
    ids = [(-1 if g is None else int(g)) for g in group_ids]
    if not 1 <= len(ids) <= 16:
        raise ValueError("between 1 and 16 groups per launch")
    if any(g >= 0 for g in ids):
        inst = _need_cuda(inst, "inst", torch.uint8)
        if inst.shape != gt.shape:
            raise ValueError("inst must have the same shape as gt")
    else:
        inst = None
    b, h, w = gt.shape
    dev, g = gt.device, len(ids)
    k = None
    if want_normals:
        k = _need_cuda(camera_matrix, "camera_matrix", torch.float32)
        if k.shape != (b, 3, 3):
            raise ValueError("camera_matrix must be B x 3 x 3")
    out = dict(out or {})

    def buf(key, shape, dtype):
        t = out.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != dev or not t.is_contiguous():
            t = out[key] = torch.empty(shape, dtype=dtype, device=dev)
        return t

    normals = buf("normals", (b, 3, h, w), torch.float32) if want_normals else None
    sums = buf("sums", (b, g, 8), torch.float64)
    metrics = buf("metrics", (b, g, 7), torch.float32)
    acc = buf("mean_acc", (1 + 7 * g,), torch.float64)
    arr = (C.c_int * g)(*ids)
    head = (_ptr(gt), _ptr(pred), _ptr(inst), _ptr(k), b, h, w, float(min_depth), float(max_depth), arr, g, _ptr(normals), _ptr(sums),
            _ptr(metrics), _ptr(acc))
    if peer is not None:
        acc_all = buf("mean_acc_all", (1 + 7 * g,), torch.float64)
        fn, name, args = _lib.lib().polcue_eval_pass_peer_f32, "polcue_eval_pass_peer_f32", head + (peer.handle, _ptr(acc_all))
    else:
        fn, name, args = _lib.lib().polcue_eval_pass_f32, "polcue_eval_pass_f32", head
    if not want_normals:
        out["normals"] = None
    if _prepare_only:
        return out, fn, name, args, (gt, pred, inst, k)
    with torch.cuda.device(dev):
        _lib.check(fn(*args, _stream(gt)), name)
    return out


class EvalPass:
    """`eval_pass` with its arguments validated, its buffers allocated and its C call bound ONCE: `run()` only enqueues the
    three launches (a few microseconds of host time instead of ~40 for the checks, buffer look-ups and context managers of
    `eval_pass`) -- the 120-image evaluation pass of the reference is launch-latency sized, and a shard of it even more so.
    The inputs are captured by reference: refill `gt` / `pred` / `inst` / `camera_matrix` in place between runs.
    `run()` must be called with the inputs' device current (`torch.cuda.set_device`); results are in `.out` (see `eval_pass`)."""

    def __init__(self, gt, pred, inst, camera_matrix, min_depth, max_depth, group_ids, want_normals=True, peer=None):
        self.out, self._fn, self._name, self._args, self._inputs = eval_pass(gt, pred, inst, camera_matrix, min_depth, max_depth, group_ids,
                                                                             want_normals=want_normals, peer=peer, _prepare_only=True)
        self.device = gt.device
        self._peer = peer                                  # keeps the exchange block alive as long as the bound call

    def run(self):
        if torch.cuda.current_device() != self.device.index:
            raise RuntimeError(f"EvalPass.run: make {self.device} the current device first")
        rc = self._fn(*self._args, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if rc != 0:
            _lib.check(rc, self._name)
        return self.out


def metrics_from_sums(sums):
    """8 additive accumulators (any leading shape) -> 7 metrics in the reference order (float64 tensor/array)."""
    n = sums[..., 0]
    sqrt = torch.sqrt if isinstance(sums, torch.Tensor) else np.sqrt
    stack = torch.stack if isinstance(sums, torch.Tensor) else np.stack
    return stack((sums[..., 6] / n, sums[..., 7] / n, sqrt(sums[..., 4] / n), sqrt(sums[..., 5] / n),
                  sums[..., 1] / n, sums[..., 2] / n, sums[..., 3] / n), -1)


# ------------------------------------------------------------------------------------------
# per-channel statistics (XOLP dataset statistics, output checksums)
# ------------------------------------------------------------------------------------------
_stat_workspaces = {}


def channel_stats(x):
    """x: B x C x ... float32 CUDA tensor -> float64 tensor [C, 2] on device: per-channel (sum, sum of squares)."""
    x = _need_cuda(x, "x", torch.float32)
    if x.dim() < 3:
        raise ValueError("x must be B x C x ...")
    b, c = x.shape[0], x.shape[1]
    hw = x.numel() // (b * c) if b * c else 0
    nbytes = int(_lib.lib().polcue_channel_stats_workspace_bytes(b, c, hw))
    key = (x.device.index, torch.cuda.current_stream(x.device).cuda_stream)
    ws = _stat_workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = _stat_workspaces[key] = torch.zeros(nbytes, dtype=torch.uint8, device=x.device)
    stats = torch.empty((c, 2), dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().polcue_channel_stats_f32(_ptr(x), b, c, hw, _ptr(ws), _ptr(stats), _stream(x)),
                   "polcue_channel_stats_f32")
    return stats


def xolp_statistics(xolp, reduce_over_ranks=True, stats13=None):
    """DoLP / AoLP mean and (population) std over a set of frames, as polarisation/xolp_mean_and_std_dev.py:26-32 prints
    them.  xolp: B x 2 x H x W.  With torch.distributed initialised the sums are all-reduced first.
    `stats13`: the by-product of `fused_mosaic(.., want_stats=True)` for the same frames -- the outputs are then not read."""
    from . import dist as D
    b = xolp.shape[0]
    per_channel = channel_stats(xolp) if stats13 is None else stats_from_stats13(stats13)[0]
    acc = torch.cat((per_channel.reshape(-1),
                     torch.tensor([float(xolp.numel() // 2)], dtype=torch.float64, device=xolp.device)))
    if reduce_over_ranks:
        D.all_reduce_sums(acc)
    n = acc[4]
    mean = acc[[0, 2]] / n
    std = torch.sqrt(torch.clamp(acc[[1, 3]] / n - mean * mean, min=0.0))
    return {"dolp_mean": float(mean[0]), "dolp_std": float(std[0]), "aolp_mean": float(mean[1]), "aolp_std": float(std[1]),
            "xolp_mean": float(0.5 * (mean[0] + mean[1])), "xolp_std": float(0.5 * (std[0] + std[1])), "frames": b}


# ------------------------------------------------------------------------------------------
# supervised normals loss (forward + backward)
# ------------------------------------------------------------------------------------------
_loss_workspaces = {}


def _loss_workspace(device):
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _loss_workspaces.get(key)
    if ws is None:
        ws = _loss_workspaces[key] = torch.zeros(int(_lib.lib().polcue_normals_loss_workspace_bytes()), dtype=torch.uint8,
                                                 device=device)
    return ws


class _NormalsLoss(torch.autograd.Function):
    """loss = sum((2 - cos(n_gt, n_pred)) * mask) / sum(mask); gradient flows to depth_pred only."""

    @staticmethod
    def forward(ctx, depth_gt, depth_pred, camera_matrix, mask):
        b, _, h, w = depth_pred.shape
        sums = torch.empty(2, dtype=torch.float64, device=depth_pred.device)
        loss = torch.empty((), dtype=torch.float32, device=depth_pred.device)
        with torch.cuda.device(depth_pred.device):
            _lib.check(_lib.lib().polcue_normals_loss_fwd_f32(_ptr(depth_gt), _ptr(depth_pred), _ptr(camera_matrix), _ptr(mask), b, h, w,
                                                              _ptr(_loss_workspace(depth_pred.device)), _ptr(sums), _ptr(loss),
                                                              _stream(depth_pred)), "polcue_normals_loss_fwd_f32")
        ctx.save_for_backward(depth_gt, depth_pred, camera_matrix, mask, sums)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        depth_gt, depth_pred, camera_matrix, mask, sums = ctx.saved_tensors
        b, _, h, w = depth_pred.shape
        grad_out = grad_out.to(torch.float32).contiguous()
        grad_pred = torch.empty_like(depth_pred)
        with torch.cuda.device(depth_pred.device):
            _lib.check(_lib.lib().polcue_normals_loss_bwd_f32(_ptr(depth_gt), _ptr(depth_pred), _ptr(camera_matrix), _ptr(mask), b, h, w,
                                                              _ptr(sums), _ptr(grad_out), _ptr(grad_pred), _stream(depth_pred)),
                       "polcue_normals_loss_bwd_f32")
        return None, grad_pred, None, None


class _SupervisedLosses(torch.autograd.Function):
    """(depth L1 loss, normals loss) of trainer.py:1240-1251 with mask = (min <= gt <= max); gradients flow to depth_pred."""

    @staticmethod
    def forward(ctx, depth_gt, depth_pred, camera_matrix, min_depth, max_depth):
        b, _, h, w = depth_pred.shape
        sums = torch.empty(3, dtype=torch.float64, device=depth_pred.device)
        losses = torch.empty(2, dtype=torch.float32, device=depth_pred.device)
        with torch.cuda.device(depth_pred.device):
            _lib.check(_lib.lib().polcue_supervised_losses_fwd_f32(_ptr(depth_gt), _ptr(depth_pred), _ptr(camera_matrix), min_depth,
                                                                   max_depth, b, h, w, _ptr(_loss_workspace(depth_pred.device)),
                                                                   _ptr(sums), _ptr(losses), _stream(depth_pred)),
                       "polcue_supervised_losses_fwd_f32")
        ctx.save_for_backward(depth_gt, depth_pred, camera_matrix, sums)
        ctx.range = (min_depth, max_depth)
        return losses[1], losses[0]          # (supervised_depth_loss, supervised_normals_loss)

    @staticmethod
    def backward(ctx, grad_depth, grad_normals):
        depth_gt, depth_pred, camera_matrix, sums = ctx.saved_tensors
        b, _, h, w = depth_pred.shape
        zero = torch.zeros((), dtype=torch.float32, device=depth_pred.device)
        grad_depth = zero if grad_depth is None else grad_depth.to(torch.float32).contiguous()
        grad_normals = zero if grad_normals is None else grad_normals.to(torch.float32).contiguous()
        grad_pred = torch.empty_like(depth_pred)
        with torch.cuda.device(depth_pred.device):
            _lib.check(_lib.lib().polcue_supervised_losses_bwd_f32(_ptr(depth_gt), _ptr(depth_pred), _ptr(camera_matrix), ctx.range[0],
                                                                   ctx.range[1], b, h, w, _ptr(sums), _ptr(grad_normals),
                                                                   _ptr(grad_depth), _ptr(grad_pred), _stream(depth_pred)),
                       "polcue_supervised_losses_bwd_f32")
        return None, grad_pred, None, None, None


def supervised_losses(depth_gt, depth_pred, camera_matrix, min_depth, max_depth):
    """The supervised block of Trainer.compute_losses (manydepth/trainer.py:1240-1251) for one scale in one forward and
    one backward kernel:
        mask = (gt >= min_depth).float() * (gt <= max_depth).float()
        supervised_depth_loss   = (|gt - pred| * mask).sum() / mask.sum()
        supervised_normals_loss = compute_supervised_normals_losses(gt, pred, K, mask)
    Returns (supervised_depth_loss, supervised_normals_loss), zero-dim float32, differentiable w.r.t. depth_pred."""
    depth_gt = _need_cuda(depth_gt, "depth_gt").detach().float().contiguous()
    depth_pred = _need_cuda(depth_pred, "depth_pred").float().contiguous()
    camera_matrix = _need_cuda(camera_matrix, "camera_matrix").detach().float().contiguous()
    if depth_pred.dim() != 4 or depth_pred.shape[1] != 1 or depth_gt.shape != depth_pred.shape:
        raise ValueError("depth_gt and depth_pred must share one B x 1 x H x W shape")
    if camera_matrix.shape != (depth_pred.shape[0], 3, 3):
        raise ValueError("camera_matrix must be B x 3 x 3")
    return _SupervisedLosses.apply(depth_gt, depth_pred, camera_matrix, float(min_depth), float(max_depth))


def normals_loss(depth_gt, depth_pred, camera_matrix, mask):
    """Trainer.compute_supervised_normals_losses (manydepth/trainer.py:1298-1309): zero-dim float32 loss, differentiable
    w.r.t. depth_pred.  depth_gt / depth_pred / mask: B x 1 x H x W, camera_matrix: B x 3 x 3 (CUDA)."""
    depth_gt = _need_cuda(depth_gt, "depth_gt").detach().float().contiguous()
    depth_pred = _need_cuda(depth_pred, "depth_pred").float().contiguous()
    camera_matrix = _need_cuda(camera_matrix, "camera_matrix").detach().float().contiguous()
    mask = _need_cuda(mask, "mask").detach().float().contiguous()
    if depth_pred.dim() != 4 or depth_pred.shape[1] != 1 or depth_gt.shape != depth_pred.shape or mask.shape != depth_pred.shape:
        raise ValueError("depth_gt, depth_pred and mask must share one B x 1 x H x W shape")
    if camera_matrix.shape != (depth_pred.shape[0], 3, 3):
        raise ValueError("camera_matrix must be B x 3 x 3")
    return _NormalsLoss.apply(depth_gt, depth_pred, camera_matrix, mask)
