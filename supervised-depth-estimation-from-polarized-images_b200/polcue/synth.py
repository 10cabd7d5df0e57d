"""Seeded synthetic inputs shaped like the reference's data.

Gen-P  physically plausible polarizer mosaics: low-frequency radiance L in [30, 170], DoLP rho* in
       [0.02, 0.45] and AoLP phi*, I_k = L (1 + rho* cos(2 (a_k - phi*))) + N(0, 1.5^2), 8-bit,
       quadrants TL/TR/BL/BR = 0/45/90/135 degrees (manydepth/datasets/indoor_dataset.py:435-439).
Gen-U  uniform random bytes (stress: 69 % of pixels leave the diffuse table, 10 % have rho > 1).
Depth  smooth metric depth with 10 % invalid (0) pixels, a noisy clamped prediction, an
       instance-id map with the material levels of manydepth/trainer.py:1389-1411 and the
       HAMMER intrinsics of manydepth/evaluation.py:50-52 rescaled to the image size.

Every frame is seeded by its own index, so any sharding of a sequence sees identical frames.
Pure numpy; the torch variant below evaluates the same formulas on a device for large batches.
"""
from __future__ import annotations

import numpy as np

FRAME_H, FRAME_W = 2048, 2448          # BASELINE.json frame (mosaic) size
TRAIN_H, TRAIN_W = 320, 480            # train_supervised_GT.sh:8
HAMMER_K = np.array([[706.75531005859375, 0.0, 545.6326819328060083],
                     [0.0, 707.5133056640625, 389.9299663507044897],
                     [0.0, 0.0, 1.0]])  # manydepth/evaluation.py:50-52 (1088 x 832 image)
MATERIAL_LEVELS = (20, 40, 60, 80, 100, 120, 140, 160, 180, 200)
_ANGLES = np.array([0.0, 45.0, 90.0, 135.0]) * np.pi / 180.0


def _fields(rng, hs, ws):
    """Three low-frequency fields in [0, 1] (sums of 3 sinusoids, periods 50-400 px)."""
    v, u = np.mgrid[0:hs, 0:ws].astype(np.float64)
    out = []
    for _ in range(3):
        acc = np.zeros((hs, ws))
        for _ in range(3):
            period = rng.uniform(50.0, 400.0)
            ang = rng.uniform(0.0, 2 * np.pi)
            phase = rng.uniform(0.0, 2 * np.pi)
            acc += np.sin(2 * np.pi * (u * np.cos(ang) + v * np.sin(ang)) / period + phase)
        out.append(0.5 + acc / 6.0)
    return out


def gen_p_planes(frame, hs, ws):
    """Four uint8 planes (I0, I45, I90, I135), each hs x ws, for frame index `frame`."""
    rng = np.random.default_rng(1234 + int(frame))
    f_l, f_r, f_p = _fields(rng, hs, ws)
    lum = 30.0 + 140.0 * f_l
    rho = 0.02 + 0.43 * f_r
    phi = np.pi * f_p
    planes = []
    for a in _ANGLES:
        val = lum * (1.0 + rho * np.cos(2.0 * (a - phi))) + rng.normal(0.0, 1.5, (hs, ws))
        planes.append(np.clip(np.rint(val), 0, 255).astype(np.uint8))
    return planes


def tile_mosaic(planes):
    """(I0, I45, I90, I135) -> 2hs x 2ws quadrant-tiled mosaic (TL, TR, BL, BR)."""
    i0, i45, i90, i135 = planes
    return np.block([[i0, i45], [i90, i135]])


def gen_p_mosaic(frame, h=FRAME_H, w=FRAME_W):
    return tile_mosaic(gen_p_planes(frame, h // 2, w // 2))


def gen_u_mosaic(frame, h=FRAME_H, w=FRAME_W):
    return np.random.default_rng(1234 + int(frame)).integers(0, 256, (h, w), dtype=np.uint8)


def gen_batch(kind, first_frame, count, h=FRAME_H, w=FRAME_W):
    fn = {"P": gen_p_mosaic, "U": gen_u_mosaic}[kind]
    return np.stack([fn(first_frame + i, h, w) for i in range(count)])


def scaled_intrinsics(h, w):
    k = HAMMER_K.copy()
    k[0] *= w / 1088.0
    k[1] *= h / 832.0
    return k


def gen_depth_sample(index, h=TRAIN_H, w=TRAIN_W):
    """(gt, pred, inst, K) for eval image `index`: float32 h x w x2, uint8 h x w, float32 3x3."""
    rng = np.random.default_rng(4321 + int(index))
    v, u = np.mgrid[0:h, 0:w].astype(np.float64)
    gx, gy = rng.uniform(-4e-4, 4e-4, 2)
    z = 0.6 + 0.3 * np.sin(u / 60.0 + rng.uniform(0, 6.28)) * np.cos(v / 45.0) + gx * u + gy * v
    z += 0.25 * (u > w * rng.uniform(0.3, 0.7))                  # a depth step (object edge)
    invalid = rng.random((h, w)) < 0.10
    gt = np.where(invalid, 0.0, z).astype(np.float32)
    pred = np.clip(z * (1.0 + 0.1 * rng.normal(size=(h, w))), 0.1, 2.0).astype(np.float32)
    blocks = rng.integers(0, 11, (h // 40 + 1, w // 40 + 1))
    inst = (np.kron(blocks, np.ones((40, 40), dtype=np.int64))[:h, :w] * 20).astype(np.uint8)
    return gt, pred, inst, scaled_intrinsics(h, w).astype(np.float32)


def gen_depth_batch(first, count, h=TRAIN_H, w=TRAIN_W):
    items = [gen_depth_sample(first + i, h, w) for i in range(count)]
    return tuple(np.stack([it[k] for it in items]) for k in range(4))


def add_hole_regions(gt, seed=0):
    """HAMMER-like missing regions on top of the 10 % random invalid pixels of `gen_depth_sample`: per image a few
    elliptic blobs (depth sensor drop-outs on glossy / transparent objects), a strip along one image border and one
    isolated valid pixel inside a blob (its eight neighbours are all invalid).  gt: B x H x W float32; returns a copy."""
    gt = np.array(gt, dtype=np.float32, copy=True)
    b, h, w = gt.shape
    v, u = np.mgrid[0:h, 0:w]
    for i in range(b):
        rng = np.random.default_rng(977 + int(seed) + i)
        for _ in range(4):
            cy, cx = rng.uniform(0, h), rng.uniform(0, w)
            ry, rx = rng.uniform(2, max(3, h / 6)), rng.uniform(2, max(3, w / 6))
            gt[i][((v - cy) / ry) ** 2 + ((u - cx) / rx) ** 2 < 1.0] = 0.0
        side = int(rng.integers(0, 4))
        width = int(rng.integers(1, max(2, min(h, w) // 10 + 1)))
        if side == 0: gt[i, :width] = 0.0
        elif side == 1: gt[i, -width:] = 0.0
        elif side == 2: gt[i, :, :width] = 0.0
        else: gt[i, :, -width:] = 0.0
        if h >= 5 and w >= 5:
            y, x = int(rng.integers(2, h - 2)), int(rng.integers(2, w - 2))
            gt[i, y - 2:y + 3, x - 2:x + 3] = 0.0
            gt[i, y, x] = 0.8                                   # lonely valid pixel
    return gt


# --------------------------------------------------------------------------------------
# torch variant of Gen-P for large resident batches (bench only; same formulas, torch RNG)
# --------------------------------------------------------------------------------------
def gen_p_batch_torch(first_frame, count, h=FRAME_H, w=FRAME_W, device="cuda"):
    import torch

    hs, ws = h // 2, w // 2
    out = torch.empty((count, h, w), dtype=torch.uint8, device=device)
    v = torch.arange(hs, device=device, dtype=torch.float32)[:, None]
    u = torch.arange(ws, device=device, dtype=torch.float32)[None, :]
    for i in range(count):
        rng = np.random.default_rng(1234 + first_frame + i)
        fields = []
        for _ in range(3):
            acc = torch.zeros((hs, ws), device=device)
            for _ in range(3):
                period = rng.uniform(50.0, 400.0)
                ang = rng.uniform(0.0, 2 * np.pi)
                phase = rng.uniform(0.0, 2 * np.pi)
                acc += torch.sin((2 * np.pi / period) * (u * float(np.cos(ang)) + v * float(np.sin(ang))) + phase)
            fields.append(0.5 + acc / 6.0)
        lum = 30.0 + 140.0 * fields[0]
        rho = 0.02 + 0.43 * fields[1]
        phi = np.pi * fields[2]
        gen = torch.Generator(device=device)
        gen.manual_seed(1234 + first_frame + i)
        for k, (r0, c0) in enumerate(((0, 0), (0, ws), (hs, 0), (hs, ws))):
            val = lum * (1.0 + rho * torch.cos(2.0 * (float(_ANGLES[k]) - phi)))
            val = val + 1.5 * torch.randn((hs, ws), device=device, generator=gen)
            out[i, r0:r0 + hs, c0:c0 + ws] = val.round().clamp(0, 255).to(torch.uint8)
    return out
