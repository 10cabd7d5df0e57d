"""Generate tests/golden/resize_golden.npz: outputs of the reference's resize configuration, executed here.

The reference resizes the four gray polarizer images with ``transforms.Resize((height, width),
interpolation=Image.ANTIALIAS)`` (manydepth/datasets/indoor_dataset.py:77,115,335-349) after ``img.convert('L')``
(:31-32) and an optional ``transpose(FLIP_LEFT_RIGHT)`` (hammer_dataset.py:72-73).  The arithmetic is Pillow's
(pinned 6.2.1; ``Image.ANTIALIAS`` is the old name of ``Image.LANCZOS``, removed in Pillow 10).  This script runs
exactly that call chain with the installed torchvision + Pillow on seeded images and stores inputs and outputs;
tests/test_oracle_golden.py checks the oracle restatement against them, the GPU tests check the kernels.

Run once, here:   python tests/golden/make_resize_golden.py
"""
import hashlib
import os
import sys

import numpy as np
import PIL
from PIL import Image
import torchvision
from torchvision import transforms

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "supervised-depth-estimation-from-polarized-images_b200"))
from polcue import synth  # noqa: E402

LANCZOS = getattr(Image, "ANTIALIAS", Image.LANCZOS)


def reference_resize(img_u8, out_hw, flip):
    pil = Image.fromarray(img_u8).convert("L")
    if flip:
        pil = pil.transpose(Image.FLIP_LEFT_RIGHT)
    try:
        resize = transforms.Resize(tuple(out_hw), interpolation=LANCZOS)
    except (TypeError, ValueError):
        resize = transforms.Resize(tuple(out_hw), interpolation=transforms.InterpolationMode.LANCZOS)
    out = np.asarray(resize(pil))
    assert np.array_equal(out, np.asarray(pil.resize((out_hw[1], out_hw[0]), Image.LANCZOS)))
    return out


def main():
    out = {"versions": np.array([PIL.__version__, torchvision.__version__])}
    rng = np.random.default_rng(77)
    cases = [((104, 136), (40, 60)), ((64, 96), (20, 30)), ((50, 70), (50, 30)), ((33, 47), (66, 94)), ((100, 100), (37, 100)),
             ((7, 9), (3, 2)), ((5, 5), (11, 13)), ((61, 83), (23, 31)), ((208, 272), (80, 120)), ((96, 16), (8, 16))]
    out["case_shapes"] = np.array([list(a) + list(b) for a, b in cases])
    for i, (in_hw, out_hw) in enumerate(cases):
        if i % 2:
            img = rng.integers(0, 256, in_hw, dtype=np.uint8)
        else:
            img = synth.gen_p_planes(900 + i, *in_hw)[i % 4]
        out[f"in_{i}"] = img
        out[f"out_{i}"] = reference_resize(img, out_hw, False)
        out[f"outflip_{i}"] = reference_resize(img, out_hw, True)
    # HAMMER geometry (832 x 1088 quadrants -> 320 x 480, xolp_and_normals.py:106-108): inputs regenerated from the seed
    planes = synth.gen_p_planes(4242, 832, 1088)
    digests = []
    for k in range(4):
        small = reference_resize(planes[k], (320, 480), bool(k & 1))
        out[f"hammer_sample_{k}"] = small[::8, ::8].copy()
        digests.append(hashlib.sha256(small.tobytes()).hexdigest())
    out["hammer_sha256"] = np.array(digests)
    out["hammer_in_sha256"] = np.array([hashlib.sha256(p.tobytes()).hexdigest() for p in planes])
    path = os.path.join(HERE, "resize_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
