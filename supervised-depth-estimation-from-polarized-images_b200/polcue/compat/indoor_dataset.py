"""Mirror of the polarization branch of manydepth/datasets/indoor_dataset.py:
   self.resize_pol = transforms.Resize((height, width), interpolation=Image.ANTIALIAS)   (:77, :115)
   inputs[("polXX_gray", i, 0)] = self.resize_pol(polXX_gray)                           (:335-349)
   IndoorDataset.get_xolp(inputs, i)                                                    (:430-442)
and the batched replacement of all three (`polarization_inputs`)."""
import numpy as np
import torch

from .. import ops
from . import to_device


class ResizePol:
    """`transforms.Resize((height, width), interpolation=Image.ANTIALIAS)` for the 8-bit gray polarizer images:
    PIL 'L' image (or H x W uint8 array) in, the same kind of object out, bit-exact with Pillow."""

    def __init__(self, size):
        self.size = (int(size[0]), int(size[1]))     # (height, width), as torchvision

    def __call__(self, img):
        is_pil = hasattr(img, "mode") and hasattr(img, "size")
        if is_pil and img.mode != "L":
            raise ValueError(f"ResizePol handles 8-bit gray ('L') images, got mode {img.mode!r}")
        arr = np.asarray(img)
        if arr.ndim != 2 or arr.dtype != np.uint8:
            raise ValueError("ResizePol needs an H x W uint8 image")
        out = ops.lanczos_resize(to_device(arr), self.size).cpu().numpy()
        if is_pil:
            from PIL import Image
            return Image.fromarray(out, "L")
        return out


def get_xolp(self, inputs, i):
    """Calculate and concatenate DOLP and AOLP (indoor_dataset.py:430-442): reads the four resized gray images
    `inputs[("polXX_gray", i, 0)]`, writes `inputs[("xolp", i, 0)]`, a 2 x H x W float64 CPU tensor like
    `to_tensor(np.stack((dolp, aolp), axis=2))`.  `self` is unused (bind it as a method or pass None)."""
    im00 = np.asarray(inputs[("pol00_gray", i, 0)])  # 0 deg
    im10 = np.asarray(inputs[("pol10_gray", i, 0)])  # 90 deg
    im01 = np.asarray(inputs[("pol01_gray", i, 0)])  # 45 deg
    im11 = np.asarray(inputs[("pol11_gray", i, 0)])  # 135 deg
    _, xolp = ops.xolp_from_planes(*(to_device(im) for im in (im00, im01, im10, im11)), want_iun=False)
    inputs[("xolp", i, 0)] = xolp[0].to(torch.float64).cpu()


def polarization_inputs(pol00, pol10, pol01, pol11, size, do_flip=None, n=1.5, normalize=True):
    """The batched, main-process replacement of the loader's polarization branch plus the encoders' front end.

    pol00, pol10, pol01, pol11: B x H x W uint8 CUDA tensors in the LOADER's naming and order (0, 90, 45, 135 deg --
    indoor_dataset.py:435-438), at full resolution, not yet flipped; `do_flip`: per-sample flags (hammer_dataset.py:72).
    Returns dict(planes [B,4,h,w] u8 in angle order, xolp [B,2,h,w], normals [B,9,h,w], xolp_norm if `normalize`)."""
    return ops.loader_front_end(pol00, pol01, pol10, pol11, size, n=n, flip=do_flip,
                                normalize_xolp=ops.XOLP_MEAN_STD if normalize else None)
