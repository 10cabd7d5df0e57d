"""Drop-in mirrors of the reference's hot-path functions.

Each module keeps the name, argument order/meaning, return order and error behaviour of the reference
module it replaces, so a call site changes only its import line (INTEGRATION.md lists them):

  reference module                               mirror
  polarisation/pol_split_and_save.py             polcue.compat.pol_split_and_save
  polarisation/xolp.py                           polcue.compat.xolp
  polarisation/xolp_and_normals.py               polcue.compat.xolp_and_normals
  ppp_code/physical_normals_channels.py          polcue.compat.physical_normals_channels
  manydepth/normals_vec.py                       polcue.compat.normals_vec
  manydepth/networks/pre_encoders.py (get_normals)  polcue.compat.normals_vec.get_normals
  manydepth/layers.py (compute_depth_errors*)    polcue.compat.layers
  kornia.geometry.depth (depth_to_normals)       polcue.compat.trainer.depth_to_normals
  manydepth/datasets/indoor_dataset.py (resize_pol, get_xolp)  polcue.compat.indoor_dataset
  manydepth/trainer.py, manydepth/evaluation.py (compute_depth_losses_from_list)  polcue.compat.trainer

numpy-signature functions take numpy arrays and return numpy float64 like the reference (the data makes one
round trip through the GPU); torch-signature functions take and return CUDA tensors.
"""
import torch


def default_device():
    if not torch.cuda.is_available():
        raise RuntimeError("polcue needs a CUDA device: there is no CPU fallback for the polarization hot path")
    return torch.device("cuda", torch.cuda.current_device())


def to_device(array, dtype=None):
    """numpy array -> CUDA tensor (contiguous)."""
    import numpy as np
    t = torch.from_numpy(np.ascontiguousarray(array))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(default_device(), non_blocking=False)


_SHADOWED = {   # reference module name -> mirror module (the two small top-level packages of the reference)
    "polarisation.xolp": "xolp",
    "polarisation.pol_split_and_save": "pol_split_and_save",
    "polarisation.xolp_and_normals": "xolp_and_normals",
    "ppp_code.physical_normals_channels": "physical_normals_channels",
}


def shadow():
    """Serve `polarisation.{xolp, pol_split_and_save, xolp_and_normals}` and `ppp_code.physical_normals_channels` under the
    reference's own module names: after `polcue.compat.shadow()`, `from polarisation.xolp import Iun_and_xolp` resolves to
    the CUDA implementation without the reference checkout on `sys.path`.  Call it BEFORE anything imports those modules;
    if the reference's own modules are already in `sys.modules` this raises (use `install()` to patch them in place).
    Returns the list of module names now served."""
    import importlib
    import sys
    import types

    for name in _SHADOWED:
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__name__", "").startswith(__name__ + "."):
            raise RuntimeError(f"{name} is already imported from {getattr(mod, '__file__', '?')}: call polcue.compat.install() instead")
    for name, mirror in _SHADOWED.items():
        pkg_name, _, leaf = name.partition(".")
        pkg = sys.modules.get(pkg_name)
        if pkg is None:
            pkg = types.ModuleType(pkg_name, f"polcue shadow of the reference's `{pkg_name}` package (hot-path modules only)")
            pkg.__path__ = []          # a package with no files: only the modules registered here can be imported from it
            sys.modules[pkg_name] = pkg
        mod = importlib.import_module("." + mirror, __name__)
        sys.modules[name] = mod
        setattr(pkg, leaf, mod)
    return sorted(_SHADOWED)


def install(verbose=False, patch_loader=False):
    """Replace, in place, the hot-path functions of the reference modules that are ALREADY imported (`sys.modules`).

    `patch_loader` (default False): also replace `manydepth.datasets.indoor_dataset.Iun_and_xolp`.  That function runs
    inside `__getitem__`, i.e. in the DataLoader's FORKED worker processes (num_workers = 12 by default,
    options.py:299-302), where CUDA cannot be (re-)initialised; the CUDA-backed mirror is only valid there with
    `num_workers=0` or a `spawn` start method, so it is opt-in.  The recommended integration keeps the loader's uint8
    planes and runs `polcue.ops.fused_planes` / `loader_front_end` on the collated batch in the main process
    (INTEGRATION.md).  The mirror itself refuses to run in a forked worker with a clear error.

    For the `manydepth` package, of which only a few functions are on the path:
      manydepth.normals_vec.{rho_diffuse, rho_spec, calc_normals}
      manydepth.networks.pre_encoders.{rho_diffuse, rho_spec, calc_normals} and ShallowNormalsEncoder.get_normals
      manydepth.layers.{compute_depth_errors, compute_depth_errors_numpy}
      manydepth.trainer.{compute_depth_errors, compute_depth_errors_numpy, Iun_and_xolp} and
          Trainer.compute_supervised_normals_losses (the GT + predicted stencils, cosine and masked mean, with backward),
          Trainer.compute_depth_losses_from_list / manydepth.evaluation.Evaluation.compute_depth_losses_from_list
      manydepth.datasets.indoor_dataset.Iun_and_xolp, polarisation.* and ppp_code.physical_normals_channels.* if imported.
    Returns the list of "module.attribute" names that were replaced.
    """
    import sys

    from . import layers, normals_vec, physical_normals_channels, pol_split_and_save, trainer, xolp, xolp_and_normals

    done = []

    def patch(mod_name, attr, value, owner=None):
        mod = sys.modules.get(mod_name)
        target = getattr(mod, owner, None) if (mod is not None and owner) else mod
        if target is not None and hasattr(target, attr):
            setattr(target, attr, value)
            done.append(f"{mod_name}.{owner + '.' if owner else ''}{attr}")

    for name in ("rho_diffuse", "rho_spec", "calc_normals"):
        patch("manydepth.normals_vec", name, getattr(normals_vec, name))
        patch("manydepth.networks.pre_encoders", name, getattr(normals_vec, name))
    patch("manydepth.networks.pre_encoders", "get_normals", staticmethod(normals_vec.get_normals), owner="ShallowNormalsEncoder")
    for mod_name in ("manydepth.layers", "manydepth.trainer", "manydepth.evaluation"):
        patch(mod_name, "compute_depth_errors", layers.compute_depth_errors)
        patch(mod_name, "compute_depth_errors_numpy", layers.compute_depth_errors_numpy)
    for mod_name in ("manydepth.trainer", "polarisation.xolp", "polarisation.xolp_and_normals") + (
            ("manydepth.datasets.indoor_dataset",) if patch_loader else ()):
        patch(mod_name, "Iun_and_xolp", xolp.Iun_and_xolp)
    patch("manydepth.trainer", "compute_supervised_normals_losses",
          lambda self, depth_gt, depth_pred, intrinsics, mask: trainer.compute_supervised_normals_losses(depth_gt, depth_pred, intrinsics, mask),
          owner="Trainer")
    patch("manydepth.trainer", "compute_depth_losses_from_list", trainer.compute_depth_losses_from_list, owner="Trainer")
    patch("manydepth.evaluation", "compute_depth_losses_from_list", trainer.compute_depth_losses_from_list, owner="Evaluation")
    for mod_name in ("polarisation.pol_split_and_save", "polarisation.xolp_and_normals"):
        patch(mod_name, "split_pol", pol_split_and_save.split_pol)
    for name in ("rho_diffuse", "rho_spec", "calc_normals"):
        patch("polarisation.xolp_and_normals", name, getattr(xolp_and_normals, name))
    for name in ("PolarisationImage_channel", "rho_diffuse_channel", "rho_spec_channel", "calc_normals_channel"):
        patch("ppp_code.physical_normals_channels", name, getattr(physical_normals_channels, name))
    if verbose:
        print("polcue.compat.install: replaced", ", ".join(done) or "nothing (import the reference modules first)")
    return done
