"""Mirror of the hot-path members of manydepth/trainer.py: `depth_to_normals` as imported at :37 (kornia.geometry.depth),
Trainer.compute_supervised_normals_losses (:1298-1309), the supervised block of compute_losses (:1240-1251) and
compute_depth_losses_from_list (:1356-1435, also evaluation.py:215-288)."""
from .. import ops


def depth_to_normals(depth, camera_matrix, normalize_points=False):
    """kornia.geometry.depth.depth_to_normals (kornia 0.5.11), forward only: the GT branch (:1305, :1477).  The predicted
    branch of the loss has its own backward inside compute_supervised_normals_losses; a tensor that requires grad raises."""
    if normalize_points:
        raise NotImplementedError("normalize_points=True is never used by the reference (trainer.py:1305-1306)")
    return ops.depth_to_normals(depth, camera_matrix)


def compute_supervised_normals_losses(depth_gt, depth_pred, intrinsics, mask):
    """
    Compute the normals loss based on pixel-wise cosine similarity.
    depth_gt, depth_pred: B x 1 x H x W; intrinsics: B x 4 x 4 (or B x 3 x 3); mask: B x 1 x H x W.
    One fused forward kernel and one fused backward kernel (gradient w.r.t. depth_pred).
    """
    camera_matrix = intrinsics[:, :3, :3]
    return ops.normals_loss(depth_gt, depth_pred, camera_matrix, mask)


def compute_supervised_losses(depth_gt, depth_pred, intrinsics, min_depth, max_depth):
    """The supervised block of Trainer.compute_losses for one scale (trainer.py:1240-1251): builds the range mask, the
    masked L1 depth loss and the normals loss in one fused forward (and one fused backward) kernel.
    Returns (supervised_depth_loss, supervised_normals_loss)."""
    return ops.supervised_losses(depth_gt, depth_pred, intrinsics[:, :3, :3], min_depth, max_depth)


# instance-id levels of the material groups (trainer.py:1389-1408, evaluation.py:237-262); "objects" is a RANGE of ids
OBJECT_IDS = {"box": 20, "bottle": 40, "can": 60, "cup": 80, "remote": 100, "teapot": 120, "cutlery": 140, "glass": 160,
              "table": 180, "wall": 200, "objects": (20, 160)}
DEPTH_METRIC_NAMES = ("de/abs_rel", "de/sq_rel", "de/rms", "de/log_rms", "da/a1", "da/a2", "da/a3")


def compute_depth_losses_from_list(self, gts, preds, losses, masks, object="all"):
    """Trainer.compute_depth_losses_from_list (trainer.py:1356-1435) / Evaluation.compute_depth_losses_from_list
    (evaluation.py:215-288): per-image masked depth metrics over lists of batches, mean over all images, written into
    `losses[metric]` as 0-d numpy arrays and printed in the reference's table format.

    `self` supplies the options exactly where the reference reads them: `self.opt.{min_depth, max_depth, height, width,
    depth_supervision, train_stereo_only}` (Trainer) or `self.{min_depth, max_depth, height, width}` (Evaluation, which
    never median-scales), and `self.depth_metric_names`.  The lists may hold CPU tensors (the reference stages them with
    `.cpu()`) or CUDA tensors; every image of a batch is evaluated (the reference loops over `batch_size`).
    An image whose mask is empty contributes NaN, as `compute_depth_errors_numpy` does on empty arrays.
    """
    import numpy as np
    import torch
    import torch.nn.functional as F

    from . import default_device

    opt = getattr(self, "opt", self)
    min_depth, max_depth = float(opt.min_depth), float(opt.max_depth)
    height, width = int(opt.height), int(opt.width)
    median_scaling = hasattr(self, "opt") and not getattr(opt, "depth_supervision", True) and not getattr(opt, "train_stereo_only", False)
    inst_id = None if object == "all" else OBJECT_IDS[object]
    dev = default_device()
    rows = []
    for k in range(len(preds)):
        pred = preds[k].detach().to(dev, torch.float32)
        if tuple(pred.shape[-2:]) != (height, width):      # identity for scale-0 predictions; torch is plumbing here
            pred = F.interpolate(pred, [height, width], mode="bilinear", align_corners=False)
        gt = gts[k].detach().to(dev, torch.float32)[:, 0]
        inst = None
        if inst_id is not None:
            inst = masks[k].detach().to(dev)[:, 0]
            if inst.dtype != torch.uint8:                 # the loader's mask is (to_tensor(..) * 255).int(): ids 0..255
                inst = inst.clamp(0, 255).to(torch.uint8)
        _, metrics = ops.depth_errors_per_image(gt, pred[:, 0], min_depth, max_depth, inst, inst_id,
                                                median_scaling=median_scaling, clamp_first=True)
        rows.append(metrics)
    mean_errors = torch.cat(rows).double().mean(0).cpu().numpy() if rows else np.full(7, np.nan)
    print("\n  " + ("{:>8} | " * 7).format("abs_rel", "sq_rel", "rmse", "rmse_log", "a1", "a2", "a3"))
    print(("&{: 8.5f}  " * 7).format(*mean_errors.tolist()) + "\\\\")
    names = getattr(self, "depth_metric_names", DEPTH_METRIC_NAMES)
    for i, metric in enumerate(names):
        losses[metric] = np.array(mean_errors[i])
    return mean_errors
