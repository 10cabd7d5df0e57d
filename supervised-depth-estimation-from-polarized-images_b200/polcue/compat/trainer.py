"""Mirror of the hot-path member of manydepth/trainer.py: Trainer.compute_supervised_normals_losses (:1298-1309)."""
from .. import ops


def compute_supervised_normals_losses(depth_gt, depth_pred, intrinsics, mask):
    """
    Compute the normals loss based on pixel-wise cosine similarity.
    depth_gt, depth_pred: B x 1 x H x W; intrinsics: B x 4 x 4 (or B x 3 x 3); mask: B x 1 x H x W.
    One fused forward kernel and one fused backward kernel (gradient w.r.t. depth_pred).
    """
    camera_matrix = intrinsics[:, :3, :3]
    return ops.normals_loss(depth_gt, depth_pred, camera_matrix, mask)
