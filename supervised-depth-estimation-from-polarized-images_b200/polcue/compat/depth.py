"""Mirror of kornia.geometry.depth.depth_to_normals as imported at manydepth/trainer.py:37 (forward / GT branch)."""
from .. import ops


def depth_to_normals(depth, camera_matrix, normalize_points=False):
    if normalize_points:
        raise NotImplementedError("normalize_points=True is never used by the reference (trainer.py:1305-1306)")
    return ops.depth_to_normals(depth, camera_matrix)
