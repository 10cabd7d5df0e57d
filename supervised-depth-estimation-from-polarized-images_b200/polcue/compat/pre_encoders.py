"""Mirror of the hot-path member of manydepth/networks/pre_encoders.py: ShallowNormalsEncoder.get_normals (:99-113)."""
from .. import ops


def get_normals(x, n=1.5):
    """input: XOLP B x 2 x H x W (CUDA); output: 3 concatenated normals B x 9 x H x W (float32)."""
    return ops.get_normals(x, n)


class GetNormalsMixin:
    """`class ShallowNormalsEncoder(GetNormalsMixin, ShallowEncoder)` keeps `self.get_normals(x)` working."""

    get_normals = staticmethod(get_normals)
