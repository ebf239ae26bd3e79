"""transformation_matrix(quaternion, translation) -> [N,4,4] | [4,4].

API of morefusion/functions/geometry/transformation_matrix.py:5-18."""

from .compose_transform import compose_transform
from .quaternion_matrix import quaternion_matrix
from . import _util


def transformation_matrix(quaternion, translation):
    quaternion = _util.as_f32(quaternion)
    translation = _util.as_f32(translation, quaternion.device)
    if quaternion.dim() == 2:
        batch_size = quaternion.shape[0]
        assert tuple(quaternion.shape) == (batch_size, 4)
        assert tuple(translation.shape) == (batch_size, 3)
        T = quaternion_matrix(quaternion)
        T = compose_transform(T[:, :3, :3], translation)
    else:
        assert quaternion.dim() == 1
        assert tuple(quaternion.shape) == (4,)
        assert tuple(translation.shape) == (3,)
        T = quaternion_matrix(quaternion[None])[0]
        T = compose_transform(T[None, :3, :3], translation[None])[0]
    return T
