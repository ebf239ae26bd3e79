"""transformation_matrix(quaternion, translation): rigid transform(s) T = translate(t) @ rotate(q).

Same call contract as the reference operator
(morefusion/functions/geometry/transformation_matrix.py:5-18): q [N,4] with t [N,3] gives
[N,4,4]; a single q [4] with t [3] gives [4,4]; any other shape combination trips an assert.
Both cases run through one batched call of the CUDA quaternion / compose kernels."""

from . import _util
from .compose_transform import compose_transform
from .quaternion_matrix import quaternion_matrix


def transformation_matrix(quaternion, translation):
    q = _util.as_f32(quaternion)
    t = _util.as_f32(translation, q.device)
    single = q.dim() == 1
    assert q.dim() in (1, 2), "quaternion must be [4] or [N,4]"
    if single:
        assert t.dim() == 1, "a single quaternion takes a single translation [3]"
        q, t = q[None], t[None]
    n = q.shape[0]
    assert tuple(q.shape) == (n, 4), "quaternion rows must have 4 components (w, x, y, z)"
    assert tuple(t.shape) == (n, 3), "one translation [3] per quaternion"
    rotation = quaternion_matrix(q)[:, :3, :3]
    T = compose_transform(rotation, t)
    return T[0] if single else T
