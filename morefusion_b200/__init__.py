"""morefusion_b200: B200-native (sm_100a) implementation of MoreFusion's volumetric-pose
hot path behind the reference's Python operator API (morefusion.functions /
morefusion.contrib), on torch CUDA tensors.  See DESIGN.md."""

__version__ = "0.1.0"


class InvalidType(TypeError):
    """Mirror of chainer.utils.type_check.InvalidType raised by check_type_forward."""


class config:
    # the reference raises ValueError("points include nan") after a device->host sync
    # (average_voxelization_3d.py:47-48); set False to skip the sync on hot paths.
    check_nan = True


from . import functions  # noqa: E402,F401
from . import contrib  # noqa: E402,F401
