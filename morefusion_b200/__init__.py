"""morefusion_b200: B200-native (sm_100a) implementation of MoreFusion's volumetric-pose
hot path behind the reference's Python operator API (morefusion.functions /
morefusion.contrib), on torch CUDA tensors.  See DESIGN.md."""

__version__ = "0.1.0"


class InvalidType(TypeError):
    """Mirror of chainer.utils.type_check.InvalidType raised by check_type_forward."""


import contextlib


class config:
    # the reference raises ValueError("points include nan") after a device->host sync
    # (average_voxelization_3d.py:47-48); set False to skip the sync on hot paths.
    check_nan = True

    @staticmethod
    @contextlib.contextmanager
    def no_nan_check():
        """Temporarily skip the NaN check (and its host sync), e.g. under CUDA-graph capture."""
        old = config.check_nan
        config.check_nan = False
        try:
            yield
        finally:
            config.check_nan = old


from . import functions  # noqa: E402,F401
from . import contrib  # noqa: E402,F401
