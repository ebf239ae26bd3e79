# same re-exports as morefusion/functions/__init__.py:3-15
from .geometry import average_voxelization_3d  # noqa: F401
from .geometry import compose_transform  # noqa: F401
from .geometry import interpolate_voxel_grid  # noqa: F401
from .geometry import max_voxelization_3d  # noqa: F401
from .geometry import occupancy_grid_3d  # noqa: F401
from .geometry import pseudo_occupancy_voxelization  # noqa: F401
from .geometry import quaternion_matrix  # noqa: F401
from .geometry import transform_points  # noqa: F401
from .geometry import transformation_matrix  # noqa: F401
from .geometry import translation_matrix  # noqa: F401
from .geometry import truncated_distance_function  # noqa: F401

from .loss import average_distance  # noqa: F401
