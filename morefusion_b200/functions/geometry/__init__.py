# same re-exports as morefusion/functions/geometry/__init__.py:3-26
from .average_voxelization_3d import average_voxelization_3d  # noqa: F401
from .compose_transform import compose_transform  # noqa: F401
from .max_voxelization_3d import max_voxelization_3d  # noqa: F401
from .occupancy_grid_3d import occupancy_grid_3d  # noqa: F401
from .interpolate_voxel_grid import interpolate_voxel_grid  # noqa: F401
from .quaternion_matrix import quaternion_matrix  # noqa: F401
from .transform_points import transform_points  # noqa: F401
from .transformation_matrix import transformation_matrix  # noqa: F401
from .translation_matrix import translation_matrix  # noqa: F401
from .truncated_distance_function import truncated_distance_function  # noqa: F401
from .truncated_distance_function import pseudo_occupancy_voxelization  # noqa: F401
