# hot-path members of morefusion/contrib/__init__.py:3-11
from .iterative_collision_check_link import IterativeCollisionCheckLink  # noqa: F401
from . import singleview_3d  # noqa: F401
from .occupancy_registration import OccupancyRegistration  # noqa: F401  
from .multi_instance_octree_mapping import MultiInstanceOctreeMapping  # noqa: F401  
