"""Pose metrics (mirror of morefusion/metrics/__init__.py:3-7)."""

from .auc_for_errors import auc_for_errors  # noqa: F401
from .average_distance import average_distance  # noqa: F401
from .ycb_video_add_auc import ycb_video_add_auc  # noqa: F401
