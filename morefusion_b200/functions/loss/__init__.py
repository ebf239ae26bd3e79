from .average_distance import average_distance  # noqa: F401
