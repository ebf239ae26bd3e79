"""2-D backbones feeding the volumetric pose head (mirror of morefusion/models/__init__.py).

SURVEY.md 8(f)-1, the row adjacent to the hot path: library convolutions (torch -> cuDNN) in
channels-last bf16 for now; the hand-written kernels start at the per-point features."""

from . import dense_fusion  # noqa: F401
from .resnet import ResNet18Extractor  # noqa: F401
