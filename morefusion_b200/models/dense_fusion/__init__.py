from .pspnet import PSPNetExtractor  # noqa: F401
