"""toc3d_b200 — B200-native (sm_100a) implementation of the ToC3D image-backbone hot path.

Public API: the two backbone plugins (registry names of the reference) and the return type.
"""
from .backbone import EVA_ViT, ToC3DEVAViT, ToC3DViTReturnType  # noqa: F401
from .configs import CONFIGS, TINY  # noqa: F401
from .neck import CPFPN  # noqa: F401
from .preprocess import ImagePreprocess  # noqa: F401

__all__ = ["ToC3DEVAViT", "EVA_ViT", "ToC3DViTReturnType", "CPFPN", "ImagePreprocess", "CONFIGS", "TINY"]
