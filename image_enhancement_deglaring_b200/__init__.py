"""B200-native UNet de-glaring hot path (drop-in for JTZ18/image-enhancement-deglaring's model layer)."""
from .model import LightweightUNet, count_parameters, get_model_size_mb  # noqa: F401

__all__ = ["LightweightUNet", "count_parameters", "get_model_size_mb"]
from .model_optimized import OptimizedUNet  # noqa: E402,F401

__all__.append("OptimizedUNet")
from .train import FusedAdamW, L1Loss  # noqa: E402,F401

__all__ += ["FusedAdamW", "L1Loss"]
from . import imageops  # noqa: E402,F401  (device-side PIL / OpenCV pre- and post-processing)

__all__.append("imageops")
