"""B200-native UNet de-glaring hot path (drop-in for JTZ18/image-enhancement-deglaring's model layer)."""
from .model import LightweightUNet, count_parameters, get_model_size_mb  # noqa: F401

__all__ = ["LightweightUNet", "count_parameters", "get_model_size_mb"]
