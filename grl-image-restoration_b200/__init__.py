"""grl-image-restoration_b200: B200-native GRL forward hot path."""
from . import configs  # noqa: F401

__all__ = ["configs"]
