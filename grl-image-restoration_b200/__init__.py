"""grl-image-restoration_b200: B200-native (sm_100a) implementation of GRL's forward hot path behind the
reference's nn.Module surface.  See DESIGN.md / INTEGRATION.md at the repository root."""
from . import checkpoint, configs, geometry, tiling  # noqa: F401
from .modules import (  # noqa: F401
    GRL, AffineTransform, AnchorLinear, AnchorProjection, AnchorStripeAttention, CAB, ChannelAttention, CPB_MLP,
    EfficientMixAttnTransformerBlock, MixedAttention, Mlp, QKVProjection, TransformerStage, Upsample, UpsampleOneStep,
    WindowAttention, build_last_conv,
)

__all__ = ["GRL", "TransformerStage", "EfficientMixAttnTransformerBlock", "MixedAttention", "WindowAttention",
           "AnchorStripeAttention", "AffineTransform", "CAB", "ChannelAttention", "Mlp", "QKVProjection",
           "AnchorProjection", "AnchorLinear", "CPB_MLP", "Upsample", "UpsampleOneStep", "build_last_conv",
           "configs", "geometry"]
