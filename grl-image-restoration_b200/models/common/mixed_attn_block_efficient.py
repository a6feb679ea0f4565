from ...geometry import _get_stripe_info  # noqa: F401
from ...modules import (AffineTransform, AnchorStripeAttention, EfficientMixAttnTransformerBlock,  # noqa: F401
                        MixedAttention, WindowAttention)
