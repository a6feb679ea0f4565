from ...modules import (CAB, CPB_MLP, AnchorLinear, AnchorProjection, ChannelAttention, QKVProjection)  # noqa: F401
