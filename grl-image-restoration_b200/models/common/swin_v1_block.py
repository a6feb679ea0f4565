from ...modules import Mlp, build_last_conv  # noqa: F401
