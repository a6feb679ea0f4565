from ...modules import Upsample, UpsampleOneStep  # noqa: F401
