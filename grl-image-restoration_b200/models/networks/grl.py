"""`<pkg>.models.networks.grl` -- same dotted path as the reference's network module."""
from ...modules import GRL, TransformerStage  # noqa: F401
