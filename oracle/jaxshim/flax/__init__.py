"""numpy stand-in for the subset of `flax` the reference hot path touches (see ../README.md)."""
from . import linen  # noqa: F401
