"""numpy stand-in for the subset of `jax` the reference hot path touches (see ../README.md)."""
import sys as _sys
import numpy as _np

_sys.modules[__name__ + ".numpy"] = _np          # `import jax.numpy as jnp` -> numpy
numpy = _np

from . import nn, lax, random, tree_util          # noqa: E402,F401


def jit(fn=None, **kw):
    if fn is None:
        return lambda f: f
    return fn


def tree_map(fn, tree, *rest):
    return tree_util.tree_map(fn, tree, *rest)
