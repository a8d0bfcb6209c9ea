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


def vmap(fn, in_axes=0, out_axes=0):
    """only ever wrapped around the scalar-ODE demo solver at import time of trainer_utils/solvers.py; never called here"""
    def not_supported(*a, **k):
        raise NotImplementedError("jaxshim: vmap is an import-time stand-in only")
    return not_supported
