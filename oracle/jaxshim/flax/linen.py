"""Minimal flax.linen on numpy: enough for the reference's enf/ modules to run unmodified.

Implements Flax's module conventions the reference depends on:
  * subclasses of Module become dataclasses; `setup()` runs lazily at first use;
  * sub-modules assigned to attributes in setup() are named after the attribute, lists of
    sub-modules `<attr>_<i>`; sub-modules created inside a @compact __call__ are auto-named
    `<ClassName>_<k>`;
  * `self.param(name, init_fn, shape)` creates (init) or looks up (apply) a leaf;
  * `Dense` (kernel (in, out), lecun_normal / zeros), `LayerNorm` (epsilon 1e-6, biased variance
    computed as E[x^2]-E[x]^2, scale + bias), `Sequential`.
"""
import dataclasses
import math
import numpy as np

from jax.nn import relu, gelu  # noqa: F401  (nn.relu, nn.gelu: flax re-exports jax.nn's, approximate=True default)
from jax import random as _random

_STACK = []          # modules whose __call__ is executing (innermost last)
_CTX = {"mode": None, "key": None}


def compact(fn):
    fn._compact = True
    return fn


def _wrap_call(fn):
    def wrapper(self, *args, **kwargs):
        _STACK.append(self)
        object.__setattr__(self, "_auto_count", {})
        try:
            self._ensure_setup()
            return fn(self, *args, **kwargs)
        finally:
            _STACK.pop()
    wrapper.__wrapped__ = fn
    return wrapper


class Module:
    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        if "__call__" in cls.__dict__:
            cls.__call__ = _wrap_call(cls.__dict__["__call__"])
        dataclasses.dataclass(cls, eq=False, repr=False)

    # -- construction ---------------------------------------------------------------------
    def __post_init__(self):
        object.__setattr__(self, "_parent", None)
        object.__setattr__(self, "_name", None)
        object.__setattr__(self, "_root_scope", None)
        object.__setattr__(self, "_setup_done", False)
        object.__setattr__(self, "_in_setup", False)
        object.__setattr__(self, "_auto_count", {})
        # created inside a running (compact) __call__ -> auto-named child of that module
        if _STACK and not _STACK[-1]._in_setup:
            top = _STACK[-1]
            k = top._auto_count.get(type(self).__name__, 0)
            top._auto_count[type(self).__name__] = k + 1
            object.__setattr__(self, "_parent", top)
            object.__setattr__(self, "_name", f"{type(self).__name__}_{k}")
        # dataclass fields that already hold sub-modules (e.g. Sequential.layers)
        for f in dataclasses.fields(self):
            self._adopt(f.name, getattr(self, f.name))

    def _adopt(self, attr, value):
        if isinstance(value, Module):
            if value._parent is None:
                object.__setattr__(value, "_parent", self)
                object.__setattr__(value, "_name", attr)
        elif isinstance(value, (list, tuple)):
            for i, v in enumerate(value):
                if isinstance(v, Module) and v._parent is None:
                    object.__setattr__(v, "_parent", self)
                    object.__setattr__(v, "_name", f"{attr}_{i}")

    def __setattr__(self, name, value):
        if "_parent" in self.__dict__:          # after __post_init__
            self._adopt(name, value)
        object.__setattr__(self, name, value)

    def _ensure_setup(self):
        if not self._setup_done:
            object.__setattr__(self, "_setup_done", True)
            if hasattr(self, "setup"):
                object.__setattr__(self, "_in_setup", True)
                try:
                    self.setup()
                finally:
                    object.__setattr__(self, "_in_setup", False)

    # -- variables --------------------------------------------------------------------------
    def _path(self):
        if self._parent is None:
            return ()
        return self._parent._path() + (self._name,)

    def _scope(self, create):
        if self._parent is None:
            return self._root_scope
        parent = self._parent._scope(create)
        if parent is None:
            return None
        if self._name not in parent:
            if not create:
                return None
            parent[self._name] = {}
        return parent[self._name]

    def param(self, name, init_fn, *init_args):
        if _CTX["mode"] == "init":
            scope = self._scope(create=True)
            if name not in scope:
                key = _CTX["key"].fold("/".join(self._path() + (name,)))
                scope[name] = np.asarray(init_fn(key, *init_args), dtype=np.float64)
            return scope[name]
        scope = self._scope(create=False)
        if scope is None or name not in scope:
            raise KeyError("missing parameter " + "/".join(self._path() + (name,)))
        return scope[name]

    def _reset(self):
        """forget lazily-built children so the same module object can be re-bound."""
        for k, v in list(self.__dict__.items()):
            if k in ("_parent",):
                continue
            vs = v if isinstance(v, (list, tuple)) else [v]
            for m in vs:
                if isinstance(m, Module) and m._parent is self:
                    m._reset()
        object.__setattr__(self, "_setup_done", False)

    def init(self, key, *args, **kwargs):
        self._reset()
        object.__setattr__(self, "_root_scope", {})
        _CTX.update(mode="init", key=key if hasattr(key, "fold") else _random.PRNGKey(key))
        try:
            self(*args, **kwargs)
        finally:
            _CTX.update(mode=None, key=None)
        return {"params": self._root_scope}

    def apply(self, variables, *args, **kwargs):
        self._reset()
        object.__setattr__(self, "_root_scope", variables["params"])
        _CTX.update(mode="apply", key=None)
        try:
            return self(*args, **kwargs)
        finally:
            _CTX.update(mode=None)


# -- initializers ---------------------------------------------------------------------------
class initializers:
    @staticmethod
    def normal(stddev=1e-2):
        return lambda key, shape, dtype=float: key.rng().standard_normal(shape) * stddev

    @staticmethod
    def constant(value):
        return lambda key, shape, dtype=float: np.full(shape, value, dtype=np.float64)

    @staticmethod
    def ones(key, shape, dtype=float):
        return np.ones(shape)

    @staticmethod
    def zeros(key, shape, dtype=float):
        return np.zeros(shape)

    @staticmethod
    def variance_scaling(scale, mode, distribution):
        def init(key, shape, dtype=float):
            fan_in, fan_out = shape[-2], shape[-1]
            denom = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2}[mode]
            var = scale / denom
            rng = key.rng()
            if distribution == "normal":
                return rng.standard_normal(shape) * math.sqrt(var)
            if distribution == "uniform":
                return rng.uniform(-1, 1, shape) * math.sqrt(3 * var)
            if distribution == "truncated_normal":
                x = rng.standard_normal(shape)
                bad = np.abs(x) > 2
                while bad.any():
                    x[bad] = rng.standard_normal(int(bad.sum()))
                    bad = np.abs(x) > 2
                return x * math.sqrt(var) / 0.87962566103423978
            raise ValueError(distribution)
        return init

    @staticmethod
    def lecun_normal():
        return initializers.variance_scaling(1.0, "fan_in", "truncated_normal")


# -- layers -----------------------------------------------------------------------------------
class Dense(Module):
    features: int
    use_bias: bool = True
    kernel_init: object = initializers.lecun_normal()
    bias_init: object = initializers.zeros

    def __call__(self, x):
        kernel = self.param("kernel", self.kernel_init, (x.shape[-1], self.features))
        y = x @ kernel
        if self.use_bias:
            y = y + self.param("bias", self.bias_init, (self.features,))
        return y


class LayerNorm(Module):
    epsilon: float = 1e-6

    def __call__(self, x):
        scale = self.param("scale", initializers.ones, (x.shape[-1],))
        bias = self.param("bias", initializers.zeros, (x.shape[-1],))
        mu = x.mean(axis=-1, keepdims=True)
        var = np.maximum((x * x).mean(axis=-1, keepdims=True) - mu * mu, 0.0)   # use_fast_variance=True
        return (x - mu) / np.sqrt(var + self.epsilon) * scale + bias


class Sequential(Module):
    layers: object

    def __call__(self, x):
        for layer in self.layers:
            x = layer(x)
        return x
