"""Latent ODE model and fixed-step solver over the CUDA kernels of libenf_b200.so (C ABI: include/enf_ode_b200.h).

Mirrors `PonitaODEGen` (experiments/fitting/ode_models/ponita_ode_g.py:201-257: same constructor fields, `init` / `apply`
on the latent tuple `(p, a, window)`, same parameter tree `{'params': {'ponita': {...}}}`) and `_solve_latent_ode`
(experiments/fitting/trainers/trainer_utils/solvers.py:111-162).  PyTorch is plumbing (device memory, streams, the autograd
tape across solver steps); the model's arithmetic, its vector-Jacobian product and the forward roll-out run in the library.
The op is once differentiable (what jax.value_and_grad at pde_trainer.py:299,328 needs).
"""
import ctypes
import math
from typing import Dict, Union

import torch

from . import _lib
from .invariant import BaseInvariant
from .nef import _as_f32, _flatten, _ptr, _unflatten

MAX_LAYERS = 8
_LAYER_LEAVES = ("conv_k", "conv_b", "ln_g", "ln_b", "l1_w", "l1_b", "l2_w", "l2_b")
_LAYER_PATHS = {"conv_k": "conv/kernel/kernel", "conv_b": "conv/bias", "ln_g": "norm/scale", "ln_b": "norm/bias",
                "l1_w": "linear_1/kernel", "l1_b": "linear_1/bias", "l2_w": "linear_2/kernel", "l2_b": "linear_2/bias"}
_HEAD = {"kb_w0": "kernel_basis/layers_1/kernel", "kb_b0": "kernel_basis/layers_1/bias",
         "kb_w1": "kernel_basis/layers_3/kernel", "kb_b1": "kernel_basis/layers_3/bias", "stem_w": "a_stem/kernel"}
_TAIL = {"ro_scalar": "readout_scalar/layers_0/kernel", "ro_rel": "readout_vec_rel/kernel", "ro_ori": "readout_vec_ori/kernel"}


class EnfOdeDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("B", "Z", "L", "hidden", "basis", "layers", "widen", "degree", "Dx",
                                               "invariant_kind")] + [("reserved", ctypes.c_int32 * 6)]


class EnfOdeLayer(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in _LAYER_LEAVES]


class EnfOdeWeights(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_void_p) for n in _HEAD] + [("layer", EnfOdeLayer * MAX_LAYERS)] +
                [(n, ctypes.c_void_p) for n in _TAIL])


class EnfMlpOdeDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("B", "Z", "P", "L", "hidden")] + [("reserved", ctypes.c_int32 * 3)]


class EnfMlpOdeWeights(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p * 4) for n in ("a_w", "a_b", "p_w", "p_b")]


EXPORTS = ("enf_ode_workspace_bytes", "enf_ode_fwd", "enf_ode_bwd", "enf_ode_solve", "enf_mlpode_workspace_bytes", "enf_mlpode_fwd",
           "enf_mlpode_bwd")
_bound = False


def _load():
    global _bound
    lib = _lib.load()
    if not _bound:
        vp = ctypes.c_void_p
        D, W = ctypes.POINTER(EnfOdeDesc), ctypes.POINTER(EnfOdeWeights)
        lib.enf_ode_workspace_bytes.restype = ctypes.c_size_t
        lib.enf_ode_workspace_bytes.argtypes = [D]
        lib.enf_ode_fwd.restype = ctypes.c_int
        lib.enf_ode_fwd.argtypes = [D, W, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
        lib.enf_ode_bwd.restype = ctypes.c_int
        lib.enf_ode_bwd.argtypes = [D, W, vp, vp, vp, vp, W, vp, vp, vp, ctypes.c_size_t, vp]
        lib.enf_ode_solve.restype = ctypes.c_int
        lib.enf_ode_solve.argtypes = [D, W, vp, vp, ctypes.c_int32, ctypes.c_float, ctypes.c_int32, vp, vp, vp, ctypes.c_size_t, vp]
        MD, MW = ctypes.POINTER(EnfMlpOdeDesc), ctypes.POINTER(EnfMlpOdeWeights)
        lib.enf_mlpode_workspace_bytes.restype = ctypes.c_size_t
        lib.enf_mlpode_workspace_bytes.argtypes = [MD]
        lib.enf_mlpode_fwd.restype = ctypes.c_int
        lib.enf_mlpode_fwd.argtypes = [MD, MW, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
        lib.enf_mlpode_bwd.restype = ctypes.c_int
        lib.enf_mlpode_bwd.argtypes = [MD, MW, vp, vp, vp, vp, MW, vp, vp, vp, ctypes.c_size_t, vp]
        _bound = True
    return lib


def leaf_paths(num_layers: int, has_ori: bool):
    """ordered (path in the Flax tree under 'ponita') of every leaf, in EnfOdeWeights order."""
    out = list(_HEAD.values())
    for i in range(num_layers):
        out += [f"interaction_layers_{i}/{_LAYER_PATHS[n]}" for n in _LAYER_LEAVES]
    out += [_TAIL["ro_scalar"], _TAIL["ro_rel"]] + ([_TAIL["ro_ori"]] if has_ori else [])
    return out


def _weights_struct(leaves, num_layers, has_ori):
    w = EnfOdeWeights()
    it = iter(leaves)
    for n in _HEAD:
        setattr(w, n, next(it).data_ptr())
    for i in range(num_layers):
        for n in _LAYER_LEAVES:
            setattr(w.layer[i], n, next(it).data_ptr())
    w.ro_scalar = next(it).data_ptr()
    w.ro_rel = next(it).data_ptr()
    w.ro_ori = next(it).data_ptr() if has_ori else 0
    return w


class _OdeFunction(torch.autograd.Function):
    """enf_ode_fwd / enf_ode_bwd on torch's current stream."""

    @staticmethod
    def forward(ctx, meta, p, a, *leaves):
        lib = _load()
        desc_kw, has_ori = meta
        desc = EnfOdeDesc(**desc_kw)
        p, a = _as_f32(p, "p"), _as_f32(a, "a")
        leaves = [_as_f32(t, "ode parameter") for t in leaves]
        nbytes = lib.enf_ode_workspace_bytes(ctypes.byref(desc))
        if nbytes == 0:
            raise _lib.EnfLibraryError("bad ODE model description: " + lib.enf_last_error().decode())
        ws = torch.empty(nbytes, dtype=torch.uint8, device=p.device)
        dp, da = torch.empty_like(p), torch.empty_like(a)
        w = _weights_struct(leaves, desc.layers, has_ori)
        stream = ctypes.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
        with torch.cuda.device(p.device):
            rc = lib.enf_ode_fwd(ctypes.byref(desc), ctypes.byref(w), _ptr(p), _ptr(a), _ptr(dp), _ptr(da), _ptr(ws), nbytes, stream)
        _lib.check(rc, "enf_ode_fwd")
        ctx.meta, ctx.ws, ctx.nbytes = meta, ws, nbytes
        ctx.save_for_backward(p, a, *leaves)
        return dp, da

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_dp, g_da):
        lib = _load()
        desc_kw, has_ori = ctx.meta
        desc = EnfOdeDesc(**desc_kw)
        p, a, *leaves = ctx.saved_tensors
        g_dp = _as_f32(g_dp if g_dp is not None else torch.zeros_like(p), "g_dp")
        g_da = _as_f32(g_da if g_da is not None else torch.zeros_like(a), "g_da")
        need_w = any(ctx.needs_input_grad[3:])
        grads = [torch.empty_like(t) for t in leaves] if need_w else None
        gp, ga = torch.empty_like(p), torch.empty_like(a)
        w = _weights_struct(leaves, desc.layers, has_ori)
        gw = _weights_struct(grads, desc.layers, has_ori) if need_w else None
        stream = ctypes.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
        with torch.cuda.device(p.device):
            rc = lib.enf_ode_bwd(ctypes.byref(desc), ctypes.byref(w), _ptr(p), _ptr(a), _ptr(g_dp), _ptr(g_da),
                                 ctypes.byref(gw) if need_w else None, _ptr(gp), _ptr(ga), _ptr(ctx.ws), ctx.nbytes, stream)
        _lib.check(rc, "enf_ode_bwd")
        return (None, gp, ga) + (tuple(grads) if need_w else (None,) * len(leaves))


class PonitaODEGen:
    """ponita_ode_g.py:201-257.  Supported: kernel_size="global", global_pool=False, vec_num_out=1 (what get_model_pde builds,
    experiments/fitting/__init__.py:48-61)."""

    def __init__(self, num_hidden: int, num_layers: int, scalar_num_out: int, vec_num_out: int, invariant: BaseInvariant,
                 basis_dim: int, degree: int, widening_factor: int, global_pool: bool = False,
                 kernel_size: Union[float, str] = "global"):
        if kernel_size != "global":
            raise NotImplementedError("kernel_size must be 'global' (the exponential envelope of ponita_ode_g.py:158-160 is not built)")
        if global_pool or vec_num_out != 1:
            raise NotImplementedError("global_pool=False and vec_num_out=1 only (what the shipped configs use)")
        if not 1 <= num_layers <= MAX_LAYERS:
            raise ValueError(f"num_layers must be 1..{MAX_LAYERS}")
        self.num_hidden, self.num_layers, self.scalar_num_out = num_hidden, num_layers, scalar_num_out
        self.invariant, self.basis_dim, self.degree, self.widening_factor = invariant, basis_dim, degree, widening_factor
        self.has_ori = invariant.num_z_ori_dims > 0
        self.inv_dim = 3 if invariant.invariant_type == "ponita" else invariant.dim      # Ponita2D
        self.poly_dim = sum(self.inv_dim ** (k + 1) for k in range(degree + 1))
        self._paths = leaf_paths(num_layers, self.has_ori)

    # -- parameters -------------------------------------------------------------------------------------------------------------
    def init(self, rng, latents, device=None) -> Dict:
        """Parameter tree with the reference's names, shapes and initialisers (lecun-normal Dense kernels, zero biases,
        `chang_xavier_uniform` conv kernels ponita_ode_g.py:9-13, variance_scaling(1e-6) read-outs :129-137).  `rng`: int seed
        or torch.Generator."""
        p = latents[0]
        device = device if device is not None else (p.device if torch.is_tensor(p) else "cuda")
        g = rng if isinstance(rng, torch.Generator) else torch.Generator().manual_seed(int(rng))
        H, Bd, W = self.num_hidden, self.basis_dim, self.widening_factor * self.num_hidden
        S = self.scalar_num_out + (1 if self.has_ori else 0)

        def tn(i, o, scale=1.0):        # variance_scaling(scale, fan_in, truncated_normal)
            x = torch.empty(i, o)
            torch.nn.init.trunc_normal_(x, 0.0, 1.0, -2.0, 2.0, generator=g)
            return x * (math.sqrt(scale / i) / 0.87962566103423978)

        flat = {"kernel_basis/layers_1/kernel": tn(self.poly_dim, H), "kernel_basis/layers_1/bias": torch.zeros(H),
                "kernel_basis/layers_3/kernel": tn(H, Bd), "kernel_basis/layers_3/bias": torch.zeros(Bd),
                "a_stem/kernel": tn(self.scalar_num_out, H)}
        for i in range(self.num_layers):
            pre = f"interaction_layers_{i}/"
            std = math.sqrt(2.0 / (Bd + H) * Bd)
            flat.update({pre + "conv/kernel/kernel": (torch.rand(Bd, H, generator=g) * 2 - 1) * std, pre + "conv/bias": torch.zeros(H),
                         pre + "norm/scale": torch.ones(H), pre + "norm/bias": torch.zeros(H),
                         pre + "linear_1/kernel": tn(H, W), pre + "linear_1/bias": torch.zeros(W),
                         pre + "linear_2/kernel": tn(W, H), pre + "linear_2/bias": torch.zeros(H)})
        flat["readout_scalar/layers_0/kernel"] = tn(H, S, 1e-6)
        flat["readout_vec_rel/kernel"] = tn(self.inv_dim + H, 1, 1e-6)
        if self.has_ori:
            flat["readout_vec_ori/kernel"] = tn(self.inv_dim + H, 1, 1e-6)
        return {"params": {"ponita": _unflatten({k: v.to(device) for k, v in flat.items()})}}

    def _leaves(self, variables):
        tree = variables["params"] if "params" in variables else variables
        flat = _flatten(tree["ponita"] if "ponita" in tree else tree)
        try:
            return [flat[k] for k in self._paths]
        except KeyError as e:
            raise KeyError(f"ODE parameter tree is missing {e}; expected the tree produced by ode_model.init") from None

    def _desc_kw(self, p, a):
        B, Z, P = p.shape
        if P != self.invariant.pose_dim or a.shape != (B, Z, self.scalar_num_out):
            raise ValueError(f"latent shapes {tuple(p.shape)}, {tuple(a.shape)} do not match the model")
        return dict(B=B, Z=Z, L=self.scalar_num_out, hidden=self.num_hidden, basis=self.basis_dim, layers=self.num_layers,
                    widen=self.widening_factor, degree=self.degree, Dx=self.invariant.num_x_pos_dims,
                    invariant_kind=_lib.INVARIANT_KINDS[self.invariant.invariant_type])

    # -- the model ----------------------------------------------------------------------------------------------------------------
    def apply(self, variables, latents):
        """(derivative_p, derivative_a, window_der) = ode_model.apply(params, (p, a, window)); window_der is zeros_like(window)
        (None for window None), ponita_ode_g.py:252-257."""
        p, a, window = latents
        dp, da = _OdeFunction.apply((self._desc_kw(p, a), self.has_ori), p, a, *self._leaves(variables))
        return dp, da, (None if window is None else torch.zeros_like(window))

    __call__ = apply

    def solve(self, variables, latents, t0, tf, h, method="rk4"):
        """Forward-only roll-out inside the library (enf_ode_solve): validation / visualisation (pde_trainer.py:380-384, 581-585)."""
        lib = _load()
        p, a, window = latents
        if method not in ("euler", "rk4"):
            raise ValueError(f"Unknown method: {method}")
        num_steps = int((tf - t0) / h)
        with torch.no_grad():
            p, a = _as_f32(p, "p"), _as_f32(a, "a")
            leaves = [_as_f32(t.detach(), "ode parameter") for t in self._leaves(variables)]
            desc = EnfOdeDesc(**self._desc_kw(p, a))
            nbytes = lib.enf_ode_workspace_bytes(ctypes.byref(desc))
            if nbytes == 0:
                raise _lib.EnfLibraryError("bad ODE model description: " + lib.enf_last_error().decode())
            ws = torch.empty(nbytes, dtype=torch.uint8, device=p.device)
            B, Z = p.shape[:2]
            p_traj = torch.empty(B, num_steps + 1, Z, p.shape[2], device=p.device)
            a_traj = torch.empty(B, num_steps + 1, Z, a.shape[2], device=p.device)
            w = _weights_struct(leaves, desc.layers, self.has_ori)
            stream = ctypes.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
            with torch.cuda.device(p.device):
                rc = lib.enf_ode_solve(ctypes.byref(desc), ctypes.byref(w), _ptr(p), _ptr(a), num_steps, float(h),
                                       0 if method == "euler" else 1, _ptr(p_traj), _ptr(a_traj), _ptr(ws), nbytes, stream)
            _lib.check(rc, "enf_ode_solve")
            w_traj = None if window is None else window[:, None].expand(B, num_steps + 1, *window.shape[1:]).contiguous()
        return p_traj, a_traj, w_traj


_MLP_PATHS = [f"mlp_{k}/layers_{i}/{leaf}" for k in "ap" for leaf in ("kernel", "bias") for i in (0, 2, 4, 6)]   # a_w, a_b, p_w, p_b


def _mlp_struct(leaves):
    w = EnfMlpOdeWeights()
    it = iter(leaves)
    for name in ("a_w", "a_b", "p_w", "p_b"):
        arr = getattr(w, name)
        for i in range(4):
            arr[i] = next(it).data_ptr()
    return w


class _MlpOdeFunction(torch.autograd.Function):
    """enf_mlpode_fwd / enf_mlpode_bwd on torch's current stream."""

    @staticmethod
    def forward(ctx, desc_kw, p, a, *leaves):
        lib = _load()
        desc = EnfMlpOdeDesc(**desc_kw)
        p, a = _as_f32(p, "p"), _as_f32(a, "a")
        leaves = [_as_f32(t, "ode parameter") for t in leaves]
        nbytes = lib.enf_mlpode_workspace_bytes(ctypes.byref(desc))
        if nbytes == 0:
            raise _lib.EnfLibraryError("bad MLPODE description: " + lib.enf_last_error().decode())
        ws = torch.empty(nbytes, dtype=torch.uint8, device=p.device)
        dp = torch.empty(p.shape[0], p.shape[1], 2, device=p.device)
        da = torch.empty_like(a)
        w = _mlp_struct(leaves)
        stream = ctypes.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
        with torch.cuda.device(p.device):
            rc = lib.enf_mlpode_fwd(ctypes.byref(desc), ctypes.byref(w), _ptr(p), _ptr(a), _ptr(dp), _ptr(da), _ptr(ws), nbytes, stream)
        _lib.check(rc, "enf_mlpode_fwd")
        ctx.desc_kw, ctx.ws, ctx.nbytes = desc_kw, ws, nbytes
        ctx.save_for_backward(p, a, *leaves)
        return dp, da

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_dp, g_da):
        lib = _load()
        desc = EnfMlpOdeDesc(**ctx.desc_kw)
        p, a, *leaves = ctx.saved_tensors
        g_dp = _as_f32(g_dp if g_dp is not None else torch.zeros(p.shape[0], p.shape[1], 2, device=p.device), "g_dp")
        g_da = _as_f32(g_da if g_da is not None else torch.zeros_like(a), "g_da")
        need_w = any(ctx.needs_input_grad[3:])
        grads = [torch.empty_like(t) for t in leaves] if need_w else None
        gp, ga = torch.empty_like(p), torch.empty_like(a)
        w = _mlp_struct(leaves)
        gw = _mlp_struct(grads) if need_w else None
        stream = ctypes.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
        with torch.cuda.device(p.device):
            rc = lib.enf_mlpode_bwd(ctypes.byref(desc), ctypes.byref(w), _ptr(p), _ptr(a), _ptr(g_dp), _ptr(g_da),
                                    ctypes.byref(gw) if need_w else None, _ptr(gp), _ptr(ga), _ptr(ctx.ws), ctx.nbytes, stream)
        _lib.check(rc, "enf_mlpode_bwd")
        return (None, gp, ga) + (tuple(grads) if need_w else (None,) * len(leaves))


class MLPODE:
    """experiments/fitting/ode_models/mlp_ode.py:5-42 (`cfg.node.name: mlp` of get_model_pde): same constructor fields (num_layers is
    unused by the reference too: both MLPs have three hidden layers), `init` / `apply` on the latent tuple."""

    def __init__(self, num_hidden: int, num_layers: int, scalar_num_out: int, vec_num_out: int):
        if vec_num_out != 1:
            raise NotImplementedError("vec_num_out must be 1 (the pose derivative has 2 * vec_num_out components and is added to 2-D poses)")
        self.num_hidden, self.num_layers, self.scalar_num_out, self.vec_num_out = num_hidden, num_layers, scalar_num_out, vec_num_out

    def init(self, rng, latents, device=None) -> Dict:
        p, a = latents[0], latents[1]
        device = device if device is not None else (p.device if torch.is_tensor(p) else "cuda")
        g = rng if isinstance(rng, torch.Generator) else torch.Generator().manual_seed(int(rng))
        n_in, H = p.shape[-1] + a.shape[-1], self.num_hidden

        def dense(i, o):        # flax default: lecun-normal kernel, zero bias
            x = torch.empty(i, o)
            torch.nn.init.trunc_normal_(x, 0.0, 1.0, -2.0, 2.0, generator=g)
            return {"kernel": (x * (math.sqrt(1.0 / i) / 0.87962566103423978)).to(device), "bias": torch.zeros(o, device=device)}

        mlp = lambda out: {"layers_0": dense(n_in, H), "layers_2": dense(H, H), "layers_4": dense(H, H), "layers_6": dense(H, out)}
        return {"params": {"mlp_a": mlp(self.scalar_num_out), "mlp_p": mlp(2 * self.vec_num_out)}}

    def apply(self, variables, latents):
        p, a, window = latents
        flat = _flatten(variables["params"] if "params" in variables else variables)
        try:
            leaves = [flat[k] for k in _MLP_PATHS]
        except KeyError as e:
            raise KeyError(f"MLPODE parameter tree is missing {e}") from None
        B, Z, P = p.shape
        desc = dict(B=B, Z=Z, P=P, L=self.scalar_num_out, hidden=self.num_hidden)
        dp, da = _MlpOdeFunction.apply(desc, p, a, *leaves)
        return dp, da, (None if window is None else torch.zeros_like(window))

    __call__ = apply


def solve_latent_ode(f, latents, t0, tf, h, method="rk4", stop_gradient=False):
    """solvers.py:111-162 with `f(z, t)` = e.g. `lambda z, t: ode_model.apply(params, z)`: differentiable trajectories
    (B, T+1, Z, .) of p, a, window (the training path; `PonitaODEGen.solve` is the fused forward-only roll-out)."""
    num_steps = int((tf - t0) / h)
    axpy = lambda x, c, k: x if k is None else x + c * k
    traj = [tuple(latents)]
    t = t0
    for _ in range(num_steps):
        cur = traj[-1]
        if stop_gradient:
            cur = tuple(None if v is None else v.detach() for v in cur)
        if method == "euler":                      # solvers.py:73-88
            k1 = f(cur, t)
            nxt = tuple(axpy(x, h, k) for x, k in zip(cur, k1))
        elif method == "rk4":                      # solvers.py:91-108
            k1 = f(cur, t)
            k2 = f(tuple(axpy(x, 0.5 * h, k) for x, k in zip(cur, k1)), t + 0.5 * h)
            k3 = f(tuple(axpy(x, 0.5 * h, k) for x, k in zip(cur, k2)), t + 0.5 * h)
            k4 = f(tuple(axpy(x, h, k) for x, k in zip(cur, k3)), t + h)
            nxt = tuple(x if ka is None else x + (h / 6.0) * (ka + 2 * kb + 2 * kc + kd)
                        for x, ka, kb, kc, kd in zip(cur, k1, k2, k3, k4))
        else:
            raise ValueError(f"Unknown method: {method}")
        traj.append(nxt)
        t += h
    return tuple(None if traj[0][k] is None else torch.stack([s[k] for s in traj], dim=1) for k in range(3))
