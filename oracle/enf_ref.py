"""CPU oracle for the ENF steerable cross-attention hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`enf_pde_b200/`) may import this
module; only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` do, and only as the checker / the reported CPU baseline.

It is an *unfused* restatement, in PyTorch (fp64 for checking, fp32 for CPU timing), of the
reference graph exactly as the reference writes it -- every (B, C, Z, .) intermediate is
materialised, no algebraic folds -- so that the CUDA path (which folds aggressively) is checked
against an independent formulation.  Gradients come from torch autograd on this graph.

Reference lines restated (paths relative to /root/reference):
  enf/models/equivariant_cross_attention_nef.py:44-67, 204-235   (NeF + cross-attn block)
  enf/steerable_attention/equivariant_cross_attention.py:10-21, 74-151  (PointwiseFFN, operator)
  enf/steerable_attention/embedding/rff.py:42-47, 61-64, 84-93   (RFFNet / Layer / RFFEmbedding)
  enf/steerable_attention/invariant/_base_invariant.py:25-43 and each invariant's __call__
Third-party semantics restated (jax / flax are not vendored by the reference, nor installed
here): flax.linen.Dense (y = x @ kernel + bias, kernel (in, out)), flax.linen.LayerNorm
(eps 1e-6, biased variance, scale+bias), jax.nn.gelu (approximate=True, tanh form),
jax.nn.softmax.

Parity pin: the reference has no tests or golden vectors.  This restatement is pinned by
`tests/golden/*.npz`, produced by executing the reference's OWN source files from
/root/reference over a numpy shim of jax/flax (`oracle/jaxshim`, `tests/golden/make_golden.py`);
only the third-party primitives listed above are restated in that shim.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------

# invariant_type -> (dim I, num_x_pos_dims, num_z_pos_dims, num_z_ori_dims, window kind)
# window kinds: "np" non-periodic, "per" periodic, "sph" spherical bump (see Appendix A.2)
INVARIANTS = {
    "rel_pos": dict(window="np"),           # rel_pos.py:41          I = n
    "norm_rel_pos": dict(window="np"),      # norm_rel_pos.py:34     I = 1
    "abs_pos": dict(window="np"),           # abs_pos.py:42          I = n
    "rel_pos_periodic": dict(window="per"),  # rel_pos_periodic.py:49-58  I = 2n (n = 2)
    "ponita": dict(window="np"),            # ponita.py:31-42 (PonitaPos2D: the cross-attn variant)
    "polar_periodic": dict(window="sph"),   # polar_periodic.py:52-68
    "latitude_periodic": dict(window="sph"),  # spherical_longitude.py:69-85
    "ball": dict(window="sph"),             # ball.py:66-96
    "ball_lat": dict(window="sph"),         # ball_lat.py:66-88
}


@dataclass
class EnfConfig:
    """Mirror of the `nef:` block consumed by get_model_pde (experiments/fitting/__init__.py:14-38)."""
    num_in: int = 2
    num_hidden: int = 128
    num_heads: int = 2
    num_out: int = 1
    latent_dim: int = 16
    invariant_type: str = "rel_pos_periodic"
    embedding_freq_multiplier: Tuple[float, float] = (0.05, 0.1)
    use_gaussian_window: bool = True
    num_layers: int = 0          # latent self-attention blocks before the cross-attention block

    # derived -------------------------------------------------------------------------------
    @property
    def inv_dim(self) -> int:
        t, n = self.invariant_type, self.num_in
        return {"rel_pos": n, "norm_rel_pos": 1, "abs_pos": n, "rel_pos_periodic": 2 * n,
                "ponita": 2, "polar_periodic": 1, "latitude_periodic": 4, "ball": 5,
                "ball_lat": 6}[t]

    @property
    def sa_inv_dim(self) -> int:
        """dim of get_sa_invariant's class (invariant/__init__.py:13-45): only `ponita` differs (Ponita2D, ponita.py:46-61)."""
        return 3 if self.invariant_type == "ponita" else self.inv_dim

    @property
    def num_z_pos_dims(self) -> int:
        t, n = self.invariant_type, self.num_in
        return {"ponita": 2, "polar_periodic": 2, "latitude_periodic": 2, "ball": 4,
                "ball_lat": 4}.get(t, n)

    @property
    def num_z_ori_dims(self) -> int:
        return 1 if self.invariant_type == "ponita" else 0

    @property
    def pose_raw_dim(self) -> int:
        return self.num_z_pos_dims + self.num_z_ori_dims


# --------------------------------------------------------------------------------------------
# third-party primitives (flax / jax.nn semantics)
# --------------------------------------------------------------------------------------------

def dense(x, p):
    return x @ p["kernel"] + p["bias"]


def layer_norm(x, p, eps=1e-6):
    mu = x.mean(dim=-1, keepdim=True)
    var = (x * x).mean(dim=-1, keepdim=True) - mu * mu      # flax "fast variance" E[x^2]-E[x]^2
    var = torch.clamp(var, min=0.0)
    return (x - mu) * torch.rsqrt(var + eps) * p["scale"] + p["bias"]


def gelu_tanh(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def pointwise_ffn(x, p):
    """equivariant_cross_attention.py:15-21  Dense -> gelu -> LayerNorm -> Dense."""
    x = dense(x, p["Dense_0"])
    x = gelu_tanh(x)
    x = layer_norm(x, p["LayerNorm_0"])
    return dense(x, p["Dense_1"])


# Test instrumentation (NOT part of the reference's arithmetic): the derivative of a relu jumps at 0, so a pre-activation that
# sits within rounding distance of 0 makes the exact gradient depend on the sign of the last bit -- for the float32
# reference as much as for any re-implementation.  RELU_KINK_SHIFT moves only the DERIVATIVE's threshold (the forward value is
# untouched): tests evaluate the gradients with the threshold at -tau, 0, +tau to bound what a sign flip inside |pre| < tau
# can change (tests/helpers.kink_allowance).  0.0 = plain relu.
RELU_KINK_SHIFT = [0.0]


class _ReluKink(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x > RELU_KINK_SHIFT[0])
        return torch.clamp(x, min=0)

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        return g * mask


def rff_net(u, p):
    """rff.py:42-47 with num_layers=2: RFFEmbedding -> Layer(relu) -> linear_final."""
    omega = p["encoding"]["coefficients"].detach()          # stop_gradient, rff.py:90
    proj = (2.0 * math.pi) * u @ omega                       # self.pi = 2*pi, rff.py:80,92
    g = torch.cat([torch.sin(proj), torch.cos(proj)], dim=-1)  # [sin | cos], rff.py:84
    pre = dense(g, p["layers_0"]["linear"])
    h = torch.relu(pre) if RELU_KINK_SHIFT[0] == 0.0 else _ReluKink.apply(pre)      # rff.py:63-64
    return dense(h, p["linear_final"])                       # rff.py:46


# --------------------------------------------------------------------------------------------
# invariants + windows (Appendix A.2)
# --------------------------------------------------------------------------------------------

def _sph_unit(phi, theta):
    return torch.stack([torch.sin(theta) * torch.cos(phi), torch.sin(theta) * torch.sin(phi),
                        torch.cos(theta)], dim=-1)


def _cos_angle(x, p):
    """polar_periodic.py:52-63: cosine similarity of the (phi, theta) unit vectors, (B,C,Z,1)."""
    xv = _sph_unit(x[:, :, 0], x[:, :, 1])
    pv = _sph_unit(p[:, :, 0], p[:, :, 1])
    num = torch.einsum("bnd,bmd->bnm", xv, pv)[:, :, :, None]
    den = torch.linalg.norm(xv, dim=-1)[:, :, None, None] * torch.linalg.norm(pv, dim=-1)[:, None, :, None]
    return num / den


def invariant(cfg: EnfConfig, x, p, sa: bool = False):
    """x (B,C,Dx), p (B,Z,P) already cos/sin-embedded for 'ponita'. Returns (B,C,Z,I).  sa: the self-attention variant
    (get_sa_invariant; x is then a set of embedded poses too): only `ponita` differs (Ponita2D, ponita.py:63-86)."""
    t = cfg.invariant_type
    n = cfg.num_in
    if sa and t == "ponita":
        rel = x[:, :, None, :2] - p[:, None, :, :2]
        xo, po = x[:, :, None, 2:], p[:, None, :, 2:]
        i1 = rel[..., 0] * po[..., 0] + rel[..., 1] * po[..., 1]
        i2 = -rel[..., 0] * po[..., 1] + rel[..., 1] * po[..., 0]
        return torch.stack([i1, i2, (xo * po).sum(dim=-1)], dim=-1)
    if t == "rel_pos":
        return x[:, :, None, :n] - p[:, None, :, :n]
    if t == "norm_rel_pos":
        return torch.linalg.norm(p[:, None, :, :] - x[:, :, None, :], dim=-1, keepdim=True)
    if t == "abs_pos":
        return x[:, :, None, :].expand(x.shape[0], x.shape[1], p.shape[1], x.shape[2])
    if t == "rel_pos_periodic":
        rel = p[:, None, :, :] - x[:, :, None, :]
        return torch.cat([torch.cos(math.pi * rel), torch.sin(math.pi * rel)], dim=-1)
    if t == "ponita":
        rel = x[:, :, None, :2] - p[:, None, :, :2]
        ori = p[:, None, :, 2:]
        i1 = rel[..., 0] * ori[..., 0] + rel[..., 1] * ori[..., 1]
        i2 = -rel[..., 0] * ori[..., 1] + rel[..., 1] * ori[..., 0]
        return torch.stack([i1, i2], dim=-1)
    if t == "polar_periodic":
        return _cos_angle(x, p)
    if t in ("latitude_periodic", "ball_lat"):
        B, C, Z = x.shape[0], x.shape[1], p.shape[1]
        phi_x = x[:, :, None, 0].expand(B, C, Z)[..., None]
        th_x = x[:, :, None, 1].expand(B, C, Z)[..., None]
        phi_p = p[:, None, :, 0].expand(B, C, Z)[..., None]
        th_p = p[:, None, :, 1].expand(B, C, Z)[..., None]
        parts = [th_x, th_p, torch.cos(phi_x - phi_p), torch.sin(phi_x - phi_p)]
        if t == "ball_lat":
            # ball_lat.py:77-87 concatenates r_x[:, :, None, None] (B,C,1,1) and r_p[:, None, :, None]
            # (B,1,Z,1) WITHOUT broadcasting them, which jnp.concatenate rejects; the evident intent
            # (and what ball.py:90-92 does explicitly) is the broadcast, restated here.
            parts += [x[:, :, None, 2].expand(B, C, Z)[..., None], p[:, None, :, 3].expand(B, C, Z)[..., None]]
        return torch.cat(parts, dim=-1)
    if t == "ball":
        xv = _sph_unit(x[:, :, 0], x[:, :, 1])
        al, be, ga, r_p = p[:, :, 0], p[:, :, 1], p[:, :, 2], p[:, :, 3]
        ca, sa, cb, sb, cg, sg = (torch.cos(al), torch.sin(al), torch.cos(be), torch.sin(be),
                                  torch.cos(ga), torch.sin(ga))
        R = torch.stack([
            torch.stack([ca * cb, ca * sb * sg - sa * cg, ca * sb * cg + sa * sg], dim=-1),
            torch.stack([sa * cb, sa * sb * sg + ca * cg, sa * sb * cg - ca * sg], dim=-1),
            torch.stack([-sb, cb * sg, cb * cg], dim=-1)], dim=-2)          # (B,Z,3,3)
        inv = torch.einsum("bnij,bcj->bcni", R, xv)
        B, C, Z = x.shape[0], x.shape[1], p.shape[1]
        r_x = x[:, :, 2][:, :, None, None].expand(B, C, Z, 1)
        r_pb = r_p[:, None, :, None].expand(B, C, Z, 1)
        return torch.cat([inv, r_x, r_pb], dim=-1)
    raise ValueError(f"Unknown invariant type: {t}.")


def gaussian_window(cfg: EnfConfig, x, p, sigma):
    """(B,C,Z,1) additive logit bias.  sigma (B,Z,1)."""
    kind = INVARIANTS[cfg.invariant_type]["window"]
    if kind == "np":      # _base_invariant.py:25-33
        nz = cfg.num_z_pos_dims
        nx = cfg.num_in
        d2 = ((p[:, None, :, :nz] - x[:, :, None, :nx]) ** 2).sum(dim=-1, keepdim=True)
        return -(1.0 / sigma[:, None, :] ** 2) * d2
    if kind == "per":     # _base_invariant.py:35-43 (note the double negation -> positive sign)
        nz = cfg.num_z_pos_dims
        nx = cfg.num_in
        nrd = -(torch.cos(math.pi * (p[:, None, :, :nz] - x[:, :, None, :nx])) ** 2).sum(dim=-1, keepdim=True)
        return -(1.0 / sigma[:, None, :] ** 2) * nrd
    # spherical bump: polar_periodic.py:35-38, spherical_longitude.py:34-55, ball.py:36-52
    c = _cos_angle(x, p)
    dist = torch.arccos(torch.clamp(c, -1 + 1e-6, 1 - 1e-6))
    return torch.exp(-dist ** 2 / (2 * sigma[:, None, :, :] ** 2))


# --------------------------------------------------------------------------------------------
# the operator and the model
# --------------------------------------------------------------------------------------------

def cross_attention(cfg: EnfConfig, ap, x, p, a, sigma, sa: bool = False):
    """EquivariantCrossAttention.__call__ (equivariant_cross_attention.py:74-151) with
    condition_value_transform=True, condition_invariant_embedding=False; project_heads only changes the width of
    out_proj (:68-72), which the parameter shapes carry.  sa: built on the self-attention invariant (x = embedded poses)."""
    H, d = cfg.num_heads, cfg.num_hidden
    inv = invariant(cfg, x, p, sa)                               # :86
    inv_emb_q = rff_net(inv, ap["invariant_embedding_query"])    # :89
    q = dense(inv_emb_q, ap["inv_emb_to_q"])                     # :92
    k = dense(a, ap["a_to_k"])                                   # :93
    v = dense(a, ap["a_to_v"])                                   # :94
    inv_emb_v = rff_net(inv, ap["invariant_embedding_value"])    # :100
    v_gamma_beta = pointwise_ffn(inv_emb_v, ap["inv_emb_to_v"])  # :112
    v_gamma, v_beta = torch.chunk(v_gamma_beta, 2, dim=-1)       # :115
    v = v[:, None, :, :] * (1 + v_gamma) + v_beta                # :118
    v = v.reshape(v.shape[:-1] + (H, d))                         # :121
    v = pointwise_ffn(v, ap["inv_emb_cond_mixer"])               # :122
    q = q.reshape(q.shape[:-1] + (H, d))                         # :130
    k = k.reshape(k.shape[:-1] + (H, d))                         # :131
    att = (q * k[:, None, ...]).sum(dim=-1) * (1.0 / d ** 0.5)   # :134, scale :59
    if cfg.use_gaussian_window:
        att = att + gaussian_window(cfg, x, p, sigma)            # :138-139
    att = torch.softmax(att, dim=-2)                             # :141 (over latents)
    y = (att[..., None] * v).sum(dim=2)                          # :144
    y = y.reshape(*y.shape[:2], H * d)                           # :147
    return dense(y, ap["out_proj"]), att                         # :150


def nef_apply(cfg: EnfConfig, params: Dict, x, p, a, sigma, return_att: bool = False):
    """EquivariantCrossAttentionNeF.__call__ (equivariant_cross_attention_nef.py:204-235)."""
    P = params["params"] if "params" in params else params
    if cfg.num_z_ori_dims > 0:                                   # :214-217
        n = cfg.num_z_pos_dims
        p = torch.cat([p[:, :, :n], torch.cos(p[:, :, n:]), torch.sin(p[:, :, n:])], dim=-1)
    a = dense(a, P["latent_stem"])                               # :220
    for i in range(cfg.num_layers):                              # :223-226 latent self-attention (x = p, residual, project_heads)
        sb = P[f"self_attention_blocks_{i}"]
        y, _ = cross_attention(cfg, sb["attn"], p, p, layer_norm(a, sb["layer_norm_attn"]), sigma, sa=True)   # :56-59
        a = a + pointwise_ffn(a + y, sb["pointwise_ffn"])        # :62-64 (block) and :225
        a = gelu_tanh(a)                                         # :226
    blk = P["cross_attention_blocks_0"]
    a_norm = layer_norm(a, blk["layer_norm_attn"])               # :56
    y, att = cross_attention(cfg, blk["attn"], x, p, a_norm, sigma)   # :59
    out = pointwise_ffn(y, blk["pointwise_ffn"])                 # :66 (residual=False)
    out = gelu_tanh(out)                                         # :230
    op = P["out_proj"]                                           # :196-202, :233
    out = gelu_tanh(dense(out, op["layers_0"]))
    out = gelu_tanh(dense(out, op["layers_2"]))
    out = dense(out, op["layers_4"])
    return (out, att) if return_att else out


# --------------------------------------------------------------------------------------------
# parameter tree (Appendix A.3) with the Flax initialisers
# --------------------------------------------------------------------------------------------

def _trunc_normal(rng, shape, std):
    # jax.nn.initializers.variance_scaling(..., "truncated_normal"): N(0,1) truncated to [-2,2],
    # divided by the std of that truncated law so the result has the requested std.
    x = rng.standard_normal(shape)
    bad = np.abs(x) > 2
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(x) > 2
    return x * std / 0.87962566103423978


def _dense_default(rng, n_in, n_out):
    return {"kernel": _trunc_normal(rng, (n_in, n_out), math.sqrt(1.0 / n_in)), "bias": np.zeros(n_out)}


def _rff_params(rng, I, d, std):
    return {
        "encoding": {"coefficients": rng.standard_normal((I, d // 2)) * std},
        "layers_0": {"linear": {"kernel": rng.standard_normal((d, d)) * math.sqrt(2.0 / d),
                                "bias": rng.standard_normal(d) * 1e-6}},
        "linear_final": {"kernel": rng.uniform(-1, 1, (d, d)) * math.sqrt(3.0 * 2.0 / d),
                         "bias": rng.standard_normal(d) * 1e-6},
    }


def _ffn_params(rng, n_in, n_hidden, n_out):
    return {"Dense_0": _dense_default(rng, n_in, n_hidden),
            "LayerNorm_0": {"scale": np.ones(n_hidden), "bias": np.zeros(n_hidden)},
            "Dense_1": _dense_default(rng, n_hidden, n_out)}


def nef_init(cfg: EnfConfig, seed: int = 0, dtype=torch.float64, perturb: float = 0.0) -> Dict:
    """Parameter tree named as Flax would name it for nef.init (Appendix A.3).

    `perturb` > 0 adds N(0, perturb^2) noise to every bias and LayerNorm scale/bias so tests
    exercise the terms that are exactly zero / one at initialisation.
    """
    rng = np.random.default_rng(seed)
    d, H, L, I, O = cfg.num_hidden, cfg.num_heads, cfg.latent_dim, cfg.inv_dim, cfg.num_out
    fq, fv = cfg.embedding_freq_multiplier
    attn = {
        "invariant_embedding_query": _rff_params(rng, I, d, fq),
        "invariant_embedding_value": _rff_params(rng, I, d, fv),
        "inv_emb_to_q": _dense_default(rng, d, H * d),
        "a_to_k": _dense_default(rng, d, H * d),
        "a_to_v": _dense_default(rng, d, H * d),
        "inv_emb_to_v": _ffn_params(rng, d, d, 2 * H * d),
        "inv_emb_cond_mixer": _ffn_params(rng, d, d, d),
        "out_proj": _dense_default(rng, H * d, H * d),
    }
    def sa_block():
        Is = cfg.sa_inv_dim
        return {"layer_norm_attn": {"scale": np.ones(d), "bias": np.zeros(d)},
                "attn": {"invariant_embedding_query": _rff_params(rng, Is, d, fq),
                         "invariant_embedding_value": _rff_params(rng, Is, d, fv),
                         "inv_emb_to_q": _dense_default(rng, d, H * d), "a_to_k": _dense_default(rng, d, H * d),
                         "a_to_v": _dense_default(rng, d, H * d), "inv_emb_to_v": _ffn_params(rng, d, d, 2 * H * d),
                         "inv_emb_cond_mixer": _ffn_params(rng, d, d, d), "out_proj": _dense_default(rng, H * d, d)},
                "pointwise_ffn": _ffn_params(rng, d, d, d)}

    tree = {
        "latent_stem": _dense_default(rng, L, d),
        **{f"self_attention_blocks_{i}": sa_block() for i in range(cfg.num_layers)},
        "cross_attention_blocks_0": {
            "layer_norm_attn": {"scale": np.ones(d), "bias": np.zeros(d)},
            "attn": attn,
            "pointwise_ffn": _ffn_params(rng, H * d, H * d, H * d),
        },
        "out_proj": {"layers_0": _dense_default(rng, H * d, d),
                     "layers_2": _dense_default(rng, d, d),
                     "layers_4": _dense_default(rng, d, O)},
    }

    def conv(node, path=()):
        if isinstance(node, dict):
            return {k: conv(v, path + (k,)) for k, v in node.items()}
        arr = np.asarray(node, dtype=np.float64)
        if perturb > 0 and path[-1] in ("bias", "scale"):
            arr = arr + rng.standard_normal(arr.shape) * perturb
        return torch.tensor(arr, dtype=dtype)

    return {"params": conv(tree)}


# --------------------------------------------------------------------------------------------
# latent initialisation (enf/latents/utils.py, autodecoder.py:38-56) -- used to build realistic
# synthetic (p, a, sigma) for tests and the benchmark's CPU leg.
# --------------------------------------------------------------------------------------------

def init_positions_grid(num_signals, num_latents, num_dims):
    """enf/latents/utils.py:73-103."""
    npd = int(round(num_latents ** (1.0 / num_dims)))
    assert npd ** num_dims == num_latents, "num_latents must be a power of the number of position dimensions"
    ax = np.linspace(-1 + 1 / npd, 1 - 1 / npd, npd)
    g = np.stack(np.meshgrid(*[ax] * num_dims, indexing="ij"), axis=-1).reshape(-1, num_dims)
    return np.repeat(g[None], num_signals, axis=0)


def init_positions_polar(num_signals, n_phi, n_theta):
    """enf/latents/utils.py:36-70 generalised to an (n_phi x n_theta) grid built the same way
    (the reference only allows n_phi = 2*n_theta with n_theta^2 = Z/2)."""
    phi = np.linspace(np.pi / n_phi, 2 * np.pi - np.pi / n_phi, n_phi)
    theta = np.linspace((np.pi / 2) / n_theta, np.pi - (np.pi / 2) / n_theta, n_theta)
    g = np.stack(np.meshgrid(phi, theta, indexing="ij"), axis=-1).reshape(-1, 2)
    return np.repeat(g[None], num_signals, axis=0)


def init_positions_ball(num_signals, num_latents):
    """enf/latents/utils.py:4-33 (fibonacci Euler angles, r = 0.75)."""
    idx = np.arange(1, num_latents + 1)
    alpha = np.arccos(1 - 2 * idx / (num_latents + 1))
    beta = np.pi * (1 + 5 ** 0.5) * idx
    gamma = np.arange(num_latents) * (2 * np.pi / num_latents)
    pos = np.stack([alpha, beta, gamma, np.full(num_latents, 0.75)], axis=-1)
    return np.repeat(pos[None], num_signals, axis=0)


def init_latents(cfg: EnfConfig, num_signals: int, num_latents: int, polar_grid=None,
                 dtype=torch.float64, jitter: float = 0.0, seed: int = 0):
    """(p, a, sigma) as PositionOrientationFeatureAutodecoder would initialise them
    (autodecoder.py:21-56); `jitter` perturbs them so tests do not sit on the symmetric init."""
    t = cfg.invariant_type
    rng = np.random.default_rng(seed + 1234)
    if t in ("polar_periodic", "latitude_periodic"):
        if polar_grid is None:
            n_theta = int(round((num_latents // 2) ** 0.5))
            polar_grid = (2 * n_theta, n_theta)
        assert polar_grid[0] * polar_grid[1] == num_latents
        p = init_positions_polar(num_signals, *polar_grid)
        sigma0 = 2 * np.pi / polar_grid[1]
    elif t in ("ball", "ball_lat"):
        p = init_positions_ball(num_signals, num_latents)
        sigma0 = 1.0
    else:
        npos = cfg.num_z_pos_dims
        p = init_positions_grid(num_signals, num_latents, npos)
        sigma0 = npos / int(round(num_latents ** (1.0 / npos)))
        if cfg.num_z_ori_dims > 0:      # utils.py:106-109
            ori = np.arctan2(p[:, :, 0], p[:, :, 1])[:, :, None]
            p = np.concatenate([p, ori], axis=-1)
    a = np.ones((num_signals, num_latents, cfg.latent_dim))
    sigma = np.full((num_signals, num_latents, 1), sigma0)
    if jitter > 0:
        p = p + rng.standard_normal(p.shape) * jitter
        a = a + rng.standard_normal(a.shape) * jitter * 5
        sigma = sigma * (1 + rng.uniform(-0.3, 0.3, sigma.shape) * min(1.0, jitter * 10))
    return (torch.tensor(p, dtype=dtype), torch.tensor(a, dtype=dtype), torch.tensor(sigma, dtype=dtype))


def make_coords(cfg: EnfConfig, grid, dtype=torch.float64):
    """Coordinate grids as the fit_*.py scripts build them (fit_navier_stokes.py:32-33,
    fit_ihc.py:33-37, datasets/pdes.py:469-488)."""
    t = cfg.invariant_type
    if t in ("polar_periodic", "latitude_periodic"):
        nphi, nth = grid
        phi = np.linspace(0, 2 * np.pi, nphi, endpoint=False)
        th = np.linspace(0, np.pi, nth + 2)[1:-1]
        g = np.stack(np.meshgrid(phi, th, indexing="ij"), axis=-1).reshape(-1, 2)
    elif t in ("ball", "ball_lat"):
        nphi, nth, nr = grid
        phi = np.linspace(0, 2 * np.pi, nphi, endpoint=False)
        th = np.linspace(1e-3, np.pi, nth, endpoint=False)
        r = np.linspace(0, 1, nr)
        g = np.stack(np.meshgrid(phi, th, r, indexing="ij"), axis=-1).reshape(-1, 3)
    else:
        axes = [np.linspace(-1, 1, n) for n in grid]
        g = np.stack(np.meshgrid(*axes), axis=-1).reshape(-1, len(grid))
    return torch.tensor(g, dtype=dtype)


# --------------------------------------------------------------------------------------------
# loss / gradients helpers used by the tests and the CPU baseline
# --------------------------------------------------------------------------------------------

def tree_map(fn, tree):
    if isinstance(tree, dict):
        return {k: tree_map(fn, v) for k, v in tree.items()}
    return fn(tree)


def tree_flatten(tree, prefix=""):
    out = {}
    for k, v in tree.items():
        name = f"{prefix}/{k}" if prefix else k
        if isinstance(v, dict):
            out.update(tree_flatten(v, name))
        else:
            out[name] = v
    return out


def tree_unflatten(flat):
    out = {}
    for name, v in flat.items():
        node = out
        parts = name.split("/")
        for k in parts[:-1]:
            node = node.setdefault(k, {})
        node[parts[-1]] = v
    return out


def inner_loop(cfg: EnfConfig, params, coords, img, p, a, sigma, lrs, num_inner_steps, masks, optimize_gaussian_window=False,
               n_pos=None):
    """Restatement of PDETrainer.inner_loop (experiments/fitting/trainers/pde_trainer.py:122-235) with explicit masks:
    per step loss = mean((nef(coords[mask]) - img[:, mask])^2) (:175-185), grads w.r.t. the per-field latents times the
    batch size (:207), window gradient zeroed unless optimize_gaussian_window (:210-212), update -lr[key] * grad
    (:215-218); returns (loss on the last mask, adapted (p, a, sigma)).  `lrs`: dict p_pos, p_ori, a (L,), gaussian_window.
    Values only (the second-order outer gradient of :255 is not restated)."""
    B = img.shape[0]
    n_pos = p.shape[-1] if n_pos is None else n_pos
    lr_p = torch.empty(p.shape[-1], dtype=p.dtype)
    lr_p[:n_pos] = float(lrs["p_pos"])
    if p.shape[-1] > n_pos:
        lr_p[n_pos:] = float(lrs["p_ori"])
    lr_a = torch.as_tensor(lrs["a"], dtype=p.dtype).reshape(1, 1, -1)
    p, a, sigma = p.clone(), a.clone(), sigma.clone()
    for step in range(num_inner_steps):
        m = masks[step]
        pp, aa, ss = p.clone().requires_grad_(True), a.clone().requires_grad_(True), sigma.clone().requires_grad_(True)
        out = nef_apply(cfg, params, coords[m][None].expand(B, -1, -1), pp, aa, ss)
        loss = ((out - img[:, m]) ** 2).mean()
        gp, ga, gs = torch.autograd.grad(loss, (pp, aa, ss), allow_unused=True)
        p = p - lr_p * gp * B
        a = a - lr_a * ga * B
        if optimize_gaussian_window and gs is not None:
            sigma = sigma - float(lrs["gaussian_window"]) * gs * B
    m = masks[num_inner_steps]
    out = nef_apply(cfg, params, coords[m][None].expand(B, -1, -1), p, a, sigma)
    return ((out - img[:, m]) ** 2).mean(), (p, a, sigma)


def outer_loss_and_grads(cfg: EnfConfig, params, coords, img, p0, a0, sigma0, lrs, num_inner_steps, masks,
                         optimize_gaussian_window=False, n_pos=None):
    """The meta-learning OUTER objective and its gradient, second order included: `jax.value_and_grad(self.enf_loss)`
    (pde_trainer.py:255) through `inner_loop` (:122-235), whose steps take `jax.grad(loss_fn)` (:188-204) and are NOT
    stop-gradiented.  p0 (1,Z,P), a0 (1,Z,L), sigma0 (1,Z,1): the shared autodecoder latents, repeated over the B fields
    (:157-159); lrs: dict of tensors p_pos (1,), p_ori (1,), a (L,), gaussian_window (1,).
    Returns (loss, {"nef": param-tree grads, "p": , "a": , "gaussian_window": , "lrs": {key: grad}}) -- torch double backward
    (create_graph=True): relu'' = 0, as in JAX."""
    B = img.shape[0]
    params = tree_map(lambda t: t.detach().clone().requires_grad_(True), params)
    p0 = p0.detach().clone().requires_grad_(True)
    a0 = a0.detach().clone().requires_grad_(True)
    s0 = sigma0.detach().clone().requires_grad_(True)
    lr = {k: torch.as_tensor(v, dtype=p0.dtype).detach().clone().requires_grad_(True) for k, v in lrs.items()}
    n_pos = p0.shape[-1] if n_pos is None else n_pos
    p, a, sigma = p0.repeat(B, 1, 1), a0.repeat(B, 1, 1), s0.repeat(B, 1, 1)
    for step in range(num_inner_steps):
        m = masks[step]
        out = nef_apply(cfg, params, coords[m][None].expand(B, -1, -1), p, a, sigma)
        loss = ((out - img[:, m]) ** 2).mean()
        gp, ga, gs = torch.autograd.grad(loss, (p, a, sigma), create_graph=True, allow_unused=True)
        lr_p = torch.cat([lr["p_pos"].expand(n_pos)] + ([lr["p_ori"].expand(p.shape[-1] - n_pos)] if p.shape[-1] > n_pos else []))
        p = p - lr_p * gp * B
        a = a - lr["a"].reshape(1, 1, -1) * ga * B
        if optimize_gaussian_window and gs is not None:
            sigma = sigma - lr["gaussian_window"] * gs * B
    m = masks[num_inner_steps]
    out = nef_apply(cfg, params, coords[m][None].expand(B, -1, -1), p, a, sigma)
    loss = ((out - img[:, m]) ** 2).mean()
    leaves = list(tree_flatten(params).items())
    lr_keys = list(lr)
    inputs = [t for _, t in leaves] + [p0, a0, s0] + [lr[k] for k in lr_keys]
    grads = torch.autograd.grad(loss, inputs, allow_unused=True)
    z = lambda g, t: g if g is not None else torch.zeros_like(t)
    n = len(leaves)
    gflat = {name: z(g, t) for (name, t), g in zip(leaves, grads[:n])}
    return loss.detach(), {"nef": tree_unflatten(gflat), "p": z(grads[n], p0), "a": z(grads[n + 1], a0),
                           "gaussian_window": z(grads[n + 2], s0),
                           "lrs": {k: z(g, lr[k]) for k, g in zip(lr_keys, grads[n + 3:])}}


def fwd_bwd(cfg: EnfConfig, params, x, p, a, sigma, d_out):
    """Forward + reverse pass with cotangent d_out.  Returns out and grads wrt (params, p, a, sigma)."""
    params = tree_map(lambda t: t.detach().clone().requires_grad_(True), params)
    p = p.detach().clone().requires_grad_(True)
    a = a.detach().clone().requires_grad_(True)
    leaves = list(tree_flatten(params).items())
    inputs = [t for _, t in leaves] + [p, a]
    if cfg.use_gaussian_window:
        sigma = sigma.detach().clone().requires_grad_(True)
        inputs.append(sigma)
    out = nef_apply(cfg, params, x, p, a, sigma)
    grads = torch.autograd.grad(out, inputs, grad_outputs=d_out, allow_unused=True)
    n = len(leaves)
    gflat = {name: (g if g is not None else torch.zeros_like(t)) for (name, t), g in zip(leaves, grads[:n])}
    dsigma = grads[n + 2] if cfg.use_gaussian_window else torch.zeros_like(sigma)
    return out.detach(), tree_unflatten(gflat), grads[n], grads[n + 1], dsigma
