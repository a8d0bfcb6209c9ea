"""CPU oracle for the latent ODE model and solver.  TEST INFRASTRUCTURE ONLY (see oracle/enf_ref.py).

Unfused PyTorch restatement (fp64 for checking) of, paths relative to /root/reference:
  experiments/fitting/ode_models/ponita_ode_g.py:15-27    PolynomialFeatures
  experiments/fitting/ode_models/ponita_ode_g.py:30-50    ConvBlock  (conv -> LayerNorm -> Dense -> gelu -> Dense, no residual)
  experiments/fitting/ode_models/ponita_ode_g.py:53-87    SepGconv   ('bsc,brsc->brc' over a bias-free Dense of the kernel basis)
  experiments/fitting/ode_models/ponita_ode_g.py:90-198   PonitaGen  (kernel basis MLP, a_stem, read-outs; kernel_size "global")
  experiments/fitting/ode_models/ponita_ode_g.py:201-257  PonitaODEGen (a - 1, angle derivative = last scalar channel, d sigma = 0)
  experiments/fitting/ode_models/mlp_ode.py:5-42          MLPODE (the factory's non-equivariant baseline)
  experiments/fitting/trainers/trainer_utils/solvers.py:73-162  tree-mapped Euler / RK4 steps and the trajectory loop
  enf/steerable_attention/invariant/__init__.py:13-45     get_sa_invariant (self-attention variants: `ponita` -> Ponita2D)
Third-party semantics restated as in oracle/enf_ref.py (flax Dense / LayerNorm eps 1e-6 fast variance, gelu tanh form).

Parity pin: tests/golden/ode_*.npz, produced by executing the reference's own source over oracle/jaxshim
(tests/golden/make_golden_ode.py): model outputs, finite-difference gradients, one Euler and one RK4 step.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict

import torch

from . import enf_ref as R


@dataclass
class OdeConfig:
    """Mirror of the `node:` block + the nef's invariant (experiments/fitting/__init__.py:48-61)."""
    invariant_type: str = "rel_pos_periodic"
    num_in: int = 2
    num_hidden: int = 128
    num_layers: int = 3
    latent_dim: int = 16          # scalar_num_out
    basis_dim: int = 64
    degree: int = 3
    widening_factor: int = 2

    @property
    def enf(self) -> R.EnfConfig:
        return R.EnfConfig(num_in=self.num_in, invariant_type=self.invariant_type)

    @property
    def inv_dim(self) -> int:      # Ponita2D (ponita.py:46-86) adds the orientation dot product
        return 3 if self.invariant_type == "ponita" else self.enf.inv_dim

    @property
    def num_pos(self) -> int:
        return self.enf.num_z_pos_dims

    @property
    def num_ori(self) -> int:
        return self.enf.num_z_ori_dims

    @property
    def poly_dim(self) -> int:
        return sum(self.inv_dim ** (k + 1) for k in range(self.degree + 1))


def embed_pose(cfg: OdeConfig, p):
    """ponita_ode_g.py:149-151: (pos, angle) -> (pos, cos, sin)."""
    if cfg.num_ori > 0:
        n = cfg.num_pos
        return torch.cat([p[..., :n], torch.cos(p[..., n:]), torch.sin(p[..., n:])], dim=-1)
    return p


def sa_invariant(cfg: OdeConfig, pe):
    """invariant(p, p) of get_sa_invariant's class: (B, Z_r, Z_s, I), r = the `x` argument, s = the `p` argument."""
    if cfg.invariant_type == "ponita":      # ponita.py:75-86
        rel = pe[:, :, None, :2] - pe[:, None, :, :2]
        xo, po = pe[:, :, None, 2:], pe[:, None, :, 2:]
        i1 = rel[..., 0] * po[..., 0] + rel[..., 1] * po[..., 1]
        i2 = -rel[..., 0] * po[..., 1] + rel[..., 1] * po[..., 0]
        i3 = (xo * po).sum(dim=-1)
        return torch.stack([i1, i2, i3], dim=-1)
    return R.invariant(cfg.enf, pe, pe)


def polynomial_features(x, degree):
    """ponita_ode_g.py:22-27: [x, x (x) x, ...] with degree + 1 entries (the loop appends `degree` outer products)."""
    feats = [x]
    for _ in range(degree):
        feats.append(torch.einsum("...i,...j->...ij", feats[-1], x).reshape(*x.shape[:-1], -1))
    return torch.cat(feats, dim=-1)


def ponita_ode(cfg: OdeConfig, params: Dict, p, a):
    """PonitaODEGen.__call__: returns (dp/dt (B,Z,P_raw), da/dt (B,Z,L)); d sigma/dt = 0 (ponita_ode_g.py:252-257)."""
    P = params["ponita"]
    a = a - 1.0                                                  # :233
    pe = embed_pose(cfg, p)
    inv = sa_invariant(cfg, pe)                                   # (B,Z,Z,I)  :154
    kb = polynomial_features(inv, cfg.degree)                     # :157 Sequential[PolynomialFeatures, Dense, gelu, Dense, gelu]
    kb = R.gelu_tanh(R.dense(kb, P["kernel_basis"]["layers_1"]))
    kb = R.gelu_tanh(R.dense(kb, P["kernel_basis"]["layers_3"]))  # (B,Z,Z,basis)
    h = a @ P["a_stem"]["kernel"]                                 # :163 (no bias)
    for i in range(cfg.num_layers):                               # :166-167
        L = P[f"interaction_layers_{i}"]
        kern = kb @ L["conv"]["kernel"]["kernel"]                 # (B,Zr,Zs,hidden)  :77
        h = torch.einsum("bsc,brsc->brc", h, kern) + L["conv"]["bias"]     # :81-85
        h = R.layer_norm(h, L["norm"])
        h = R.dense(R.gelu_tanh(R.dense(h, L["linear_1"])), L["linear_2"])  # :45-49
    scalar = h @ P["readout_scalar"]["layers_0"]["kernel"]       # :170
    n = cfg.num_pos
    rel_pos = pe[:, :, None, :n] - pe[:, None, :, :n]             # :174
    B, Z = p.shape[0], p.shape[1]
    inv2 = torch.cat([inv, h[:, None, :, :].expand(B, Z, Z, h.shape[-1])], dim=-1)   # :177-179
    vec = ((inv2 @ P["readout_vec_rel"]["kernel"]) * rel_pos).mean(dim=-2)           # :181-182
    if cfg.num_ori > 0:                                           # :185-190
        p_ori = pe[:, None, :, n:].expand(rel_pos.shape)
        vec = vec + ((inv2 @ P["readout_vec_ori"]["kernel"]) * p_ori).mean(dim=-2)
        return torch.cat([vec, scalar[..., -1:]], dim=-1), scalar[..., :-1]          # :239-244
    return vec, scalar


def mlp_ode(params: Dict, p, a):
    """MLPODE.__call__ (experiments/fitting/ode_models/mlp_ode.py:30-42): (derivative_p_pos (B,Z,2), derivative_a (B,Z,L))."""
    x = torch.cat([p, a - 1.0], dim=-1)                          # :35, :38

    def mlp(t, x):
        for i in (0, 2, 4):
            x = R.gelu_tanh(R.dense(x, t[f"layers_{i}"]))
        return R.dense(x, t["layers_6"])
    return mlp(params["mlp_p"], x), mlp(params["mlp_a"], x)


def ode_step(cfg: OdeConfig, params, state, h, method):
    """solvers.py:73-108 with f = the model; the window's derivative is zero."""
    p, a, s = state
    f = lambda p_, a_: ponita_ode(cfg, params, p_, a_)
    if method == "euler":
        dp, da = f(p, a)
        return p + h * dp, a + h * da, s
    if method == "rk4":
        k1 = f(p, a)
        k2 = f(p + 0.5 * h * k1[0], a + 0.5 * h * k1[1])
        k3 = f(p + 0.5 * h * k2[0], a + 0.5 * h * k2[1])
        k4 = f(p + h * k3[0], a + h * k3[1])
        return (p + (h / 6.0) * (k1[0] + 2 * k2[0] + 2 * k3[0] + k4[0]),
                a + (h / 6.0) * (k1[1] + 2 * k2[1] + 2 * k3[1] + k4[1]), s)
    raise ValueError(f"Unknown method: {method}")


def solve_latent_ode(cfg: OdeConfig, params, latents, t0, tf, h, method="rk4", stop_gradient=False):
    """solvers.py:111-162: trajectories (B, T+1, Z, .) of p, a, sigma; T = int((tf - t0) / h)."""
    num_steps = int((tf - t0) / h)
    traj = [tuple(latents)]
    for _ in range(num_steps):
        cur = traj[-1]
        if stop_gradient:
            cur = tuple(t.detach() for t in cur)
        traj.append(ode_step(cfg, params, cur, h, method))
    return tuple(torch.stack([t[k] for t in traj], dim=1) for k in range(3))


def ode_init(cfg: OdeConfig, seed: int = 0, dtype=torch.float64, readout_scale: float = 1.0) -> Dict:
    """Parameter tree with the reference's names / shapes (tests/golden/ode_*.npz hold the reference's own).  Initialisers:
    lecun-normal Dense kernels, zero biases, `chang_xavier_uniform` for the conv kernel (ponita_ode_g.py:9-13),
    variance_scaling(1e-6) read-outs (:129,135,137; `readout_scale` lifts them for tests)."""
    import math
    import numpy as np
    rng = np.random.default_rng(seed)
    t = lambda x: torch.as_tensor(np.asarray(x), dtype=dtype)
    dn = lambda i, o, s=1.0: t(rng.standard_normal((i, o)) * math.sqrt(s / i))
    H, Bd = cfg.num_hidden, cfg.basis_dim
    P = {"kernel_basis": {"layers_1": {"kernel": dn(cfg.poly_dim, H), "bias": t(np.zeros(H))},
                          "layers_3": {"kernel": dn(H, Bd), "bias": t(np.zeros(Bd))}},
         "a_stem": {"kernel": dn(cfg.latent_dim, H)}}
    for i in range(cfg.num_layers):
        std = math.sqrt(2.0 / (Bd + H) * Bd)
        P[f"interaction_layers_{i}"] = {
            "conv": {"kernel": {"kernel": t(rng.uniform(-std, std, (Bd, H)))}, "bias": t(np.zeros(H))},
            "norm": {"scale": t(np.ones(H)), "bias": t(np.zeros(H))},
            "linear_1": {"kernel": dn(H, cfg.widening_factor * H), "bias": t(np.zeros(cfg.widening_factor * H))},
            "linear_2": {"kernel": dn(cfg.widening_factor * H, H), "bias": t(np.zeros(H))}}
    S = cfg.latent_dim + (1 if cfg.num_ori > 0 else 0)
    rs = 1e-6 * readout_scale
    P["readout_scalar"] = {"layers_0": {"kernel": dn(H, S, rs)}}
    P["readout_vec_rel"] = {"kernel": dn(cfg.inv_dim + H, 1, rs)}
    if cfg.num_ori > 0:
        P["readout_vec_ori"] = {"kernel": dn(cfg.inv_dim + H, 1, rs)}
    return {"ponita": P}
