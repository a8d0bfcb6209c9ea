"""Generate the golden vectors that pin `oracle/enf_ref.py` (run in the BUILD container only).

    python tests/golden/make_golden.py            # rewrites tests/golden/ref_*.npz

What runs: the reference's OWN, unmodified source files under /root/reference/enf/
(`EquivariantCrossAttentionNeF`, `EquivariantCrossAttention`, `PointwiseFFN`, `RFFNet`, every
invariant class and `get_ca_invariant`, the latent initialisers) imported with
`oracle/jaxshim` standing in for the un-installable `jax` / `flax` (numpy float64).  `nef.init`
creates the parameter tree (so the Flax-style names and shapes come out of the reference's
module structure, not from our reading of it) and `nef.apply` produces the outputs.

Each fixture holds inputs, the flattened parameter tree, the decoded field `out`, and central
finite-difference directional derivatives (float64, step 1e-6) of a fixed scalar functional
`sum(out * cot)` with respect to p, a, sigma and a random direction in parameter space, so the
oracle's autograd can be pinned to the reference's code as well.

/root/reference does not exist on the GPU box; the .npz files are committed.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "jaxshim"))
sys.path.insert(0, "/root/reference")

import jax  # noqa: E402  (the shim)
from enf.models.equivariant_cross_attention_nef import EquivariantCrossAttentionNeF  # noqa: E402
from enf.steerable_attention.invariant import get_ca_invariant, get_sa_invariant  # noqa: E402
from enf.latents import utils as latent_utils  # noqa: E402

# name -> (nef cfg, B, C, Z, coordinate kind)
CASES = {
    # mirrors config_navier_stokes.yaml (d shrunk for fixture size), periodic window
    "rel_pos_periodic": dict(invariant_type="rel_pos_periodic", num_in=2, d=32, H=2, L=8, O=1, B=2, C=24, Z=4,
                             freq=(0.05, 0.1), window=True),
    # config_diff_plane.yaml: ponita with orientation angle, non-periodic window
    "ponita": dict(invariant_type="ponita", num_in=2, d=16, H=2, L=6, O=1, B=2, C=20, Z=4,
                   freq=(0.05, 0.2), window=True),
    # config_diff_sphere.yaml: window disabled, sigma = None
    "polar_periodic": dict(invariant_type="polar_periodic", num_in=2, d=16, H=2, L=4, O=1, B=2, C=18, Z=8,
                           freq=(0.3, 0.5), window=False),
    # same invariant with the spherical bump window on
    "polar_periodic_win": dict(invariant_type="polar_periodic", num_in=2, d=16, H=2, L=4, O=2, B=1, C=18, Z=8,
                               freq=(0.3, 0.5), window=True),
    # config_shallow_water.yaml
    "latitude_periodic": dict(invariant_type="latitude_periodic", num_in=2, d=32, H=2, L=8, O=3, B=2, C=16, Z=8,
                              freq=(0.05, 0.2), window=True),
    # config_ihc.yaml: H = 3
    "ball": dict(invariant_type="ball", num_in=3, d=32, H=3, L=8, O=1, B=2, C=27, Z=5,
                 freq=(0.2, 0.5), window=True),
    "rel_pos": dict(invariant_type="rel_pos", num_in=3, d=16, H=1, L=4, O=2, B=2, C=10, Z=3,
                    freq=(0.2, 0.3), window=True),
    "norm_rel_pos": dict(invariant_type="norm_rel_pos", num_in=2, d=16, H=2, L=4, O=1, B=1, C=10, Z=4,
                         freq=(0.2, 0.3), window=True),
    "abs_pos": dict(invariant_type="abs_pos", num_in=2, d=16, H=2, L=4, O=1, B=2, C=10, Z=4,
                    freq=(0.2, 0.3), window=True),
    # BallLatInvariant: the reference's __call__ (ball_lat.py:77-87) concatenates the un-broadcast radii (B,C,1,1) / (B,1,Z,1)
    # and raises.  The fixture runs the reference's NeF, its window (ball_lat.py:36-52) and the first four invariants as
    # written, through a subclass whose __call__ is the reference's body with ONLY those two radii broadcast to (B,C,Z,1)
    # -- what ball.py:90-92 does for the same quantities (BallLatFixed below).
    "ball_lat": dict(invariant_type="ball_lat", num_in=3, d=16, H=2, L=4, O=2, B=2, C=14, Z=5,
                     freq=(0.2, 0.5), window=True),
    # latent self-attention blocks before the cross-attention block (num_layers > 0, equivariant_cross_attention_nef.py:159-167,
    # 223-226): x = p, residual, project_heads; `ponita` switches to Ponita2D (get_sa_invariant)
    "sa1_rel_pos_periodic": dict(invariant_type="rel_pos_periodic", num_in=2, d=16, H=2, L=6, O=1, B=2, C=12, Z=5,
                                 freq=(0.05, 0.1), window=True, layers=1),
    "sa2_ponita": dict(invariant_type="ponita", num_in=2, d=16, H=2, L=5, O=1, B=2, C=10, Z=4,
                       freq=(0.05, 0.2), window=True, layers=2),
    "sa1_latitude_periodic": dict(invariant_type="latitude_periodic", num_in=2, d=16, H=1, L=4, O=2, B=1, C=9, Z=8,
                                  freq=(0.05, 0.2), window=True, layers=1),
    "sa1_rel_pos_nowin": dict(invariant_type="rel_pos", num_in=2, d=16, H=2, L=4, O=1, B=2, C=8, Z=3,
                              freq=(0.2, 0.3), window=False, layers=1),
}


def flatten(tree, prefix=""):
    out = {}
    for k, v in tree.items():
        name = f"{prefix}/{k}" if prefix else k
        if isinstance(v, dict):
            out.update(flatten(v, name))
        else:
            out[name] = v
    return out


def unflatten(flat):
    out = {}
    for name, v in flat.items():
        node = out
        parts = name.split("/")
        for k in parts[:-1]:
            node = node.setdefault(k, {})
        node[parts[-1]] = v
    return out


def ball_lat_fixed():
    import jax.numpy as jnp
    from enf.steerable_attention.invariant.ball_lat import BallLatInvariant

    class BallLatFixed(BallLatInvariant):
        def __call__(self, x, p):
            B, C, Z = x.shape[0], x.shape[1], p.shape[1]
            try:                               # the reference's own body: fails on the concatenate (documented quirk)
                return super().__call__(x, p)
            except ValueError:
                pass
            phi_x = jnp.broadcast_to(x[:, :, None, 0], (B, C, Z))[..., None]
            theta_x = jnp.broadcast_to(x[:, :, None, 1], (B, C, Z))[..., None]
            phi_p = jnp.broadcast_to(p[:, None, :, 0], (B, C, Z))[..., None]
            theta_p = jnp.broadcast_to(p[:, None, :, 1], (B, C, Z))[..., None]
            r_x = jnp.broadcast_to(x[:, :, 2][:, :, None, None], (B, C, Z, 1))
            r_p = jnp.broadcast_to(p[:, :, 3][:, None, :, None], (B, C, Z, 1))
            return jnp.concatenate([theta_x, theta_p, jnp.cos(phi_x - phi_p), jnp.sin(phi_x - phi_p), r_x, r_p], axis=-1)

    return BallLatFixed()


def build_case(name, c, rng):
    cfg = types.SimpleNamespace(invariant_type=c["invariant_type"], num_in=c["num_in"])
    ca_inv = ball_lat_fixed() if c["invariant_type"] == "ball_lat" else get_ca_invariant(cfg)
    nef = EquivariantCrossAttentionNeF(
        num_hidden=c["d"], num_heads=c["H"], num_layers=c.get("layers", 0), num_out=c["O"], latent_dim=c["L"],
        self_attn_invariant=ca_inv if c["invariant_type"] == "ball_lat" else get_sa_invariant(cfg), cross_attn_invariant=ca_inv,
        embedding_type="rff", embedding_freq_multiplier=list(c["freq"]),
        condition_value_transform=True, use_gaussian_window=c["window"])
    B, C, Z = c["B"], c["C"], c["Z"]
    t = c["invariant_type"]
    # latent poses from the reference's own initialisers, then jittered
    if t in ("polar_periodic", "latitude_periodic"):
        p = latent_utils.init_positions_polar(None, (B, Z, 2))
        x = np.stack([rng.uniform(0, 2 * np.pi, (B, C)), rng.uniform(0.05, np.pi - 0.05, (B, C))], -1)
        sigma0 = 2 * np.pi / int(round((Z // 2) ** 0.5))
    elif t in ("ball", "ball_lat"):
        p = latent_utils.init_positions_ball(None, (B, Z, 4))
        x = np.stack([rng.uniform(0, 2 * np.pi, (B, C)), rng.uniform(0.05, np.pi - 0.05, (B, C)),
                      rng.uniform(0, 1, (B, C))], -1)
        sigma0 = 1.0
    elif t == "ponita":
        pos = latent_utils.init_positions_grid(None, (B, Z, 2))
        ori = latent_utils.init_ori_rotation_invariant_s2(None, (B, Z, 2))
        p = np.concatenate([pos, ori], -1)
        x = rng.uniform(-1, 1, (B, C, 2))
        sigma0 = 2 / int(round(Z ** 0.5))
    else:
        n = c["num_in"]
        p = rng.uniform(-1, 1, (B, Z, n))
        x = rng.uniform(-1, 1, (B, C, n))
        sigma0 = 0.7
    p = np.asarray(p, np.float64) + rng.standard_normal(np.shape(p)) * 0.05
    a = 1.0 + rng.standard_normal((B, Z, c["L"])) * 0.5
    sigma = sigma0 * (1 + rng.uniform(-0.3, 0.3, (B, Z, 1)))
    sig_arg = sigma if c["window"] else None

    variables = nef.init(jax.random.PRNGKey(hash(name) % 1000), x[:1], p[:1], a[:1], None if sig_arg is None else sigma[:1])
    flat = flatten(variables["params"])
    # move biases / LayerNorm affine off their trivial init so every term is exercised
    for k in flat:
        if k.endswith("bias") or k.endswith("scale"):
            flat[k] = flat[k] + rng.standard_normal(flat[k].shape) * 0.1
    variables = {"params": unflatten(flat)}

    def f(x_, p_, a_, s_, flat_):
        return nef.apply({"params": unflatten(flat_)}, x_, p_, a_, None if not c["window"] else s_)

    out = f(x, p, a, sigma, flat)
    cot = rng.standard_normal(out.shape)

    def functional(p_, a_, s_, flat_):
        return float(np.sum(f(x, p_, a_, s_, flat_) * cot))

    eps = 1e-6

    def fd_grad(arr, setter):
        g = np.zeros_like(arr)
        it = np.nditer(arr, flags=["multi_index"])
        for _ in it:
            idx = it.multi_index
            hi = arr.copy(); hi[idx] += eps
            lo = arr.copy(); lo[idx] -= eps
            g[idx] = (setter(hi) - setter(lo)) / (2 * eps)
        return g

    dp = fd_grad(p, lambda v: functional(v, a, sigma, flat))
    da = fd_grad(a, lambda v: functional(p, v, sigma, flat))
    ds = fd_grad(sigma, lambda v: functional(p, a, v, flat)) if c["window"] else np.zeros_like(sigma)
    # one random direction in parameter space (frozen RFF coefficients excluded: stop_gradient)
    direction = {k: (rng.standard_normal(v.shape) if not k.endswith("coefficients") else np.zeros_like(v))
                 for k, v in flat.items()}
    hi = {k: flat[k] + eps * direction[k] for k in flat}
    lo = {k: flat[k] - eps * direction[k] for k in flat}
    dtheta_dir = (functional(p, a, sigma, hi) - functional(p, a, sigma, lo)) / (2 * eps)

    rec = dict(x=x, p=p, a=a, sigma=sigma, out=out, cot=cot, dp=dp, da=da, dsigma=ds,
               dtheta_dir=np.float64(dtheta_dir))
    for k, v in flat.items():
        rec["param:" + k] = v
        rec["dir:" + k] = direction[k]
    meta = dict(c)
    meta["freq"] = list(c["freq"])
    rec["meta"] = np.array(repr(meta))
    return rec


def latent_fixture():
    """Outputs of the reference's own latent initialisers (enf/latents/utils.py) for the BASELINE shapes: pins
    enf_pde_b200/latents.py (the product's restatement) independently of the oracle's."""
    rec = {}
    for Z, n in ((25, 2), (64, 2), (27, 3)):
        rec[f"grid_{Z}_{n}"] = np.asarray(latent_utils.init_positions_grid(None, (2, Z, n)), np.float64)
    for Z in (18, 8, 128):
        rec[f"polar_{Z}"] = np.asarray(latent_utils.init_positions_polar(None, (2, Z, 2)), np.float64)
    for Z in (25, 256):
        rec[f"ball_{Z}"] = np.asarray(latent_utils.init_positions_ball(None, (2, Z, 4)), np.float64)
    rec["ori_25"] = np.asarray(latent_utils.init_ori_rotation_invariant_s2(None, (2, 25, 2)), np.float64)
    return rec


def main():
    np.savez_compressed(os.path.join(HERE, "latent_init.npz"), **latent_fixture())
    only = sys.argv[1:]
    for name, c in CASES.items():
        if only and name not in only:
            continue
        rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
        rec = build_case(name, c, rng)
        path = os.path.join(HERE, f"ref_{name}.npz")
        np.savez_compressed(path, **rec)
        print(f"{name:22s} out{rec['out'].shape} |out|max={np.abs(rec['out']).max():.4f} "
              f"params={sum(v.size for k, v in rec.items() if k.startswith('param:'))} -> {os.path.basename(path)}")
        print("   leaves:", ", ".join(sorted(k[6:] for k in rec if k.startswith("param:")))[:400])


if __name__ == "__main__":
    main()
