"""Shared helpers for the test-suite (golden loading, error metrics, case construction)."""
import ast
import glob
import os

import numpy as np
import torch

from oracle import enf_ref as R

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(f)[4:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, "ref_*.npz")))


def load_golden(name, dtype=torch.float64):
    z = np.load(os.path.join(GOLDEN_DIR, f"ref_{name}.npz"))
    meta = ast.literal_eval(str(z["meta"]))
    cfg = R.EnfConfig(num_in=meta["num_in"], num_hidden=meta["d"], num_heads=meta["H"], num_out=meta["O"],
                      latent_dim=meta["L"], invariant_type=meta["invariant_type"],
                      embedding_freq_multiplier=tuple(meta["freq"]), use_gaussian_window=meta["window"])
    t = lambda k: torch.tensor(z[k], dtype=dtype)
    params = {"params": R.tree_unflatten({k[6:]: t(k) for k in z.files if k.startswith("param:")})}
    direction = R.tree_unflatten({k[4:]: t(k) for k in z.files if k.startswith("dir:")})
    rec = {k: t(k) for k in ("x", "p", "a", "sigma", "out", "cot", "dp", "da", "dsigma")}
    rec["dtheta_dir"] = float(z["dtheta_dir"])
    return cfg, params, direction, rec


def rel_err(got, want):
    """max |got - want| / max |want|  (the metric every tolerance in tests/ refers to)."""
    got = torch.as_tensor(got).double().cpu()
    want = torch.as_tensor(want).double().cpu()
    denom = want.abs().max().item()
    if denom == 0.0:
        return (got - want).abs().max().item()
    return (got - want).abs().max().item() / denom


def rms_err(got, want):
    """rms(got - want) / rms(want): the second, averaged bound kept next to rel_err's max-norm one."""
    got = torch.as_tensor(got).double().cpu()
    want = torch.as_tensor(want).double().cpu()
    denom = want.pow(2).mean().sqrt().item()
    num = (got - want).pow(2).mean().sqrt().item()
    return num if denom == 0.0 else num / denom


LEAF_FLOOR = 1e-2


def leaf_errs(got, want, floor=LEAF_FLOOR):
    """PER-LEAF weight-gradient errors: max |got - want| over a leaf divided by that leaf's own largest reference entry,
    with a floor of `floor` x the largest entry over all leaves (a leaf whose true gradient is (near) zero -- the frozen RFF
    coefficients, a bias behind a dead unit -- has no scale of its own).  `got`, `want`: flat {name: tensor} dicts."""
    gmax = max(float(v.abs().max()) for v in want.values())
    out = {}
    for k, w in want.items():
        w = w.double().cpu()
        g = torch.as_tensor(got[k]).double().cpu()
        scale = max(float(w.abs().max()), floor * gmax, 1e-300)
        out[k] = float((g - w).abs().max()) / scale
    return out


def worst_leaf(got, want, floor=LEAF_FLOOR):
    e = leaf_errs(got, want, floor)
    k = max(e, key=e.get)
    return e[k], k


def make_case(cfg, B, C, Z, seed=0, dtype=torch.float64, polar_grid=None, perturb=0.1, jitter=0.05):
    """Seeded synthetic inputs + weights for a config (poses from the reference's initialisers, jittered)."""
    g = torch.Generator().manual_seed(seed)
    params = R.nef_init(cfg, seed=seed, dtype=dtype, perturb=perturb)
    p, a, sigma = R.init_latents(cfg, B, Z, polar_grid=polar_grid, dtype=dtype, jitter=jitter, seed=seed)
    t = cfg.invariant_type
    if t in ("polar_periodic", "latitude_periodic"):
        x = torch.stack([torch.rand(B, C, generator=g, dtype=dtype) * 2 * np.pi,
                         0.05 + torch.rand(B, C, generator=g, dtype=dtype) * (np.pi - 0.1)], -1)
    elif t in ("ball", "ball_lat"):
        x = torch.stack([torch.rand(B, C, generator=g, dtype=dtype) * 2 * np.pi,
                         0.05 + torch.rand(B, C, generator=g, dtype=dtype) * (np.pi - 0.1),
                         torch.rand(B, C, generator=g, dtype=dtype)], -1)
    else:
        x = torch.rand(B, C, cfg.num_in, generator=g, dtype=dtype) * 2 - 1
    d_out = torch.randn(B, C, cfg.num_out, generator=g, dtype=dtype) / (B * C)
    # make every input exactly representable in float32, so the float64 oracle and the float32 CUDA path see
    # the SAME numbers (e.g. the ball initialiser's beta ~ 4e2 rad moves by 1e-5 rad when rounded to float32)
    r32 = lambda t: t.float().to(dtype)
    params = R.tree_map(r32, params)
    x, p, a, sigma, d_out = (r32(t) for t in (x, p, a, sigma, d_out))
    return params, x, p, a, sigma, d_out
