"""Diagnostics (GPU): per-stage errors of the library's workspace buffers against the float64 folded model for a named case.

    python tools/diag_stages.py <case> [fp32|bf16]       cases: fuzz<seed>, ball_lat, ihc_z256, plane64_sub

Prints every stage error (tests/gpu_helpers.run_stages), the dLam record error and the per-component dp / dsigma errors."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import enf_ref as R          # noqa: E402
from helpers import make_case, rel_err   # noqa: E402
import gpu_helpers                       # noqa: E402


def case(name):
    if name.startswith("fuzz"):
        from test_gpu_fuzz import _draw
        seed = int(name[4:])
        kw, B, C, Z = _draw(seed)
        cfg = R.EnfConfig(**kw)
        return (cfg,) + make_case(cfg, B, C, Z, seed=100 + seed)
    if name == "ball_lat":
        cfg = R.EnfConfig(num_in=3, num_hidden=32, num_heads=2, num_out=2, latent_dim=8, invariant_type="ball_lat",
                          embedding_freq_multiplier=(0.2, 0.5))
        return (cfg,) + make_case(cfg, 2, 150, 12, seed=3)
    if name.startswith("ihc_z"):
        Z = int(name[5:])
        cfg = R.EnfConfig(num_in=3, num_hidden=32, num_heads=3, num_out=1, latent_dim=32, invariant_type="ball",
                          embedding_freq_multiplier=(0.2, 0.5))
        params, _, p, a, sigma, _ = make_case(cfg, 1, 4, Z, seed=31)
        coords = R.make_coords(cfg, (64, 40, 40)).float().double()
        g = torch.Generator().manual_seed(7)
        rows = torch.randperm(coords.shape[0], generator=g)[:48]
        x = coords[rows][None]
        d_out = (torch.randn(1, 48, 1, generator=g, dtype=torch.float64) / 48).float().double()
        return cfg, params, x, p, a, sigma, d_out
    raise SystemExit("unknown case")


def main():
    name = sys.argv[1]
    prec = 1 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else 0
    cfg, params, x, p, a, sigma, d_out = case(name)
    res, errs = gpu_helpers.run_stages(cfg, params, x, p, a, sigma, d_out, precision=prec)
    print(f"== {name} precision={'bf16' if prec else 'fp32'} B,C,Z={x.shape[0]},{x.shape[1]},{p.shape[1]} d={cfg.num_hidden} H={cfg.num_heads} {cfg.invariant_type}")
    for k, v in errs.items():
        flag = "  <--" if v > (2e-3 if prec else 1e-4) else ""
        print(f"   {k:14s} {v:.3e}{flag}")
    out_ref, g_ref, dp_ref, da_ref, ds_ref = R.fwd_bwd(cfg, params, x, p, a, sigma, d_out)
    dp = res["dp"].double().cpu()
    P = dp.shape[-1]
    for i in range(P):
        e = (dp[..., i] - dp_ref[..., i]).abs().max().item()
        print(f"   dp[{i}] max|err| {e:.3e}  max|ref| {dp_ref[..., i].abs().max().item():.3e}  (global max {dp_ref.abs().max().item():.3e})")
    bad = (dp - dp_ref).abs().amax(-1)
    b, z = np.unravel_index(int(bad.argmax()), bad.shape)
    print(f"   worst latent (b={b}, z={z}): p = {p[b, z].tolist()}  dp = {dp[b, z].tolist()}  ref = {dp_ref[b, z].tolist()}")


if __name__ == "__main__":
    main()
