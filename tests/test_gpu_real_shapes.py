"""Every BASELINE.json configuration at its REAL shape (bench.py's CONFIGS), in the mode bench.py runs it
(`precision="bf16"`: tcgen05 kernels where the shape has them, fp32 FMA kernels otherwise), against the oracle.  `-m gpu`.

The unfused fp64 oracle cannot evaluate a full-size problem, so parity uses the one size-independent property the path has:
coordinate queries are independent.  A cotangent supported on a few random rows (of every field) makes
  * the decoded field on those rows,
  * the latent gradients dp, da, dsigma of the checked fields, and
  * ALL 46 weight gradients (a sum over exactly those rows of all fields)
equal to the oracle evaluated on the sub-sampled problem, while the CUDA path still walks every tile / item / latent of the
full-size launch.  Tolerances: the bucket of the kernels that run (2e-3 tensor-core, 1e-4 fp32), per leaf for the weights."""
import types

import numpy as np
import pytest
import torch

from oracle import enf_ref as R
from helpers import rel_err, make_case, leaf_errs

pytestmark = pytest.mark.gpu

# name -> (EnfConfig kwargs, B, grid, Z, polar_grid, rows per field, oracle field chunk)
REAL = {
    "plane64": (dict(num_in=2, num_hidden=64, num_heads=2, num_out=1, latent_dim=16, invariant_type="ponita",
                     embedding_freq_multiplier=(0.05, 0.01)), 32, (64, 64), 25, None, 24, 32),
    "ns64": (dict(num_in=2, num_hidden=128, num_heads=2, num_out=1, latent_dim=16, invariant_type="rel_pos_periodic",
                  embedding_freq_multiplier=(0.05, 0.1)), 32, (64, 64), 64, None, 24, 8),
    "sphere": (dict(num_in=2, num_hidden=16, num_heads=2, num_out=1, latent_dim=4, invariant_type="polar_periodic",
                    embedding_freq_multiplier=(0.01, 0.01), use_gaussian_window=False), 8, (128, 64), 18, (6, 3), 48, 8),
    "sw192": (dict(num_in=2, num_hidden=128, num_heads=2, num_out=3, latent_dim=32, invariant_type="latitude_periodic",
                   embedding_freq_multiplier=(0.05, 0.2)), 4, (192, 96), 144, (16, 9), 32, 2),
    "ihc": (dict(num_in=3, num_hidden=32, num_heads=3, num_out=1, latent_dim=32, invariant_type="ball",
                 embedding_freq_multiplier=(0.2, 0.5)), 1, (64, 40, 40), 256, None, 48, 1),
}


def _oracle_on_rows(cfg, params, x_rows, p, a, sigma, d_rows, chunk):
    """fwd_bwd of the oracle on the sub-sampled problem, field chunk by field chunk (memory), weight gradients summed."""
    B = p.shape[0]
    outs, dps, das, dss, gsum = [], [], [], [], None
    for b0 in range(0, B, chunk):
        sl = slice(b0, min(B, b0 + chunk))
        o, g, dp, da, ds = R.fwd_bwd(cfg, params, x_rows[sl], p[sl], a[sl], sigma[sl], d_rows[sl])
        outs.append(o); dps.append(dp); das.append(da); dss.append(ds)
        flat = R.tree_flatten(g["params"])
        gsum = flat if gsum is None else {k: gsum[k] + flat[k] for k in flat}
    return torch.cat(outs), gsum, torch.cat(dps), torch.cat(das), torch.cat(dss)


@pytest.mark.parametrize("name", list(REAL))
def test_baseline_config_at_real_shape(name):
    import enf_pde_b200 as E
    from enf_pde_b200 import _lib
    kw, B, grid, Z, polar, nrows, chunk = REAL[name]
    cfg = R.EnfConfig(**kw)
    C = int(np.prod(grid))
    params, _, p, a, sigma, _ = make_case(cfg, B, 4, Z, seed=31, polar_grid=polar)
    coords = R.make_coords(cfg, grid).float().double()                      # (C, Dx), exactly representable in fp32
    g = torch.Generator().manual_seed(7)
    rows = torch.randperm(C, generator=g)[:nrows]
    d_out = torch.zeros(B, C, cfg.num_out, dtype=torch.float64)
    d_out[:, rows] = (torch.randn(B, nrows, cfg.num_out, generator=g, dtype=torch.float64) / (B * nrows)).float().double()

    inv = E.get_ca_invariant(types.SimpleNamespace(invariant_type=cfg.invariant_type, num_in=cfg.num_in))
    nef = E.EquivariantCrossAttentionNeF(cfg.num_hidden, cfg.num_heads, 0, cfg.num_out, cfg.latent_dim, inv, inv, "rff",
                                         cfg.embedding_freq_multiplier, True, cfg.use_gaussian_window, precision="bf16")
    f32 = lambda t: t.to("cuda", torch.float32)
    P = R.tree_map(lambda t: f32(t).contiguous().requires_grad_(True), params)
    pg, ag = f32(p).requires_grad_(True), f32(a).requires_grad_(True)
    sg = f32(sigma).requires_grad_(True) if cfg.use_gaussian_window else None
    out = nef.apply(P, f32(coords), pg, ag, sg)                             # one shared grid (x_batch_stride = 0), as bench.py
    out.backward(f32(d_out))
    torch.cuda.synchronize()
    desc = _lib.EnfDesc(B=B, C=C, Z=Z, d=cfg.num_hidden, H=cfg.num_heads, L=cfg.latent_dim, O=cfg.num_out, Dx=cfg.num_in,
                        invariant_kind=_lib.INVARIANT_KINDS[cfg.invariant_type], use_window=int(cfg.use_gaussian_window),
                        precision=_lib.PREC_BF16, flags=0)
    fwd_tc, bwd_tc = _lib.dispatch(desc)
    tol = 2e-3 if fwd_tc else 1e-4

    x_rows = coords[rows][None].expand(B, -1, -1)
    out_ref, g_ref, dp_ref, da_ref, ds_ref = _oracle_on_rows(cfg, params, x_rows, p, a, sigma, d_out[:, rows], chunk)
    errs = dict(out=rel_err(out.detach()[:, rows.cuda()], out_ref), dp=rel_err(pg.grad, dp_ref), da=rel_err(ag.grad, da_ref))
    if cfg.use_gaussian_window:
        errs["ds"] = rel_err(sg.grad, ds_ref)
    le = leaf_errs({k: v.grad for k, v in R.tree_flatten(P["params"]).items()}, g_ref)
    worst = max(le, key=le.get)
    errs["dtheta"] = le[worst]
    print(name, f"tcgen05 fwd/bwd = {fwd_tc}/{bwd_tc}", {k: f"{v:.2e}" for k, v in errs.items()}, "worst leaf:", worst,
          "launches:", E.last_launch_counts())
    assert all(v < tol for v in errs.values()), (errs, worst)
