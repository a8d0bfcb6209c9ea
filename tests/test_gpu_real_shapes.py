"""Every BASELINE.json configuration at its REAL shape (bench.py's CONFIGS), in the mode bench.py runs it
(`precision="bf16"`: tcgen05 kernels where the shape has them, fp32 FMA kernels otherwise), against the oracle.  `-m gpu`.

The unfused fp64 oracle cannot evaluate a full-size problem, so parity uses the one size-independent property the path has:
coordinate queries are independent.  A cotangent supported on a few random rows (of every field) makes
  * the decoded field on those rows,
  * the latent gradients dp, da, dsigma of the checked fields, and
  * ALL 46 weight gradients (a sum over exactly those rows of all fields)
equal to the oracle evaluated on the sub-sampled problem, while the CUDA path still walks every tile / item / latent of the
full-size launch.  Tolerances: the bucket of the kernels that run -- tensor-core: 2e-3 on the decoded field and the latent
gradients, weight gradients 2e-3 of the largest entry and 1e-2 per leaf; fp32: 1e-4 on everything (helpers.TOL_*)."""
import types

import numpy as np
import pytest
import torch

from oracle import enf_ref as R
from helpers import rel_err, make_case, leaf_errs, Checker, compare, TOL_FP32, TOL_TC, TOL_TC_LEAF

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
                 embedding_freq_multiplier=(0.2, 0.5)), 1, (64, 40, 40), 256, None, 256, 1),
}


class _RowsChecker(Checker):
    """Checker whose reference is the oracle on the sub-sampled problem, evaluated field chunk by field chunk (memory) with
    the weight gradients summed over chunks."""

    def __init__(self, cfg, case, chunk):
        self.chunk = chunk
        super().__init__(cfg, case, ref=self._chunked(cfg, case, chunk))

    @staticmethod
    def _chunked(cfg, case, chunk):
        params, x_rows, p, a, sigma, d_rows = case
        B = p.shape[0]
        outs, dps, das, dss, gsum = [], [], [], [], None
        for b0 in range(0, B, chunk):
            sl = slice(b0, min(B, b0 + chunk))
            o, g, dp, da, ds = R.fwd_bwd(cfg, params, x_rows[sl], p[sl], a[sl], sigma[sl], d_rows[sl])
            outs.append(o); dps.append(dp); das.append(da); dss.append(ds)
            flat = R.tree_flatten(g["params"])
            gsum = flat if gsum is None else {k: gsum[k] + flat[k] for k in flat}
        return torch.cat(outs), {"params": R.tree_unflatten(gsum)}, torch.cat(dps), torch.cat(das), torch.cat(dss)

    def allow(self):
        if self._allow is None:
            tot = None
            for shift in (4e-6, -4e-6):
                R.RELU_KINK_SHIFT[0] = shift
                try:
                    g = self._chunked(self.cfg, self.case, self.chunk)
                finally:
                    R.RELU_KINK_SHIFT[0] = 0.0
                fb, fg = R.tree_flatten(self.ref[1]["params"]), R.tree_flatten(g[1]["params"])
                cur = ({k: (fg[k] - fb[k]).abs() for k in fb}, (g[2] - self.ref[2]).abs(), (g[3] - self.ref[3]).abs(), (g[4] - self.ref[4]).abs())
                tot = cur if tot is None else ({k: tot[0][k] + cur[0][k] for k in cur[0]}, tot[1] + cur[1], tot[2] + cur[2], tot[3] + cur[3])
            self._allow = tot
        return self._allow


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
    tol, tol_leaf = (TOL_TC, TOL_TC_LEAF) if fwd_tc else (TOL_FP32, TOL_FP32)
    assert fwd_tc == bwd_tc                                              # never a 16-bit forward behind an fp32 backward

    x_rows = coords[rows][None].expand(B, -1, -1)
    chk = _RowsChecker(cfg, (params, x_rows, p, a, sigma, d_out[:, rows]), chunk)
    gf = {k: v.grad for k, v in R.tree_flatten(P["params"]).items()}
    errs, worst, ok = compare(chk, out.detach(), pg.grad, ag.grad, sg.grad if sg is not None else None, gf, tol, tol_leaf,
                              use_window=cfg.use_gaussian_window, rows=rows.cuda())
    errs["dtheta_global"] = max(leaf_errs(gf, R.tree_flatten(chk.ref[1]["params"]), floor=1.0).values())
    print(name, f"tcgen05 fwd/bwd = {fwd_tc}/{bwd_tc}", {k: f"{v:.2e}" for k, v in errs.items()}, "worst leaf:", worst,
          "kink allowance used:", chk.used_allowance, "launches:", E.last_launch_counts())
    assert ok and errs["dtheta_global"] < tol, (errs, worst)
