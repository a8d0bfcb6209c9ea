"""The folded algorithm (blueprint of the CUDA kernels) against the unfused oracle, float64, CPU."""
import pytest
import torch

from oracle import enf_ref as R
from folded_model import Folded, LEAF_PATHS
from helpers import golden_names, load_golden, rel_err


@pytest.mark.parametrize("name", golden_names())
def test_folded_forward_and_backward_match_oracle(name):
    cfg, params, _, rec = load_golden(name)
    x, p, a, sigma, cot = rec["x"], rec["p"], rec["a"], rec["sigma"], rec["cot"]
    out_ref, dtheta, dp, da, dsigma = R.fwd_bwd(cfg, params, x, p, a, sigma, cot)
    m = Folded(cfg, params)
    out = m.forward(x, p, a, sigma)
    assert rel_err(out, rec["out"]) < 1e-10
    G, dp2, da2, ds2 = m.backward(x, p, a, sigma, cot)
    assert rel_err(dp2, dp) < 1e-8
    assert rel_err(da2, da) < 1e-8
    if cfg.use_gaussian_window:
        assert rel_err(ds2, dsigma) < 1e-8
    flat = R.tree_flatten(dtheta["params"])
    assert sorted(G) == sorted(LEAF_PATHS)
    for leaf, path in LEAF_PATHS.items():
        assert G[leaf].shape == flat[path].shape, leaf
        scale = max(1e-12, max(float(v.abs().max()) for v in flat.values()))
        err = float((G[leaf] - flat[path]).abs().max()) / scale
        assert err < 1e-9, (leaf, err)
