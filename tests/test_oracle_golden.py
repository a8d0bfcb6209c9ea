"""The oracle restatement against golden vectors produced by the reference's own source
(tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

from oracle import enf_ref as R
from helpers import golden_names, load_golden, rel_err


@pytest.mark.parametrize("name", golden_names() + golden_names(sa=True))
def test_forward_matches_reference_source(name):
    cfg, params, _, rec = load_golden(name)
    out = R.nef_apply(cfg, params, rec["x"], rec["p"], rec["a"], rec["sigma"])
    assert out.shape == rec["out"].shape
    assert rel_err(out, rec["out"]) < 1e-11        # both float64


@pytest.mark.parametrize("name", golden_names() + golden_names(sa=True))
def test_autograd_matches_reference_finite_differences(name):
    cfg, params, direction, rec = load_golden(name)
    out, dtheta, dp, da, dsigma = R.fwd_bwd(cfg, params, rec["x"], rec["p"], rec["a"], rec["sigma"], rec["cot"])
    # central differences with step 1e-6 in float64: ~1e-8 absolute noise on O(1) derivatives
    assert rel_err(dp, rec["dp"]) < 2e-6
    assert rel_err(da, rec["da"]) < 2e-6
    if cfg.use_gaussian_window:
        assert rel_err(dsigma, rec["dsigma"]) < 2e-6
    flat_g = R.tree_flatten(dtheta["params"])
    flat_d = R.tree_flatten(direction)
    ddir = sum(float((flat_g[k] * flat_d[k]).sum()) for k in flat_g)
    assert abs(ddir - rec["dtheta_dir"]) < 2e-6 * max(1.0, abs(rec["dtheta_dir"]))
    # frozen RFF coefficients: stop_gradient (rff.py:90)
    for k, g in flat_g.items():
        if k.endswith("coefficients"):
            assert float(g.abs().max()) == 0.0


@pytest.mark.parametrize("name", ["rel_pos_periodic", "sa2_ponita", "sa1_latitude_periodic"])
def test_param_tree_names_match_reference_module_structure(name):
    cfg, params, _, _ = load_golden(name)
    ours = R.tree_flatten(R.nef_init(cfg)["params"])
    theirs = R.tree_flatten(params["params"])
    assert sorted(ours) == sorted(theirs)
    for k in ours:
        assert tuple(ours[k].shape) == tuple(theirs[k].shape), k


def test_latent_initialisers_match_reference_source():
    """enf_pde_b200/latents.py (product) and the oracle's copy, each against outputs of the reference's own
    enf/latents/utils.py (tests/golden/latent_init.npz, written by make_golden.py)."""
    import os
    import numpy as np
    from helpers import GOLDEN_DIR
    from enf_pde_b200 import latents as PL
    z = np.load(os.path.join(GOLDEN_DIR, "latent_init.npz"))
    for mod in (PL, R):
        for key in z.files:
            kind, *dims = key.split("_")
            want = z[key]
            if kind == "grid":
                got = mod.init_positions_grid(2, int(dims[0]), int(dims[1]))
            elif kind == "polar":
                n_theta = int(round((int(dims[0]) // 2) ** 0.5))
                got = mod.init_positions_polar(2, 2 * n_theta, n_theta)
            elif kind == "ball":
                got = mod.init_positions_ball(2, int(dims[0]))
            else:        # orientation angle of the ponita poses (utils.py:106-109)
                g = mod.init_positions_grid(2, int(dims[0]), 2)
                got = np.arctan2(g[:, :, 0], g[:, :, 1])[:, :, None]
            assert got.shape == want.shape, (mod.__name__, key)
            assert np.abs(got - want).max() < 1e-12, (mod.__name__, key)
