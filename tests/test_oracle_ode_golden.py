"""The latent-ODE oracle (oracle/ode_ref.py) against golden vectors produced by the reference's own PonitaODEGen and
solver step functions (tests/golden/make_golden_ode.py).  CPU only."""
import pytest
import torch

from oracle import enf_ref as R
from oracle import ode_ref as O
from helpers import load_ode_golden, ode_golden_names, rel_err


@pytest.mark.parametrize("name", ode_golden_names())
def test_ode_forward_matches_reference_source(name):
    cfg, params, _, rec = load_ode_golden(name)
    dp, da = O.ponita_ode(cfg, params, rec["p"], rec["a"])
    assert dp.shape == rec["dp"].shape and da.shape == rec["da"].shape
    assert rel_err(dp, rec["dp"]) < 1e-11 and rel_err(da, rec["da"]) < 1e-11        # both float64


@pytest.mark.parametrize("name", ode_golden_names())
def test_ode_autograd_matches_reference_finite_differences(name):
    cfg, params, direction, rec = load_ode_golden(name)
    P = R.tree_map(lambda t: t.clone().requires_grad_(True), params)
    p, a = rec["p"].clone().requires_grad_(True), rec["a"].clone().requires_grad_(True)
    dp, da = O.ponita_ode(cfg, P, p, a)
    ((dp * rec["cot_p"]).sum() + (da * rec["cot_a"]).sum()).backward()
    assert rel_err(p.grad, rec["gp"]) < 2e-6
    assert rel_err(a.grad, rec["ga"]) < 2e-6
    flat_p, flat_d = R.tree_flatten(P), R.tree_flatten(direction)
    ddir = sum(float((flat_p[k].grad * flat_d[k]).sum()) for k in flat_p)
    assert abs(ddir - rec["dtheta_dir"]) < 2e-6 * max(1.0, abs(rec["dtheta_dir"]))


@pytest.mark.parametrize("name", ode_golden_names())
@pytest.mark.parametrize("method", ["euler", "rk4"])
def test_solver_step_matches_reference_step_functions(name, method):
    cfg, params, _, rec = load_ode_golden(name)
    p1, a1, s1 = O.ode_step(cfg, params, (rec["p"], rec["a"], rec["sigma"]), rec["h"], method)
    assert rel_err(p1, rec[f"{method}_p"]) < 1e-11 and rel_err(a1, rec[f"{method}_a"]) < 1e-11
    assert torch.equal(s1, rec[f"{method}_sigma"])                                      # d sigma / dt = 0
    # the trajectory loop: (B, T + 1, Z, .), first entry the initial state, T = int((tf - t0) / h)
    pt, at, st = O.solve_latent_ode(cfg, params, (rec["p"], rec["a"], rec["sigma"]), 0.0, 2 * rec["h"], rec["h"], method)
    assert pt.shape == (rec["p"].shape[0], 3) + tuple(rec["p"].shape[1:])
    assert torch.equal(pt[:, 0], rec["p"]) and rel_err(pt[:, 1], rec[f"{method}_p"]) < 1e-11 and rel_err(at[:, 1], rec[f"{method}_a"]) < 1e-11


def test_ode_param_tree_names_match_reference_module_structure():
    cfg, params, _, _ = load_ode_golden("ponita")
    ours, theirs = R.tree_flatten(O.ode_init(cfg)), R.tree_flatten(params)
    assert sorted(ours) == sorted(theirs)
    for k in ours:
        assert tuple(ours[k].shape) == tuple(theirs[k].shape), k


def _load_mlp():
    import ast, os
    import numpy as np
    from helpers import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "mlpode.npz"))
    t = lambda k: torch.tensor(z[k], dtype=torch.float64)
    params = R.tree_unflatten({k[6:]: t(k) for k in z.files if k.startswith("param:")})
    direction = R.tree_unflatten({k[4:]: t(k) for k in z.files if k.startswith("dir:")})
    rec = {k: t(k) for k in ("p", "a", "sigma", "dp", "da", "cot_p", "cot_a", "gp", "ga")}
    rec["dtheta_dir"] = float(z["dtheta_dir"])
    return ast.literal_eval(str(z["meta"])), params, direction, rec


def test_mlpode_matches_reference_source():
    """MLPODE (mlp_ode.py) restated in the oracle vs the reference's own module over the shim: outputs and FD gradients."""
    meta, params, direction, rec = _load_mlp()
    P = R.tree_map(lambda t: t.clone().requires_grad_(True), params)
    p, a = rec["p"].clone().requires_grad_(True), rec["a"].clone().requires_grad_(True)
    dp, da = O.mlp_ode(P, p, a)
    assert rel_err(dp, rec["dp"]) < 1e-11 and rel_err(da, rec["da"]) < 1e-11
    ((dp * rec["cot_p"]).sum() + (da * rec["cot_a"]).sum()).backward()
    assert rel_err(p.grad, rec["gp"]) < 2e-6 and rel_err(a.grad, rec["ga"]) < 2e-6
    fp, fd = R.tree_flatten(P), R.tree_flatten(direction)
    ddir = sum(float((fp[k].grad * fd[k]).sum()) for k in fp)
    assert abs(ddir - rec["dtheta_dir"]) < 2e-6 * max(1.0, abs(rec["dtheta_dir"]))
