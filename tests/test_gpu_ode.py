"""Latent ODE model + solver on the GPU through the C ABI (include/enf_ode_b200.h) against the fixtures produced by
the reference's own source (tests/golden/ode_*.npz) and against the oracle (oracle/ode_ref.py) at the Navier-Stokes config's real
shape.  fp32 arithmetic: the <= 1e-4 bucket of north_star (max-norm relative, helpers.rel_err), measured ~1e-6.  `-m gpu`."""
import types

import pytest
import torch

from oracle import enf_ref as R
from oracle import ode_ref as O
from helpers import load_ode_golden, ode_golden_names, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4          # fp32 bucket
TOL_LEAF = 2e-4     # per weight-gradient leaf (sums over B*Z*Z pair rows in fp32, atomics)


def _model(cfg):
    import enf_pde_b200 as E
    inv = E.get_sa_invariant(types.SimpleNamespace(invariant_type=cfg.invariant_type, num_in=cfg.num_in))
    return E.PonitaODEGen(cfg.num_hidden, cfg.num_layers, cfg.latent_dim, 1, inv, cfg.basis_dim, cfg.degree, cfg.widening_factor,
                          global_pool=False, kernel_size="global")


def _cuda(params):
    return R.tree_map(lambda t: t.to("cuda", torch.float32).contiguous().requires_grad_(True), params)


f32 = lambda t: t.to("cuda", torch.float32)


def _oracle_grads(cfg, params, p, a, cot_p, cot_a):
    P = R.tree_map(lambda t: t.clone().requires_grad_(True), params)
    pp, aa = p.clone().requires_grad_(True), a.clone().requires_grad_(True)
    dp, da = O.ponita_ode(cfg, P, pp, aa)
    ((dp * cot_p).sum() + (da * cot_a).sum()).backward()
    return dp.detach(), da.detach(), pp.grad, aa.grad, {k: v.grad for k, v in R.tree_flatten(P).items()}


def _check(cfg, params, p, a, cot_p, cot_a, want=None, gp_floor=1e-5):
    dp_o, da_o, gp_o, ga_o, gth_o = _oracle_grads(cfg, params, p, a, cot_p, cot_a)
    model = _model(cfg)
    P = _cuda({"params": params})
    pg, ag = f32(p).requires_grad_(True), f32(a).requires_grad_(True)
    dp, da, dw = model.apply(P, (pg, ag, torch.ones(p.shape[0], p.shape[1], 1, device="cuda")))
    assert dw.shape == (p.shape[0], p.shape[1], 1) and float(dw.abs().max()) == 0.0
    ((dp * f32(cot_p)).sum() + (da * f32(cot_a)).sum()).backward()
    # max-norm relative error with an absolute floor: a single latent has dp/dt = 0 and d(invariant)/dp = 0 analytically (the
    # query and latent roles cancel); float32 leaves ~1e-8 there, which is not a relative error of anything
    rel = lambda got, ref, floor=1e-5: float((got.double().cpu() - ref).abs().max()) / max(float(ref.abs().max()), floor)
    errs = dict(dp=rel(dp.detach(), dp_o), da=rel(da.detach(), da_o), gp=rel(pg.grad, gp_o, gp_floor), ga=rel(ag.grad, ga_o))
    if want is not None:        # the reference's own outputs
        errs["dp_ref"], errs["da_ref"] = rel_err(dp.detach(), want["dp"]), rel_err(da.detach(), want["da"])
    leaves = R.tree_flatten(P["params"])
    scale = max(float(g.abs().max()) for g in gth_o.values())
    leaf_errs = {}
    for k, g in gth_o.items():
        got = leaves[k].grad
        assert got is not None, k
        # per leaf, with a floor at 1e-3 of the largest leaf (a leaf whose gradient is ~0 is compared absolutely)
        leaf_errs[k] = float((got.double().cpu() - g).abs().max()) / max(float(g.abs().max()), 1e-3 * scale)
    worst = max(leaf_errs, key=leaf_errs.get)
    print(cfg.invariant_type, {k: f"{v:.2e}" for k, v in errs.items()}, "worst leaf", worst, f"{leaf_errs[worst]:.2e}")
    assert all(v < TOL for v in errs.values()), errs
    assert leaf_errs[worst] < TOL_LEAF, (worst, leaf_errs[worst])


@pytest.mark.parametrize("name", ode_golden_names())
def test_ode_fwd_bwd_against_reference_fixtures(name):
    cfg, params, _, rec = load_ode_golden(name)
    # the ball initialiser's Euler angles reach 1e2..1e3 rad: rounding them to float32 (what the kernels are given) moves them by
    # ~3e-5 rad, so that fixture is compared through the oracle on the rounded inputs only (tests/helpers.make_case does the same)
    r32 = lambda t: t.float().double()
    if name == "ball":
        _check(cfg, R.tree_map(r32, params), r32(rec["p"]), r32(rec["a"]), rec["cot_p"], rec["cot_a"])
    else:
        _check(cfg, params, rec["p"], rec["a"], rec["cot_p"], rec["cot_a"], want=rec)


def _ns_case(B=4, Z=64, seed=0):
    """config_navier_stokes.yaml `node:` block at its real sizes (hidden 128, 3 layers, basis 64, degree 3, widening 2), 64 latents."""
    cfg = O.OdeConfig(invariant_type="rel_pos_periodic", num_in=2, num_hidden=128, num_layers=3, latent_dim=16, basis_dim=64, degree=3,
                      widening_factor=2)
    params = O.ode_init(cfg, seed=seed, readout_scale=1e6)
    g = torch.Generator().manual_seed(seed)
    p = torch.as_tensor(R.init_positions_grid(B, Z, 2), dtype=torch.float64) + 0.02 * torch.randn(B, Z, 2, generator=g, dtype=torch.float64)
    a = 1.0 + 0.3 * torch.randn(B, Z, 16, generator=g, dtype=torch.float64)
    for k, v in R.tree_flatten(params).items():
        if k.endswith("bias") or k.endswith("scale"):
            v += 0.1 * torch.randn(v.shape, generator=g, dtype=torch.float64)
    return cfg, params, p, a, g


def test_ode_real_shape_navier_stokes():
    cfg, params, p, a, g = _ns_case()
    cot_p = torch.randn(p.shape, generator=g, dtype=torch.float64)
    cot_a = torch.randn(a.shape, generator=g, dtype=torch.float64)
    _check(cfg, params, p, a, cot_p, cot_a)


@pytest.mark.parametrize("method", ["euler", "rk4"])
def test_fused_rollout_matches_oracle_solver(method):
    """enf_ode_solve (forward-only roll-out in the library) vs solvers.py's loop restated in the oracle; also against the
    reference's own single step (fixtures)."""
    for name in ("ponita", "rel_pos_periodic", "latitude_periodic"):
        cfg, params, _, rec = load_ode_golden(name)
        model = _model(cfg)
        P = _cuda({"params": params})
        h, T = rec["h"], 3
        pt, at, st = model.solve(P, (f32(rec["p"]), f32(rec["a"]), f32(rec["sigma"])), 0.0, T * h, h, method)
        po, ao, so = O.solve_latent_ode(cfg, params, (rec["p"], rec["a"], rec["sigma"]), 0.0, T * h, h, method)
        assert pt.shape == po.shape and at.shape == ao.shape and st.shape == so.shape
        e = dict(p=rel_err(pt, po), a=rel_err(at, ao), s=rel_err(st, so), p1=rel_err(pt[:, 1], rec[f"{method}_p"]),
                 a1=rel_err(at[:, 1], rec[f"{method}_a"]))
        print(name, method, {k: f"{v:.2e}" for k, v in e.items()})
        assert all(v < TOL for v in e.values()), e


@pytest.mark.parametrize("method,stop", [("euler", False), ("rk4", False), ("rk4", True)])
def test_differentiable_solver_gradients(method, stop):
    """solve_latent_ode (the training path: torch tape across steps, enf_ode_bwd per model call) vs autograd through the oracle's
    unrolled solver: gradients of a trajectory functional w.r.t. the ODE parameters and the initial latents."""
    import enf_pde_b200 as E
    cfg, params, _, rec = load_ode_golden("ponita")
    g = torch.Generator().manual_seed(3)
    h, T = 0.2, 2
    cp = torch.randn(rec["p"].shape[0], T + 1, *rec["p"].shape[1:], generator=g, dtype=torch.float64)
    ca = torch.randn(rec["a"].shape[0], T + 1, *rec["a"].shape[1:], generator=g, dtype=torch.float64)
    # oracle
    Po = R.tree_map(lambda t: t.clone().requires_grad_(True), params)
    p0, a0 = rec["p"].clone().requires_grad_(True), rec["a"].clone().requires_grad_(True)
    pt, at, _ = O.solve_latent_ode(cfg, Po, (p0, a0, rec["sigma"]), 0.0, T * h, h, method, stop_gradient=stop)
    ((pt * cp).sum() + (at * ca).sum()).backward()
    # CUDA
    model = _model(cfg)
    P = _cuda({"params": params})
    pg, ag = f32(rec["p"]).requires_grad_(True), f32(rec["a"]).requires_grad_(True)
    ptc, atc, stc = E.solve_latent_ode(lambda z, t: model.apply(P, z), (pg, ag, f32(rec["sigma"])), 0.0, T * h, h, method, stop_gradient=stop)
    ((ptc * f32(cp)).sum() + (atc * f32(ca)).sum()).backward()
    errs = dict(p_traj=rel_err(ptc.detach(), pt.detach()), a_traj=rel_err(atc.detach(), at.detach()),
                gp=rel_err(pg.grad, p0.grad), ga=rel_err(ag.grad, a0.grad))
    go, gc = R.tree_flatten(Po), R.tree_flatten(P["params"])
    scale = max(float(v.grad.abs().max()) for v in go.values())
    leaf = {k: float((gc[k].grad.double().cpu() - go[k].grad).abs().max()) / max(float(go[k].grad.abs().max()), 1e-3 * scale) for k in go}
    worst = max(leaf, key=leaf.get)
    print(method, stop, {k: f"{v:.2e}" for k, v in errs.items()}, worst, f"{leaf[worst]:.2e}")
    assert all(v < TOL for v in errs.values()), errs
    assert leaf[worst] < TOL_LEAF, (worst, leaf[worst])
    assert float(stc[:, -1].sub(f32(rec["sigma"])).abs().max()) == 0.0


def test_ode_latents_only_backward_and_errors():
    """dW = NULL skips the weight gradients (frozen ODE parameters); bad descriptions fail loudly."""
    import ctypes
    from enf_pde_b200 import ode, _lib
    cfg, params, _, rec = load_ode_golden("rel_pos_periodic")
    _, _, gp_o, ga_o, _ = _oracle_grads(cfg, params, rec["p"], rec["a"], rec["cot_p"], rec["cot_a"])
    model = _model(cfg)
    P = R.tree_map(lambda t: t.to("cuda", torch.float32).contiguous(), {"params": params})      # no requires_grad
    pg, ag = f32(rec["p"]).requires_grad_(True), f32(rec["a"]).requires_grad_(True)
    dp, da, _ = model.apply(P, (pg, ag, None))
    ((dp * f32(rec["cot_p"])).sum() + (da * f32(rec["cot_a"])).sum()).backward()
    assert rel_err(pg.grad, gp_o) < TOL and rel_err(ag.grad, ga_o) < TOL
    lib = ode._load()
    bad = ode.EnfOdeDesc(B=1, Z=4, L=4, hidden=16, basis=8, layers=9, widen=2, degree=3, Dx=2, invariant_kind=3)
    assert lib.enf_ode_workspace_bytes(ctypes.byref(bad)) == 0 and b"layers" in lib.enf_last_error()
    with pytest.raises(RuntimeError):
        model.apply(P, (rec["p"].float(), rec["a"].float(), None))          # CPU tensors: no fallback


@pytest.mark.parametrize("seed", range(8))
def test_ode_random_shapes(seed):
    """ragged / degenerate sizes: one latent, Z not a multiple of the warp size, odd widths, degree 0..3, every invariant."""
    import random
    rnd = random.Random(1000 + seed)
    inv = ["rel_pos_periodic", "ponita", "polar_periodic", "latitude_periodic", "rel_pos", "norm_rel_pos", "abs_pos", "ball"][seed % 8]
    num_in = 3 if inv == "ball" else (rnd.choice([1, 2, 3]) if inv in ("rel_pos", "norm_rel_pos", "abs_pos") else 2)
    cfg = O.OdeConfig(invariant_type=inv, num_in=num_in, num_hidden=rnd.choice([4, 12, 20, 36]), num_layers=rnd.choice([1, 2, 4]),
                      latent_dim=rnd.choice([1, 3, 7]), basis_dim=rnd.choice([1, 5, 9]), degree=rnd.choice([0, 1, 2, 3]),
                      widening_factor=rnd.choice([1, 3]))
    B, Z = rnd.choice([1, 2, 5]), rnd.choice([1, 2, 33, 70])
    g = torch.Generator().manual_seed(seed)
    params = O.ode_init(cfg, seed=seed, readout_scale=1e6)
    for k, v in R.tree_flatten(params).items():
        if k.endswith("bias") or k.endswith("scale"):
            v += 0.1 * torch.randn(v.shape, generator=g, dtype=torch.float64)
    P = cfg.enf.pose_raw_dim
    if inv in ("polar_periodic", "latitude_periodic", "ball"):
        cols = [torch.rand(B, Z, generator=g, dtype=torch.float64) * 6.28, 0.2 + torch.rand(B, Z, generator=g, dtype=torch.float64) * 2.7]
        cols += [torch.rand(B, Z, generator=g, dtype=torch.float64) for _ in range(P - 2)]
        p = torch.stack(cols, -1)
    else:
        p = torch.rand(B, Z, P, generator=g, dtype=torch.float64) * 2 - 1
    a = 1.0 + 0.3 * torch.randn(B, Z, cfg.latent_dim, generator=g, dtype=torch.float64)
    r32 = lambda t: t.float().double()
    params, p, a = R.tree_map(r32, params), r32(p), r32(a)
    cot_p, cot_a = torch.randn(p.shape, generator=g, dtype=torch.float64), torch.randn(a.shape, generator=g, dtype=torch.float64)
    print(inv, num_in, cfg, B, Z)
    # (norm_rel_pos: d|p_r - p_s|/dp is undefined on the diagonal r = s -- the reference's jnp.linalg.norm gives NaN there; the
    # oracle and the kernels both use 0)
    # a single latent: the pose gradient cancels to 0 analytically between O(1) terms; float32 leaves ~1e-8 of them
    _check(cfg, params, p, a, cot_p, cot_a, gp_floor=1e-3 if Z == 1 else 1e-5)


def test_mlpode_against_reference_fixture():
    """MLPODE (`node.name: mlp`): the library's two MLPs + vector-Jacobian product vs the reference-source fixture and the oracle."""
    import enf_pde_b200 as E
    from test_oracle_ode_golden import _load_mlp
    meta, params, _, rec = _load_mlp()
    Po = R.tree_map(lambda t: t.clone().requires_grad_(True), params)
    po, ao = rec["p"].clone().requires_grad_(True), rec["a"].clone().requires_grad_(True)
    dpo, dao = O.mlp_ode(Po, po, ao)
    ((dpo * rec["cot_p"]).sum() + (dao * rec["cot_a"]).sum()).backward()
    model = E.MLPODE(meta["hidden"], 3, meta["L"], 1)
    P = _cuda({"params": params})
    pg, ag = f32(rec["p"]).requires_grad_(True), f32(rec["a"]).requires_grad_(True)
    dp, da, dw = model.apply(P, (pg, ag, f32(rec["sigma"])))
    assert float(dw.abs().max()) == 0.0
    ((dp * f32(rec["cot_p"])).sum() + (da * f32(rec["cot_a"])).sum()).backward()
    errs = dict(dp=rel_err(dp.detach(), rec["dp"]), da=rel_err(da.detach(), rec["da"]), gp=rel_err(pg.grad, po.grad), ga=rel_err(ag.grad, ao.grad))
    go, gc = R.tree_flatten(Po), R.tree_flatten(P["params"])
    leaf = {k: rel_err(gc[k].grad, go[k].grad) for k in go}
    print({k: f"{v:.2e}" for k, v in errs.items()}, max(leaf.values()))
    assert all(v < TOL for v in errs.values()) and max(leaf.values()) < TOL_LEAF, (errs, leaf)
    ours = R.tree_flatten(model.init(0, (rec["p"], rec["a"], rec["sigma"]), device="cpu")["params"])
    assert sorted(ours) == sorted(R.tree_flatten(params)) and all(tuple(ours[k].shape) == tuple(v.shape) for k, v in R.tree_flatten(params).items())


@pytest.mark.parametrize("precision,tol,tol_leaf", [("fp32", 1e-4, 2e-4), ("bf16", 2e-3, 1e-2)])
def test_ode_phase_loss_against_oracle(precision, tol, tol_leaf):
    """The ODE-phase objective of the trainer (`ode_loss`, pde_trainer.py:412-479) end to end on the GPU path: initial latents ->
    _solve_latent_ode (Euler, cfg.node.method) -> flatten (B, T) -> nef.apply on every frame -> mean squared error; gradients with
    respect to the ODE parameters, the NeF parameters and the initial latents against autograd through the oracle's restatement of
    the same graph.  fp32: both models in the 1e-4 bucket; bf16: the decode runs the tcgen05 kernels (d = 128)."""
    import enf_pde_b200 as E
    ecfg = R.EnfConfig(num_in=2, num_hidden=128, num_heads=2, num_out=1, latent_dim=8, invariant_type="rel_pos_periodic",
                       embedding_freq_multiplier=(0.05, 0.1))
    ocfg = O.OdeConfig(invariant_type="rel_pos_periodic", num_in=2, num_hidden=32, num_layers=2, latent_dim=8, basis_dim=16, degree=2,
                       widening_factor=2)
    B, Z, C, T, h = 2, 9, 150, 3, 0.5
    nparams, x, p0, a0, sigma, _ = make_case_enf(ecfg, B, C, Z)
    oparams = R.tree_map(lambda t: t.float().double(), O.ode_init(ocfg, seed=5, readout_scale=3e5))
    g = torch.Generator().manual_seed(9)
    y = torch.randn(B * (T + 1), C, 1, generator=g, dtype=torch.float64)

    # oracle
    Pn = R.tree_map(lambda t: t.clone().requires_grad_(True), nparams)
    Po = R.tree_map(lambda t: t.clone().requires_grad_(True), oparams)
    p_r, a_r = p0.clone().requires_grad_(True), a0.clone().requires_grad_(True)
    pt, at, st = O.solve_latent_ode(ocfg, Po, (p_r, a_r, sigma), 0.0, T * h, h, "euler")
    flat = lambda t: t.reshape(-1, *t.shape[2:])
    xs = x[:1].expand(B * (T + 1), -1, -1)
    loss_ref = ((R.nef_apply(ecfg, Pn, xs, flat(pt), flat(at), flat(st)) - y) ** 2).mean()
    loss_ref.backward()

    # GPU path
    ns = types.SimpleNamespace(invariant_type="rel_pos_periodic", num_in=2)
    nef = E.EquivariantCrossAttentionNeF(128, 2, 0, 1, 8, E.get_ca_invariant(ns), E.get_sa_invariant(ns), "rff", ecfg.embedding_freq_multiplier,
                                         True, True, precision=precision)
    model = _model(ocfg)
    Gn, Go = _cuda(nparams), _cuda({"params": oparams})
    pg, ag = f32(p0).requires_grad_(True), f32(a0).requires_grad_(True)
    ptc, atc, stc = E.solve_latent_ode(lambda z, t: model.apply(Go, z), (pg, ag, f32(sigma)), 0.0, T * h, h, "euler")
    out = nef.apply(Gn, f32(xs), flat(ptc), flat(atc), flat(stc))
    loss = ((out - f32(y)) ** 2).mean()
    loss.backward()

    errs = dict(loss=abs(float(loss.detach()) - float(loss_ref)) / abs(float(loss_ref)), gp=rel_err(pg.grad, p_r.grad), ga=rel_err(ag.grad, a_r.grad))
    leaf = {}
    for name, got, want in (("ode", R.tree_flatten(Go["params"]), R.tree_flatten(Po)), ("nef", R.tree_flatten(Gn["params"]), R.tree_flatten(Pn["params"]))):
        scale = max(float(v.grad.abs().max()) for v in want.values() if v.grad is not None)
        for k, v in want.items():
            if v.grad is None:                  # frozen RFF coefficients (stop_gradient, rff.py:90)
                continue
            leaf[name + ":" + k] = float((got[k].grad.double().cpu() - v.grad).abs().max()) / max(float(v.grad.abs().max()), 1e-3 * scale)
    worst = max(leaf, key=leaf.get)
    print(precision, {k: f"{v:.2e}" for k, v in errs.items()}, "worst leaf", worst, f"{leaf[worst]:.2e}")
    assert all(v < tol for v in errs.values()), errs
    assert leaf[worst] < tol_leaf, (worst, leaf[worst])


def make_case_enf(cfg, B, C, Z):
    from helpers import make_case
    return make_case(cfg, B, C, Z, seed=23)
