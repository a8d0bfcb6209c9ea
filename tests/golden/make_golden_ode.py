"""Golden vectors for the latent ODE model and solver steps; run in the BUILD container only.

    python tests/golden/make_golden_ode.py        # rewrites tests/golden/ode_*.npz

What runs: the reference's OWN, unmodified `PonitaODEGen` / `PonitaGen` / `SepGconv` / `ConvBlock` / `PolynomialFeatures`
(/root/reference/experiments/fitting/ode_models/ponita_ode_g.py), the self-attention invariants
(`enf.steerable_attention.invariant.get_sa_invariant`) and the tree-mapped Euler / RK4 step functions
(/root/reference/experiments/fitting/trainers/trainer_utils/solvers.py:73-108) over `oracle/jaxshim` (numpy float64).
The trajectory loop of `_solve_latent_ode` (solvers.py:111-162) only fills arrays with `.at[i].set` (no numpy equivalent);
the fixtures therefore hold single steps of the reference's step functions, which is all that loop composes.

Each fixture: inputs, flattened parameter tree, (dp/dt, da/dt), central finite-difference gradients (float64, step 1e-6)
of `sum(dp * cot_p) + sum(da * cot_a)` with respect to p, a and one random direction in parameter space, and one Euler
and one RK4 step of the latent state with step size h.
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
from enf.steerable_attention.invariant import get_sa_invariant  # noqa: E402
from enf.latents import utils as latent_utils  # noqa: E402
from experiments.fitting.ode_models.ponita_ode_g import PonitaODEGen  # noqa: E402
from experiments.fitting.trainers.trainer_utils.solvers import _euler_step_treemapped, _rk4_step_treemapped  # noqa: E402

sys.path.insert(0, HERE)
from make_golden import flatten, unflatten  # noqa: E402

# mirrors the `node:` block of the experiment YAMLs (hidden sizes shrunk for fixture size)
CASES = {
    "rel_pos_periodic": dict(invariant_type="rel_pos_periodic", num_in=2, hidden=16, layers=2, L=6, basis=8, degree=3, widen=2, B=2, Z=5),
    "ponita": dict(invariant_type="ponita", num_in=2, hidden=16, layers=2, L=5, basis=8, degree=3, widen=2, B=2, Z=4),
    "polar_periodic": dict(invariant_type="polar_periodic", num_in=2, hidden=8, layers=1, L=4, basis=8, degree=3, widen=2, B=2, Z=8),
    "latitude_periodic": dict(invariant_type="latitude_periodic", num_in=2, hidden=16, layers=3, L=4, basis=8, degree=2, widen=1, B=1, Z=8),
    "rel_pos": dict(invariant_type="rel_pos", num_in=3, hidden=8, layers=1, L=3, basis=4, degree=2, widen=2, B=2, Z=3),
    "norm_rel_pos": dict(invariant_type="norm_rel_pos", num_in=2, hidden=8, layers=2, L=3, basis=4, degree=3, widen=2, B=1, Z=4),
    "abs_pos": dict(invariant_type="abs_pos", num_in=2, hidden=8, layers=1, L=3, basis=4, degree=2, widen=2, B=2, Z=4),
    # config_ihc.yaml: the ball invariant between latent poses (x = p: Euler angles alpha, beta read as (phi, theta), gamma as the radius)
    "ball": dict(invariant_type="ball", num_in=3, hidden=8, layers=2, L=4, basis=4, degree=2, widen=2, B=2, Z=5),
}


def build_case(name, c, rng):
    cfg = types.SimpleNamespace(invariant_type=c["invariant_type"], num_in=c["num_in"])
    inv = get_sa_invariant(cfg)
    model = PonitaODEGen(num_hidden=c["hidden"], num_layers=c["layers"], scalar_num_out=c["L"], vec_num_out=1, invariant=inv,
                         basis_dim=c["basis"], degree=c["degree"], widening_factor=c["widen"], global_pool=False,
                         kernel_size="global")
    B, Z, t = c["B"], c["Z"], c["invariant_type"]
    if t in ("polar_periodic", "latitude_periodic"):
        p = latent_utils.init_positions_polar(None, (B, Z, 2))
    elif t == "ball":
        p = latent_utils.init_positions_ball(None, (B, Z, 4))
    elif t == "ponita":
        p = np.concatenate([latent_utils.init_positions_grid(None, (B, Z, 2)),
                            latent_utils.init_ori_rotation_invariant_s2(None, (B, Z, 2))], -1)
    else:
        p = rng.uniform(-1, 1, (B, Z, c["num_in"]))
    p = np.asarray(p, np.float64) + rng.standard_normal(np.shape(p)) * 0.05
    a = 1.0 + rng.standard_normal((B, Z, c["L"])) * 0.5
    sigma = 0.5 * (1 + rng.uniform(-0.3, 0.3, (B, Z, 1)))

    variables = model.init(jax.random.PRNGKey(hash(name) % 1000), (p[:1], a[:1], sigma[:1]))
    flat = flatten(variables["params"])
    for k in flat:
        if k.endswith("bias") or k.endswith("scale"):
            flat[k] = flat[k] + rng.standard_normal(flat[k].shape) * 0.1
        if "readout" in k:          # variance_scaling(1e-6): lift the read-outs to O(1) so the fixture exercises them
            flat[k] = flat[k] * 1e3 / max(np.abs(flat[k]).max(), 1e-30) * 0.3 * 1e-3 + rng.standard_normal(flat[k].shape) * 0.2

    def f(p_, a_, flat_):
        dp, da, dw = model.apply({"params": unflatten(flat_)}, (p_, a_, sigma))
        assert np.all(dw == 0) and dw.shape == sigma.shape
        return dp, da

    dp, da = f(p, a, flat)
    cot_p, cot_a = rng.standard_normal(dp.shape), rng.standard_normal(da.shape)

    def functional(p_, a_, flat_):
        o = f(p_, a_, flat_)
        return float(np.sum(o[0] * cot_p) + np.sum(o[1] * cot_a))

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

    gp = fd_grad(p, lambda v: functional(v, a, flat))
    ga = fd_grad(a, lambda v: functional(p, v, flat))
    direction = {k: rng.standard_normal(v.shape) for k, v in flat.items()}
    hi = {k: flat[k] + eps * direction[k] for k in flat}
    lo = {k: flat[k] - eps * direction[k] for k in flat}
    dtheta_dir = (functional(p, a, hi) - functional(p, a, lo)) / (2 * eps)

    # one step of the reference's own step functions (solvers.py:73-108) with f = the model
    h = 0.25
    fz = lambda z, t_: model.apply({"params": unflatten(flat)}, z)
    e_p, e_a, e_s = _euler_step_treemapped(fz, (p, a, sigma), 0.0, h)
    r_p, r_a, r_s = _rk4_step_treemapped(fz, (p, a, sigma), 0.0, h)

    rec = dict(p=p, a=a, sigma=sigma, dp=dp, da=da, cot_p=cot_p, cot_a=cot_a, gp=gp, ga=ga, dtheta_dir=np.float64(dtheta_dir),
               h=np.float64(h), euler_p=e_p, euler_a=e_a, euler_sigma=e_s, rk4_p=r_p, rk4_a=r_a, rk4_sigma=r_s)
    for k, v in flat.items():
        rec["param:" + k] = v
        rec["dir:" + k] = direction[k]
    rec["meta"] = np.array(repr(dict(c)))
    return rec


def build_mlp_case(rng):
    """MLPODE (experiments/fitting/ode_models/mlp_ode.py), the `node.name: mlp` option of get_model_pde."""
    from experiments.fitting.ode_models.mlp_ode import MLPODE
    B, Z, L, hidden = 2, 5, 6, 12
    model = MLPODE(num_hidden=hidden, num_layers=3, scalar_num_out=L, vec_num_out=1)
    p = rng.uniform(-1, 1, (B, Z, 2))
    a = 1.0 + rng.standard_normal((B, Z, L)) * 0.5
    sigma = np.full((B, Z, 1), 0.3)
    variables = model.init(jax.random.PRNGKey(11), (p[:1], a[:1], sigma[:1]))
    flat = flatten(variables["params"])
    for k in flat:
        if k.endswith("bias"):
            flat[k] = flat[k] + rng.standard_normal(flat[k].shape) * 0.1

    def f(p_, a_, flat_):
        dp, da, dw = model.apply({"params": unflatten(flat_)}, (p_, a_, sigma))
        assert np.all(dw == 0)
        return dp, da

    dp, da = f(p, a, flat)
    cot_p, cot_a = rng.standard_normal(dp.shape), rng.standard_normal(da.shape)
    fun = lambda p_, a_, fl: float(np.sum(f(p_, a_, fl)[0] * cot_p) + np.sum(f(p_, a_, fl)[1] * cot_a))
    eps = 1e-6

    def fd(arr, setter):
        g = np.zeros_like(arr)
        it = np.nditer(arr, flags=["multi_index"])
        for _ in it:
            i = it.multi_index
            hi = arr.copy(); hi[i] += eps
            lo = arr.copy(); lo[i] -= eps
            g[i] = (setter(hi) - setter(lo)) / (2 * eps)
        return g

    gp, ga = fd(p, lambda v: fun(v, a, flat)), fd(a, lambda v: fun(p, v, flat))
    direction = {k: rng.standard_normal(v.shape) for k, v in flat.items()}
    dth = (fun(p, a, {k: flat[k] + eps * direction[k] for k in flat}) - fun(p, a, {k: flat[k] - eps * direction[k] for k in flat})) / (2 * eps)
    rec = dict(p=p, a=a, sigma=sigma, dp=dp, da=da, cot_p=cot_p, cot_a=cot_a, gp=gp, ga=ga, dtheta_dir=np.float64(dth))
    for k, v in flat.items():
        rec["param:" + k] = v
        rec["dir:" + k] = direction[k]
    rec["meta"] = np.array(repr(dict(hidden=hidden, L=L, B=B, Z=Z)))
    return rec


if __name__ == "__main__":
    only = sys.argv[1:]
    if not only or "mlp" in only:
        rec = build_mlp_case(np.random.default_rng(4242))
        np.savez_compressed(os.path.join(HERE, "mlpode.npz"), **rec)
        print("mlpode dp", rec["dp"].shape, "da", rec["da"].shape, sorted(k[6:] for k in rec if k.startswith("param:"))[:4])
    for name, c in CASES.items():
        if only and name not in only:
            continue
        rng = np.random.default_rng(abs(hash("ode" + name)) % (2 ** 32) if False else sum(map(ord, "ode" + name)))
        rec = build_case(name, c, rng)
        np.savez_compressed(os.path.join(HERE, f"ode_{name}.npz"), **rec)
        shapes = {k[6:]: v.shape for k, v in rec.items() if k.startswith("param:")}
        print(name, "dp", rec["dp"].shape, "da", rec["da"].shape, "|dp|", float(np.abs(rec["dp"]).max()), "|da|", float(np.abs(rec["da"]).max()))
        if name == "ponita":
            for k, s in shapes.items():
                print("   ", k, s)
