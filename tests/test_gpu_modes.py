"""Modes of the C ABI beyond the plain fwd + bwd: repeated backward on one forward, bounded-memory (recompute) training,
bfloat16 forward-only output.  `-m gpu`."""
import ctypes

import pytest
import torch

from oracle import enf_ref as R
from helpers import rel_err, make_case, worst_leaf, Checker, compare, TOL_FP32, TOL_TC, TOL_TC_LEAF, TOL_TC_SMALL

pytestmark = pytest.mark.gpu


def _nef(cfg, precision, **kw):
    import types
    import enf_pde_b200 as E
    inv = E.get_ca_invariant(types.SimpleNamespace(invariant_type=cfg.invariant_type, num_in=cfg.num_in))
    return E.EquivariantCrossAttentionNeF(cfg.num_hidden, cfg.num_heads, 0, cfg.num_out, cfg.latent_dim, inv, inv, "rff",
                                          cfg.embedding_freq_multiplier, True, cfg.use_gaussian_window, precision=precision, **kw)


def _cuda(params):
    return R.tree_map(lambda t: t.to("cuda", torch.float32).contiguous().requires_grad_(True), params)


f32 = lambda t: t.to("cuda", torch.float32)


@pytest.mark.parametrize("precision,hidden", [("fp32", 64), ("bf16", 128), ("bf16", 64)])
def test_backward_is_repeatable(precision, hidden):
    """enf_xattn_bwd does not consume the forward state: two backward calls on ONE forward (PyTorch retain_graph=True, a JAX
    vjp applied to two cotangents) give each cotangent's own gradients -- checked against the oracle for the second one
    (round 1 reused the forward's W3 buffer as scratch, which made a second backward silently wrong)."""
    cfg = R.EnfConfig(num_in=2, num_hidden=hidden, num_heads=2, num_out=1, latent_dim=16, invariant_type="ponita",
                      embedding_freq_multiplier=(0.05, 0.05))
    params, x, p, a, sigma, d1 = make_case(cfg, 2, 200, 9, seed=21)
    d2 = (torch.randn(d1.shape, generator=torch.Generator().manual_seed(5), dtype=torch.float64) / d1.numel()).float().double()
    chk = Checker(cfg, (params, x, p, a, sigma, d2))
    nef = _nef(cfg, precision)
    P = _cuda(params)
    pg, ag, sg = (f32(t).requires_grad_(True) for t in (p, a, sigma))
    out = nef.apply(P, f32(x), pg, ag, sg)
    out.backward(f32(d1), retain_graph=True)
    leaves = R.tree_flatten(P["params"])
    for t in (pg, ag, sg, *leaves.values()):
        t.grad = None
    out.backward(f32(d2))
    tol, tol_leaf = (TOL_FP32, TOL_FP32) if precision == "fp32" else (TOL_TC if hidden == 128 else TOL_TC_SMALL, TOL_TC_LEAF)
    errs, worst, ok = compare(chk, out.detach(), pg.grad, ag.grad, sg.grad, {k: v.grad for k, v in leaves.items()}, tol, tol_leaf)
    print(precision, hidden, {k: f"{v:.2e}" for k, v in errs.items()}, worst, chk.used_allowance)
    assert ok, (errs, worst)


def test_backward_rejects_a_mismatched_call():
    """the backward must be given the forward's description AND x_batch_stride (the stride was not checked)."""
    from enf_pde_b200 import _lib
    from gpu_helpers import desc_for
    from enf_pde_b200.nef import _weights_struct, params_to_leaves
    lib = _lib.load()
    cfg = R.EnfConfig(num_in=2, num_hidden=32, num_heads=2, num_out=1, latent_dim=8, invariant_type="rel_pos_periodic")
    params, x, p, a, sigma, d_out = make_case(cfg, 2, 40, 4, seed=2)
    desc = desc_for(cfg, 2, 40, 4)
    leaves = [f32(t).contiguous() for t in params_to_leaves(params)]
    w = _weights_struct(leaves)
    xg, pg, ag, sg, dg = (f32(t).contiguous() for t in (x, p, a, sigma, d_out))
    n = lib.enf_xattn_workspace_bytes(ctypes.byref(desc))
    ws = torch.empty(n + 256, dtype=torch.uint8, device="cuda")
    out = torch.empty(2, 40, 1, device="cuda")
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    assert lib.enf_xattn_fwd(ctypes.byref(desc), ctypes.byref(w), ptr(xg), 80, ptr(pg), ptr(ag), ptr(sg), ptr(out), ptr(ws), n, None) == 0
    dp, da, ds = torch.empty_like(pg), torch.empty_like(ag), torch.empty_like(sg)
    bwd = lambda xbs, wsp: lib.enf_xattn_bwd(ctypes.byref(desc), ctypes.byref(w), ptr(xg), xbs, ptr(pg), ptr(ag), ptr(sg), ptr(dg),
                                             None, ptr(dp), ptr(da), ptr(ds), wsp, n, None)
    assert bwd(0, ptr(ws)) == -7                       # forward ran with per-field coordinates, backward claims a shared grid
    assert bwd(7, ptr(ws)) == -1                       # not a legal stride at all
    assert bwd(80, ctypes.c_void_p(ws.data_ptr() + 16)) == -4      # misaligned workspace
    assert bwd(80, ptr(ws)) == 0
    lib.enf_workspace_release(ptr(ws))
    assert bwd(80, ptr(ws)) == -7                      # state released
    torch.cuda.synchronize()


@pytest.mark.parametrize("chunk", [1, 2, 0])
def test_recompute_mode_matches_stash_mode(chunk):
    """ENF_FLAG_RECOMPUTE: no per-(query, latent) stash; the backward re-runs the pair forward per chunk of fields.  Same
    kernels on the same inputs, so the gradients agree with the default mode to summation-order noise, with a smaller workspace."""
    from enf_pde_b200.nef import _XAttnFunction
    cfg = R.EnfConfig(num_in=2, num_hidden=128, num_heads=2, num_out=1, latent_dim=16, invariant_type="rel_pos_periodic",
                      embedding_freq_multiplier=(0.05, 0.1))
    params, x, p, a, sigma, d_out = make_case(cfg, 5, 300, 16, seed=4)
    res = []
    for kw in (dict(), dict(recompute=True, chunk_fields=chunk)):
        nef = _nef(cfg, "bf16", **kw)
        P = _cuda(params)
        pg, ag, sg = (f32(t).requires_grad_(True) for t in (p, a, sigma))
        out = nef.apply(P, f32(x), pg, ag, sg)
        nbytes = _XAttnFunction.last_ws[2]
        out.backward(f32(d_out))
        res.append((out.detach(), pg.grad, ag.grad, sg.grad, {k: v.grad for k, v in R.tree_flatten(P["params"]).items()}, nbytes))
    (o0, dp0, da0, ds0, g0, n0), (o1, dp1, da1, ds1, g1, n1) = res
    assert torch.equal(o0, o1)
    assert rel_err(dp1, dp0) < 2e-5 and rel_err(da1, da0) < 2e-5 and rel_err(ds1, ds0) < 2e-5
    assert worst_leaf(g1, g0)[0] < 2e-4        # fp32 atomics in a different order (per-leaf scale, small leaves)
    assert n1 < n0, (n0, n1)


def test_recompute_mode_under_a_workspace_cap_against_oracle():
    """workspace_cap_bytes -> enf_xattn_chunk_for_cap picks the chunk; result still inside the 2e-3 bucket of the oracle."""
    from enf_pde_b200.nef import _XAttnFunction
    cfg = R.EnfConfig(num_in=2, num_hidden=128, num_heads=2, num_out=3, latent_dim=32, invariant_type="latitude_periodic",
                      embedding_freq_multiplier=(0.05, 0.2))
    params, x, p, a, sigma, d_out = make_case(cfg, 3, 900, 18, seed=5, polar_grid=(6, 3))
    chk = Checker(cfg, (params, x, p, a, sigma, d_out))
    full = _nef(cfg, "bf16")
    P = _cuda(params)
    full.apply(P, f32(x), f32(p).requires_grad_(True), f32(a), f32(sigma))
    cap = int(0.8 * _XAttnFunction.last_ws[2])
    nef = _nef(cfg, "bf16", workspace_cap_bytes=cap)
    pg, ag, sg = (f32(t).requires_grad_(True) for t in (p, a, sigma))
    out = nef.apply(P, f32(x), pg, ag, sg)
    assert _XAttnFunction.last_ws[2] <= cap and 1 <= _XAttnFunction.last_ws[0]["chunk_fields"] < 3
    out.backward(f32(d_out))
    errs, worst, ok = compare(chk, out.detach(), pg.grad, ag.grad, sg.grad, {k: v.grad for k, v in R.tree_flatten(P["params"]).items()},
                              TOL_TC, TOL_TC_LEAF)
    print({k: f"{v:.2e}" for k, v in errs.items()}, worst, chk.used_allowance)
    assert ok, (errs, worst)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_bf16_output_option_forward_only(precision):
    """ENF_FLAG_OUT_BF16 (validation roll-outs): same forward, decoded field rounded once to bfloat16."""
    cfg = R.EnfConfig(num_in=2, num_hidden=128, num_heads=2, num_out=3, latent_dim=16, invariant_type="rel_pos_periodic",
                      embedding_freq_multiplier=(0.05, 0.1))
    params, x, p, a, sigma, _ = make_case(cfg, 2, 200, 16, seed=9)
    P = _cuda(params)
    with torch.no_grad():
        o32 = _nef(cfg, precision).apply(P, f32(x), f32(p), f32(a), f32(sigma))
        o16 = _nef(cfg, precision, out_bf16=True).apply(P, f32(x), f32(p), f32(a), f32(sigma))
    assert o16.dtype == torch.bfloat16 and o32.dtype == torch.float32
    assert torch.equal(o16, o32.to(torch.bfloat16))


@pytest.mark.parametrize("optimize_window", [False, True])
def test_second_order_outer_gradient_against_oracle(optimize_window):
    """enf_pde_b200.outer_step_gradients vs the oracle's double backward through the restated inner loop
    (jax.value_and_grad at pde_trainer.py:255 through jax.grad at :188-204), K = 3 Meta-SGD steps: gradients w.r.t. the 46 NeF
    leaves, the shared autodecoder latents (p, a, gaussian_window) and the Meta-SGD learning rates.
    The Hessian-vector products are 4th-order central differences of the fp32 backward with the relu pattern frozen
    (ENF_FLAG_FROZEN_RELU); tolerance 1e-4 -- the fp32 bucket -- on every quantity (per leaf for the weights; measured 3e-6), and
    the test also shows that the
    first-order (FOMAML) estimate misses by far more than that, i.e. that the second-order terms are being checked."""
    import enf_pde_b200 as E
    from helpers import leaf_errs
    cfg = R.EnfConfig(num_in=2, num_hidden=64, num_heads=2, num_out=1, latent_dim=16, invariant_type="ponita",
                      embedding_freq_multiplier=(0.05, 0.05))
    B, C, Z, K = 3, 256, 9, 3
    params, _, p, a, sigma, _ = make_case(cfg, 1, 8, Z, seed=13)
    coords = R.make_coords(cfg, (16, 16)).float().double()
    g = torch.Generator().manual_seed(17)
    img = torch.randn(B, C, cfg.num_out, generator=g, dtype=torch.float64).float().double()
    masks = [torch.randperm(C, generator=g)[:180] for _ in range(K + 1)]
    lrs = {"p_pos": torch.tensor([0.3]), "p_ori": torch.tensor([0.2]), "a": torch.full((cfg.latent_dim,), 1.5),
           "gaussian_window": torch.tensor([0.05])}
    loss_ref, g_ref = R.outer_loss_and_grads(cfg, params, coords, img, p, a, sigma, {k: v.double() for k, v in lrs.items()}, K, masks,
                                             optimize_gaussian_window=optimize_window, n_pos=2)
    nef = _nef(cfg, "fp32")
    P = _cuda(params)
    loss, grads, _ = E.outer_step_gradients(nef, P, f32(coords), f32(img), f32(p), f32(a), f32(sigma),
                                            {k: v.cuda() for k, v in lrs.items()}, K, [m.cuda() for m in masks],
                                            optimize_gaussian_window=optimize_window)
    from enf_pde_b200 import _lib
    want = R.tree_flatten(g_ref["nef"]["params"])
    got = {_lib.LEAF_PATHS[n]: t for n, t in zip(_lib.LEAVES, grads["nef"])}
    le = leaf_errs(got, want)
    errs = dict(loss=abs(float(loss) - float(loss_ref)) / abs(float(loss_ref)), p=rel_err(grads["p"], g_ref["p"]),
                a=rel_err(grads["a"], g_ref["a"]), window=rel_err(grads["gaussian_window"], g_ref["gaussian_window"]),
                dtheta=max(le.values()))
    for k in lrs:
        if float(g_ref["lrs"][k].abs().max()) > 0:
            errs["lr_" + k] = rel_err(grads["lrs"][k], g_ref["lrs"][k])
        else:
            assert float(grads["lrs"][k].abs().max()) == 0.0, k
    print("outer gradient", optimize_window, {k: f"{v:.2e}" for k, v in errs.items()}, "worst leaf:", max(le, key=le.get))
    assert all(v < TOL_FP32 for v in errs.values()), errs
    # the second-order terms matter on this problem: the first-order (FOMAML) estimate -- the last apply's own weight gradient
    # at the adapted latents -- misses the oracle's outer gradient by far more than the tolerance above
    _, (pK, aK, sK) = R.inner_loop(cfg, params, coords, img, p.repeat(B, 1, 1), a.repeat(B, 1, 1), sigma.repeat(B, 1, 1), lrs, K, masks,
                                   optimize_gaussian_window=optimize_window, n_pos=2)
    xs = coords[masks[K]][None].expand(B, -1, -1)
    out = R.nef_apply(cfg, params, xs, pK, aK, sK)
    _, g_fo, _, _, _ = R.fwd_bwd(cfg, params, xs, pK, aK, sK, 2 * (out - img[:, masks[K]]) / out.numel())
    fo_err = max(leaf_errs(R.tree_flatten(g_fo["params"]), want).values())
    print("first-order estimate misses by", f"{fo_err:.2e}")
    assert fo_err > 2e-2


def test_forward_only_field_chunks_are_bit_identical():
    """forward_chunk_fields: validation roll-outs decode B*T fields in chunks (workspace of one chunk); same bits as one call."""
    cfg = R.EnfConfig(num_in=2, num_hidden=128, num_heads=2, num_out=1, latent_dim=16, invariant_type="rel_pos_periodic",
                      embedding_freq_multiplier=(0.05, 0.1))
    params, x, p, a, sigma, _ = make_case(cfg, 7, 300, 16, seed=31)
    P = R.tree_map(lambda t: t.to("cuda", torch.float32).contiguous(), params)
    with torch.no_grad():
        whole = _nef(cfg, "bf16").apply(P, f32(x), f32(p), f32(a), f32(sigma))
        from enf_pde_b200.nef import _XAttnFunction
        ws_whole = _XAttnFunction.last_ws[2]
        parts = _nef(cfg, "bf16", forward_chunk_fields=3).apply(P, f32(x), f32(p), f32(a), f32(sigma))      # 3 + 3 + 1 fields
        ws_part = _XAttnFunction.last_ws[2]
        shared = _nef(cfg, "bf16", forward_chunk_fields=2).apply(P, f32(x[:1]).expand(7, -1, -1), f32(p), f32(a), f32(sigma))
        shared_ref = _nef(cfg, "bf16").apply(P, f32(x[:1]).expand(7, -1, -1), f32(p), f32(a), f32(sigma))
    assert torch.equal(whole, parts) and torch.equal(shared, shared_ref)
    assert ws_part < ws_whole


@pytest.mark.parametrize("hidden,heads", [(128, 2), (64, 2), (32, 3)])
def test_latents_only_backward_matches_full_backward(hidden, heads):
    """dW = NULL (Meta-SGD inner steps, ODE phase): the tcgen05 backward kernels B / C skip the shared-weight gradient MMAs and their
    flush; the latent gradients are the full backward's (up to the order of the float atomics that reduce them over tiles)."""
    inv = "ball" if hidden == 32 else "rel_pos_periodic"
    cfg = R.EnfConfig(num_in=3 if inv == "ball" else 2, num_hidden=hidden, num_heads=heads, num_out=1, latent_dim=16, invariant_type=inv,
                      embedding_freq_multiplier=(0.05, 0.1))
    params, x, p, a, sigma, d_out = make_case(cfg, 2, 300, 16, seed=41)
    nef = _nef(cfg, "bf16")
    grads = []
    for with_weights in (True, False):
        P = R.tree_map(lambda t: t.to("cuda", torch.float32).contiguous().requires_grad_(with_weights), params)
        pg, ag, sg = (f32(t).requires_grad_(True) for t in (p, a, sigma))
        nef.apply(P, f32(x), pg, ag, sg).backward(f32(d_out))
        grads.append((pg.grad, ag.grad, sg.grad))
    for full, lat in zip(*grads):
        assert rel_err(lat, full) < 1e-5
