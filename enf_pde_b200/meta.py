"""Meta-SGD inner loop over the accelerated NeF (first-order part).

Mirrors `PDETrainer.inner_loop` (experiments/fitting/trainers/pde_trainer.py:122-235): starting from the shared
autodecoder latents, `num_inner_steps` SGD steps on each field's own latents (p, a, gaussian_window) against a
different random subset of the field's samples per step, per-key meta learning rates
(`meta_sgd_lrs`: p_pos (1,), p_ori (1,), a (L,), gaussian_window (1,), pde_trainer.py:83-97), gradients scaled by
the batch size because the loss is a mean over fields (pde_trainer.py:207), window updates zeroed unless
`optimize_gaussian_window` (pde_trainer.py:210-212); then the reconstruction loss of the adapted latents on one
more subset.

Every step is `nef.apply` + the latents-only backward of the C ABI (`dW = NULL`: no weight-gradient work) on the
caller's stream; nothing synchronises with the host.  `inner_loop` is FIRST ORDER: the latents returned are exact, and so
is the loss value, but the loss is differentiable w.r.t. the NeF parameters only through the last `apply` (FOMAML); it
raises if the initial latents or the learning rates ask for a gradient.  Exact uses: test-time adaptation in
`validate_epoch` / `visualize_batch`, the latents fed to the ODE phase, the non-MAML trainer's latent fitting.
The reference's training step differentiates through the whole loop (second order, `jax.value_and_grad` at
pde_trainer.py:255): that is `outer_step_gradients` below.
"""
from typing import Dict, Optional, Sequence, Tuple

import torch


def _lr_vector(inv, lrs: Dict[str, torch.Tensor], pose_dim: int, device) -> torch.Tensor:
    """per-component learning rate of the raw pose (positions, then orientation angles)"""
    n_pos = inv.num_z_pos_dims
    lr = torch.empty(pose_dim, device=device, dtype=torch.float32)
    lr[:n_pos] = lrs["p_pos"].to(device).reshape(-1)[0]
    if pose_dim > n_pos:
        lr[n_pos:] = lrs["p_ori"].to(device).reshape(-1)[0]
    return lr


def inner_loop(nef, variables, coords: torch.Tensor, img: torch.Tensor, p: torch.Tensor, a: torch.Tensor,
               gaussian_window: Optional[torch.Tensor], meta_sgd_lrs: Dict[str, torch.Tensor], num_inner_steps: int,
               masks: Optional[Sequence[torch.Tensor]] = None, optimize_gaussian_window: bool = False,
               generator: Optional[torch.Generator] = None, max_num_sampled_points: Optional[int] = None
               ) -> Tuple[torch.Tensor, Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]]:
    """coords (C, Dx) shared grid; img (B, C, O); p (B, Z, P), a (B, Z, L), gaussian_window (B, Z, 1): the broadcast
    autodecoder latents (inner_autodecoder_params, pde_trainer.py:155-158).  `masks`: num_inner_steps + 1 index
    tensors into C (pde_trainer.py:146-152); drawn from `generator` when None.
    Returns (loss of the adapted latents on the last mask, (p, a, gaussian_window) adapted)."""
    inv = nef.cross_attn_invariant
    B, C = img.shape[0], coords.shape[0]
    dev = p.device
    wants_grad = [n for n, t in (("p", p), ("a", a), ("gaussian_window", gaussian_window), *meta_sgd_lrs.items())
                  if t is not None and getattr(t, "requires_grad", False)]
    if wants_grad:
        # the reference trains the shared latents and the Meta-SGD rates by differentiating through this loop; this routine
        # detaches every step, so it must not hand back silent `None` gradients for them
        raise RuntimeError(f"inner_loop is first order: it cannot give gradients w.r.t. {wants_grad}; use "
                           "enf_pde_b200.outer_step_gradients (second order through the loop) for the outer step")
    if masks is None:
        M = C if max_num_sampled_points is None else min(C, max_num_sampled_points)
        masks = [torch.randperm(C, generator=generator)[:M].to(dev) for _ in range(num_inner_steps + 1)]
    if len(masks) != num_inner_steps + 1:
        raise ValueError("need num_inner_steps + 1 masks")
    lr_p = _lr_vector(inv, meta_sgd_lrs, p.shape[-1], dev)
    lr_a = meta_sgd_lrs["a"].to(dev).reshape(1, 1, -1)
    lr_w = meta_sgd_lrs["gaussian_window"].to(dev).reshape(-1)[0] if gaussian_window is not None else None
    frozen = {"params": _detach_tree(variables["params"])}          # inner steps: latents-only backward (dW = NULL)
    p, a = p.detach().clone(), a.detach().clone()
    w = None if gaussian_window is None else gaussian_window.detach().clone()
    for step in range(num_inner_steps):
        m = masks[step]
        xs, ys = coords[m].contiguous(), img[:, m].contiguous()
        p.requires_grad_(True); a.requires_grad_(True)
        if w is not None:
            w.requires_grad_(True)
        out = nef.apply(frozen, xs, p, a, w)
        # d(mean((out - y)^2))/d(out), times B (the loss is a mean over fields, pde_trainer.py:207)
        d_out = (out.detach() - ys) * (2.0 * B / out.numel())
        out.backward(d_out)
        with torch.no_grad():
            p_new = p - lr_p * p.grad
            a_new = a - lr_a * a.grad
            w_new = None if w is None else (w - lr_w * w.grad if optimize_gaussian_window else w.detach())
        p, a = p_new.detach(), a_new.detach()
        w = None if w_new is None else w_new.detach()
    m = masks[num_inner_steps]
    out = nef.apply(variables, coords[m].contiguous(), p, a, w)
    loss = ((out - img[:, m]) ** 2).mean()
    return loss, (p, a, w)


def _detach_tree(tree):
    if isinstance(tree, dict):
        return {k: _detach_tree(v) for k, v in tree.items()}
    return tree.detach()


# ---------------------------------------------------------------------------------------------------------------------------
# second order: the gradient of the OUTER objective through the inner loop
# ---------------------------------------------------------------------------------------------------------------------------

def _fd_weights(order):
    # central differences of the gradient along a direction: sum_i c_i G(z + s_i t v) / t
    return ((1.0, 0.5), (-1.0, -0.5)) if order == 2 else ((2.0, -1.0 / 12), (1.0, 8.0 / 12), (-1.0, -8.0 / 12), (-2.0, 1.0 / 12))


def outer_step_gradients(nef, variables, coords: torch.Tensor, img: torch.Tensor, p0: torch.Tensor, a0: torch.Tensor,
                         gaussian_window0: Optional[torch.Tensor], meta_sgd_lrs: Dict[str, torch.Tensor], num_inner_steps: int,
                         masks: Sequence[torch.Tensor], optimize_gaussian_window: bool = False, pos_noise: Optional[torch.Tensor] = None,
                         hvp_eps: float = 1e-2, hvp_order: int = 4):
    """Loss and gradient of the meta-learning OUTER step, second order included: what
    `jax.value_and_grad(self.enf_loss)` (pde_trainer.py:255) returns for `params = {nef, autodecoder, meta_sgd_lrs}` when it
    differentiates through `inner_loop` (:122-235) and its un-stopped `jax.grad` steps (:188-222).

    p0 (1,Z,P), a0 (1,Z,L), gaussian_window0 (1,Z,1): the shared autodecoder latents (repeated over the B fields, :157-159;
    `pos_noise` (B,Z,n_pos) is the optional position noise of :162-167).  Returns
        loss, {"nef": [46 leaf gradients, EnfWeights order], "p": (1,Z,P), "a": (1,Z,L), "gaussian_window": (1,Z,1),
               "lrs": {key: gradient shaped like meta_sgd_lrs[key]}}, (p_K, a_K, window_K) adapted latents.

    How: the inner loop runs forward on the accelerated path (latents-only backward, as `inner_loop`), keeping every step's
    latents z_k and scaled gradient g_k = B dl_k/dz.  The reverse sweep is the adjoint of z_{k+1} = z_k - lr * g_k:
        lambda_K = dL/dz_K, G_theta = dL/dtheta                        (one ordinary backward with weight gradients)
        for k = K-1 .. 0:   u = lr * lambda_{k+1}
                            G_lr     -= sum(lambda_{k+1} * g_k)  per key
                            [h_theta, h_z] = B d/dt grad_{theta,z} l_k(theta, z_k + t u) at t = 0       (Hessian-vector product)
                            G_theta  -= h_theta ;  lambda_k = lambda_{k+1} - h_z
        dL/d(p0, a0, window0) = sum over fields of lambda_0.
    The Hessian-vector product is a central finite difference (order 2 or 4) of `enf_xattn_bwd` gradients, ALWAYS on the fp32
    kernels (a difference of 16-bit-operand gradients is noise), with the relu activation pattern frozen at z_k
    (ENF_FLAG_FROZEN_RELU): the difference then differentiates one linear branch of every relu, which is what
    reverse-over-reverse autodiff computes (relu'' = 0) -- a plain finite difference would add the curvature concentrated at the
    kinks.  Every other operation of the path is smooth.  Cost: `hvp_order` extra forward + backward passes per inner step, on
    the per-step query subsets."""
    import copy
    from . import _lib
    from .nef import params_to_leaves
    inv = nef.cross_attn_invariant
    B, dev = img.shape[0], p0.device
    K = num_inner_steps
    if len(masks) != K + 1:
        raise ValueError("need num_inner_steps + 1 masks")
    use_w = gaussian_window0 is not None
    n_pos = inv.num_z_pos_dims
    lr_p = _lr_vector(inv, meta_sgd_lrs, p0.shape[-1], dev)
    lr_a = meta_sgd_lrs["a"].to(dev).reshape(1, 1, -1).float()
    lr_w = meta_sgd_lrs["gaussian_window"].to(dev).reshape(-1)[0].float() if use_w else None
    leaves = [t.detach() for t in params_to_leaves(variables)]
    frozen = {"params": _detach_tree(variables["params"])}
    nef32 = copy.copy(nef)                                   # fp32 twin for the Hessian-vector products
    nef32.precision, nef32.recompute, nef32.out_bf16 = _lib.PREC_FP32, False, False

    def grads_at(step, p, a, w, with_weights, module, mask_pose=None):
        """B * d l_step / d(theta, p, a, w) at the given latents (l = mean squared error over the step's subset)."""
        m = masks[step]
        xs, ys = coords[m].contiguous(), img[:, m].contiguous()
        p = p.detach().requires_grad_(True); a = a.detach().requires_grad_(True)
        w = None if w is None else w.detach().requires_grad_(True)
        if with_weights:
            P = [t.detach().requires_grad_(True) for t in leaves]
            from .nef import leaves_to_params
            var = leaves_to_params(P)
        else:
            P, var = None, frozen
        out = module.apply(var, xs, p, a, w, relu_mask_pose=mask_pose) if mask_pose is not None else module.apply(var, xs, p, a, w)
        diff = out.detach() - ys
        out.backward(diff * (2.0 * B / out.numel()))
        loss = (diff * diff).mean()
        gw = None if w is None else w.grad
        return loss, ([t.grad for t in P] if with_weights else None), p.grad, a.grad, gw

    # ---- forward sweep: the inner loop (pde_trainer.py:191-222) ----------------------------------------------------------
    p = p0.detach().repeat(B, 1, 1).contiguous()
    if pos_noise is not None:
        p[..., :n_pos] += pos_noise
    a = a0.detach().repeat(B, 1, 1).contiguous()
    w = gaussian_window0.detach().repeat(B, 1, 1).contiguous() if use_w else None
    traj = []
    for k in range(K):
        _, _, gp, ga, gw = grads_at(k, p, a, w, False, nef)
        if use_w and not optimize_gaussian_window:
            gw = torch.zeros_like(gw)                        # pde_trainer.py:210-212
        traj.append((p, a, w, gp, ga, gw))
        p = p - lr_p * gp
        a = a - lr_a * ga
        if use_w and optimize_gaussian_window:
            w = w - lr_w * gw
    # ---- outer loss and its direct gradient -------------------------------------------------------------------------------
    loss, g_theta, lam_p, lam_a, lam_w = grads_at(K, p, a, w, True, nef)
    inv_B = 1.0 / B                                          # grads_at scales by B (inner-loop convention); the outer loss does not
    g_theta = [g * inv_B for g in g_theta]
    lam_p, lam_a = lam_p * inv_B, lam_a * inv_B
    lam_w = lam_w * inv_B if use_w else None
    adapted = (p, a, w)
    g_lr = {k: torch.zeros_like(v, dtype=torch.float32, device=dev) for k, v in meta_sgd_lrs.items()}
    # ---- reverse sweep ---------------------------------------------------------------------------------------------------
    for k in range(K - 1, -1, -1):
        pk, ak, wk, gp, ga, gw = traj[k]
        g_lr["p_pos"] -= (lam_p[..., :n_pos] * gp[..., :n_pos]).sum().reshape(g_lr["p_pos"].shape)
        if "p_ori" in g_lr and pk.shape[-1] > n_pos:
            g_lr["p_ori"] -= (lam_p[..., n_pos:] * gp[..., n_pos:]).sum().reshape(g_lr["p_ori"].shape)
        g_lr["a"] -= (lam_a * ga).sum((0, 1)).reshape(g_lr["a"].shape)
        if use_w and optimize_gaussian_window:
            g_lr["gaussian_window"] -= (lam_w * gw).sum().reshape(g_lr["gaussian_window"].shape)
        up, ua = lr_p * lam_p, lr_a * lam_a
        uw = lr_w * lam_w if (use_w and optimize_gaussian_window) else None
        vmax = max(float(up.abs().max()), float(ua.abs().max()), float(uw.abs().max()) if uw is not None else 0.0)
        if vmax == 0.0:
            continue
        t = hvp_eps / vmax
        h_theta = [torch.zeros_like(g) for g in g_theta]
        h_p, h_a = torch.zeros_like(lam_p), torch.zeros_like(lam_a)
        h_w = torch.zeros_like(lam_w) if use_w else None
        for s, c in _fd_weights(hvp_order):
            we = None if wk is None else (wk + s * t * uw if uw is not None else wk)
            _, gth, gpp, gaa, gww = grads_at(k, pk + s * t * up, ak + s * t * ua, we, True, nef32, mask_pose=pk)
            for h, g in zip(h_theta, gth):
                h += (c / t) * g
            h_p += (c / t) * gpp
            h_a += (c / t) * gaa
            if use_w:
                h_w += (c / t) * gww
        g_theta = [g - h for g, h in zip(g_theta, h_theta)]
        lam_p, lam_a = lam_p - h_p, lam_a - h_a
        if use_w:
            # the window's own update is zeroed unless optimize_gaussian_window, but the loss still depends on it directly
            lam_w = lam_w - h_w
    grads = {"nef": g_theta, "p": lam_p.sum(0, keepdim=True), "a": lam_a.sum(0, keepdim=True),
             "gaussian_window": lam_w.sum(0, keepdim=True) if use_w else None, "lrs": g_lr}
    return loss, grads, adapted
