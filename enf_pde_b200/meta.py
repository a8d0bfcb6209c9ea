"""Meta-SGD inner loop over the accelerated NeF (first-order part).

Mirrors `PDETrainer.inner_loop` (experiments/fitting/trainers/pde_trainer.py:122-235): starting from the shared
autodecoder latents, `num_inner_steps` SGD steps on each field's own latents (p, a, gaussian_window) against a
different random subset of the field's samples per step, per-key meta learning rates
(`meta_sgd_lrs`: p_pos (1,), p_ori (1,), a (L,), gaussian_window (1,), pde_trainer.py:83-97), gradients scaled by
the batch size because the loss is a mean over fields (pde_trainer.py:207), window updates zeroed unless
`optimize_gaussian_window` (pde_trainer.py:210-212); then the reconstruction loss of the adapted latents on one
more subset.

Every step is `nef.apply` + the latents-only backward of the C ABI (`dW = NULL`: no weight-gradient work) on the
caller's stream; nothing synchronises with the host.  FIRST ORDER: the latents returned are exact, and so is the
loss value, but the loss is differentiable w.r.t. the NeF parameters only through the last `apply` (FOMAML) -- the
reference differentiates through the whole loop (second order, `jax.value_and_grad` at pde_trainer.py:255), which
needs the double-backward kernels listed as "next" in DESIGN.md.  Exact uses: test-time adaptation in
`validate_epoch` / `visualize_batch`, the latents fed to the ODE phase, the non-MAML trainer's latent fitting.
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
