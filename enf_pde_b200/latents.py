"""Latent pytree (p, a, gaussian_window) initialisation -- the layout kept from enf/latents.

Restates the host-side initialisers of enf/latents/utils.py:4-109 and the window-size heuristic of
enf/latents/autodecoder.py:38-56 (numpy, no device arithmetic); used by bench.py and examples to
build realistic synthetic latents.  Returns float32 torch tensors shaped (S, Z, P), (S, Z, L), (S, Z, 1).
"""
import numpy as np
import torch


def init_positions_grid(num_signals, num_latents, num_dims):          # utils.py:73-103
    npd = int(round(num_latents ** (1.0 / num_dims)))
    assert npd ** num_dims == num_latents, "num_latents must be a power of the number of position dimensions"
    ax = np.linspace(-1 + 1 / npd, 1 - 1 / npd, npd)
    g = np.stack(np.meshgrid(*[ax] * num_dims, indexing="ij"), axis=-1).reshape(-1, num_dims)
    return np.repeat(g[None], num_signals, axis=0)


def init_positions_polar(num_signals, n_phi, n_theta):                # utils.py:36-70 (n_phi x n_theta grid)
    phi = np.linspace(np.pi / n_phi, 2 * np.pi - np.pi / n_phi, n_phi)
    theta = np.linspace((np.pi / 2) / n_theta, np.pi - (np.pi / 2) / n_theta, n_theta)
    g = np.stack(np.meshgrid(phi, theta, indexing="ij"), axis=-1).reshape(-1, 2)
    return np.repeat(g[None], num_signals, axis=0)


def init_positions_ball(num_signals, num_latents):                    # utils.py:4-33
    idx = np.arange(1, num_latents + 1)
    alpha = np.arccos(1 - 2 * idx / (num_latents + 1))
    beta = np.pi * (1 + 5 ** 0.5) * idx
    gamma = np.arange(num_latents) * (2 * np.pi / num_latents)
    pos = np.stack([alpha, beta, gamma, np.full(num_latents, 0.75)], axis=-1)
    return np.repeat(pos[None], num_signals, axis=0)


def init_latents(invariant, num_signals, num_latents, latent_dim, polar_grid=None, device="cpu"):
    """(p, a, gaussian_window) as PositionOrientationFeatureAutodecoder initialises them (autodecoder.py:21-56)."""
    t = invariant.invariant_type
    if t in ("polar_periodic", "latitude_periodic"):
        if polar_grid is None:
            n_theta = int(round((num_latents // 2) ** 0.5))
            polar_grid = (2 * n_theta, n_theta)
        assert polar_grid[0] * polar_grid[1] == num_latents
        p = init_positions_polar(num_signals, *polar_grid)
        sigma0 = 2 * np.pi / polar_grid[1]
    elif t in ("ball", "ball_lat"):
        p = init_positions_ball(num_signals, num_latents)
        sigma0 = 1.0
    else:
        npos = invariant.num_z_pos_dims
        p = init_positions_grid(num_signals, num_latents, npos)
        sigma0 = npos / int(round(num_latents ** (1.0 / npos)))
        if invariant.num_z_ori_dims > 0:                               # utils.py:106-109
            p = np.concatenate([p, np.arctan2(p[:, :, 0], p[:, :, 1])[:, :, None]], axis=-1)
    a = np.ones((num_signals, num_latents, latent_dim))
    sigma = np.full((num_signals, num_latents, 1), sigma0)
    f = lambda v: torch.tensor(v, dtype=torch.float32, device=device)
    return f(p), f(a), f(sigma)


def make_coords(invariant, grid):
    """Coordinate grids as the reference's fit_*.py scripts build them (fit_navier_stokes.py:32-33,
    fit_ihc.py:33-37, datasets/pdes.py:469-488).  Returns a float32 (C, Dx) tensor shared by all fields."""
    t = invariant.invariant_type
    if t in ("polar_periodic", "latitude_periodic"):
        nphi, nth = grid
        phi = np.linspace(0, 2 * np.pi, nphi, endpoint=False)
        th = np.linspace(0, np.pi, nth + 2)[1:-1]
        g = np.stack(np.meshgrid(phi, th, indexing="ij"), axis=-1).reshape(-1, 2)
    elif t in ("ball", "ball_lat"):
        nphi, nth, nr = grid
        phi = np.linspace(0, 2 * np.pi, nphi, endpoint=False)
        th = np.linspace(1e-3, np.pi, nth, endpoint=False)
        r = np.linspace(0, 1, nr)
        g = np.stack(np.meshgrid(phi, th, r, indexing="ij"), axis=-1).reshape(-1, 3)
    else:
        axes = [np.linspace(-1, 1, n) for n in grid]
        g = np.stack(np.meshgrid(*axes), axis=-1).reshape(-1, len(grid))
    return torch.tensor(g, dtype=torch.float32)
