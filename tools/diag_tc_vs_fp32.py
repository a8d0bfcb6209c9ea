"""Diagnostics (GPU): tensor-core path vs fp32 path of the library on one problem at FULL size, stage buffer by stage buffer.

    python tools/diag_tc_vs_fp32.py ihc [nrows]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import enf_ref as R                      # noqa: E402
from helpers import make_case, rel_err               # noqa: E402
import gpu_helpers                                   # noqa: E402
from test_gpu_real_shapes import REAL                # noqa: E402
from enf_pde_b200 import _lib                        # noqa: E402
from enf_pde_b200.nef import _weights_struct, params_to_leaves   # noqa: E402

name = sys.argv[1]
nrows = int(sys.argv[2]) if len(sys.argv) > 2 else 256
kw, B, grid, Z, polar, _, _ = REAL[name]
cfg = R.EnfConfig(**kw)
C = int(np.prod(grid))
params, _, p, a, sigma, _ = make_case(cfg, B, 4, Z, seed=31, polar_grid=polar)
coords = R.make_coords(cfg, grid).float()
g = torch.Generator().manual_seed(7)
rows = torch.randperm(C, generator=g)[:nrows]
d_out = torch.zeros(B, C, cfg.num_out)
d_out[:, rows] = torch.randn(B, nrows, cfg.num_out, generator=g) / (B * nrows)
lib = _lib.load()
dev = torch.device("cuda:0")
f32 = lambda t: t.to(device=dev, dtype=torch.float32).contiguous()
leaves = [f32(t) for t in params_to_leaves(params)]
names = ["nbar", "lse", "g_W3", "g_b3", "g_U", "g_kappa", "g_lam", "g_sigma", "dv0", "dk", "dahat"]
res = {}
for prec in (0, 1):
    desc = gpu_helpers.desc_for(cfg, B, C, Z, prec)
    n = lib.enf_xattn_workspace_bytes(ctypes.byref(desc))
    ws = torch.zeros(n, dtype=torch.uint8, device=dev)
    out = torch.empty(B, C, cfg.num_out, device=dev)
    w = _weights_struct(leaves)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    xg, pg, ag, sg, dg = f32(coords), f32(p), f32(a), f32(sigma), f32(d_out)
    sig = ptr(sg) if cfg.use_gaussian_window else ctypes.c_void_p(0)
    assert lib.enf_xattn_fwd(ctypes.byref(desc), ctypes.byref(w), ptr(xg), 0, ptr(pg), ptr(ag), sig, ptr(out), ptr(ws), n, None) == 0
    grads = [torch.zeros_like(t) for t in leaves]
    gw = _weights_struct(grads)
    dp, da, ds = torch.empty_like(pg), torch.empty_like(ag), torch.empty_like(sg)
    assert lib.enf_xattn_bwd(ctypes.byref(desc), ctypes.byref(w), ptr(xg), 0, ptr(pg), ptr(ag), sig, ptr(dg), ctypes.byref(gw),
                             ptr(dp), ptr(da), ptr(ds), ptr(ws), n, None) == 0
    torch.cuda.synchronize()
    r = {k: gpu_helpers.ws_view(lib, desc, ws, k).clone().cpu() for k in names}
    r.update(out=out.cpu(), dp=dp.cpu(), da=da.cpu(), ds=ds.cpu())
    for nm, t in zip(_lib.LEAVES, grads):
        r["gw_" + nm] = t.cpu()
    res[prec] = r
    del ws
print(f"== {name} rows {nrows}: tensor-core vs fp32 (max-norm relative)")
for k in res[0]:
    e = rel_err(res[1][k], res[0][k])
    print(f"   {k:14s} {e:.3e}" + ("   <--" if e > 3e-3 else ""))
gu0, gu1 = res[0]["g_U"].reshape(B, Z, cfg.num_heads, cfg.num_hidden), res[1]["g_U"].reshape(B, Z, cfg.num_heads, cfg.num_hidden)
err = (gu1 - gu0).abs().amax(dim=(0, 2, 3)) / gu0.abs().max()
print("   g_U worst latents:", [(int(i), f"{float(v):.2e}") for v, i in zip(*torch.topk(err, 5))])
