"""Diagnostics (GPU): per-(query, latent) hand-over tensors of the tensor-core backward (ds: A -> C, duv: B -> C) against the
float64 folded model, to see WHERE an error sits (which latents, which query tiles).   python tools/diag_pairs.py fuzz1"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gpu_helpers                       # noqa: E402
from folded_model import ln_core_bwd, gelu_grad    # noqa: E402
import diag_stages                       # noqa: E402
from enf_pde_b200 import _lib            # noqa: E402

name = sys.argv[1]
cfg, params, x, p, a, sigma, d_out = diag_stages.case(name)
res, errs = gpu_helpers.run_stages(cfg, params, x, p, a, sigma, d_out, precision=1)
lib = _lib.load()
m, ws, desc = res["model"], res["ws"], res["desc"]
B, C, Z, H, d = x.shape[0], x.shape[1], p.shape[1], cfg.num_heads, cfg.num_hidden
view = lambda n: gpu_helpers.ws_view(lib, desc, ws, n).double().cpu()
gmax = float(view("gmax")[0])
e = math.frexp(gmax)[1]
gs = 2.0 ** (4 - e)
S, L, f, w = m.S, m.L, m.f, m.w
dnbar, _, _ = m.tail_bwd(S["nbar"], d_out.double())
att, n = S["att"], S["n"]
Dd = (dnbar * S["nbar"]).sum(-1)
ds = att * (torch.einsum("bchj,bczhj->bczh", dnbar, n) - Dd[:, :, None])
dn = att[..., None] * dnbar[:, :, None]
dmpre = ln_core_bwd(dn, n, S["n_rstd"]) * gelu_grad(S["mpre"])
dthat = torch.einsum("bczhj,bzhij->bczi", dmpre, L["W3"])
dtpre = ln_core_bwd(dthat, S["that"], S["t_rstd"]) * gelu_grad(S["tpre"])
dzv = (dtpre @ f["Wp"].T) * (S["h1v"] > 0)
dgv = dzv @ w["v_w1"].T
hd = d // 2
sin, cos = S["gv"][..., :hd], S["gv"][..., hd:]
du_v = 2 * math.pi * ((cos * dgv[..., :hd] - sin * dgv[..., hd:]) @ w["v_omega"].T)       # (B,C,Z,I)
I = du_v.shape[-1]
ds_gpu = view("ds_tc")[: B * Z * C * H].reshape(B, Z, C, H).permute(0, 2, 1, 3) / gs
duv_gpu = view("duv")[: B * Z * C * 8].reshape(B, Z, C, 8).permute(0, 2, 1, 3)[..., :I] / gs
print(f"== {name}: gmax {gmax:.3e} gs 2^{4 - e}")
for nm, got, ref in (("ds", ds_gpu, ds), ("du_v", duv_gpu, du_v)):
    err = (got - ref).abs()
    print(f"   {nm}: max|err|/max|ref| = {err.max() / ref.abs().max():.3e}   rms err / rms ref = {err.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt():.3e}")
    per_z = err.amax(dim=(0, 1, 3)) / ref.abs().max()
    top = torch.topk(per_z, min(5, Z))
    print("      worst latents:", [(int(i), f"{float(v):.2e}") for v, i in zip(top.values, top.indices)])
    per_tile = torch.stack([err[:, t * 128:(t + 1) * 128].amax() for t in range((C + 127) // 128)]) / ref.abs().max()
    print("      per query tile:", [f"{float(v):.2e}" for v in per_tile])
    # fraction of the total squared error carried by the 0.1 % worst pairs
    flat = err.pow(2).sum(-1).flatten()
    k = max(1, flat.numel() // 1000)
    print(f"      share of squared error in the worst 0.1% of pairs: {float(torch.topk(flat, k).values.sum() / flat.sum()):.3f}")
# ---- zoom: the worst (latent, tile) of du_v, row by row ------------------------------------------------------------------
err = (duv_gpu - du_v).abs().amax(-1)[0]            # (C, Z) of field 0
zc = int(err.amax(0).argmax())
t = int(err[:, zc].argmax()) // 128
rows = slice(t * 128, min(C, (t + 1) * 128))
print(f"   zoom: field 0, latent {zc} (p = {p[0, zc].tolist()}, sigma = {float(sigma[0, zc, 0]):.4f}), tile {t}")
scale = float(du_v.abs().max())
e_rows = err[rows, zc] / scale
print("      rows with err > 1e-2 of max|du_v|:", [(int(i), f"{float(v):.2e}") for i, v in enumerate(e_rows) if v > 1e-2])
bad = [int(i) for i, v in enumerate(e_rows) if v > 1e-2][:6]
for r in bad:
    c = t * 128 + r
    print(f"      row {r}: x = {x[0, c].tolist()} u = {S['u'][0, c, zc].tolist()} att = {att[0, c, zc].tolist()}")
    print(f"              du_v gpu {duv_gpu[0, c, zc].tolist()}  ref {du_v[0, c, zc].tolist()}")
    pre = S["gv"][0, c, zc] @ w["v_w1"] + w["v_b1"]
    print(f"              min |pre-activation h1v| = {float(pre.abs().min()):.3e}; #units with |pre| < 1e-3: {int((pre.abs() < 1e-3).sum())}; dzv nonzeros {int((dzv[0, c, zc] != 0).sum())}")
    print(f"              |dtpre| max {float(dtpre[0, c, zc].abs().max()):.3e} (scaled {float(dtpre[0, c, zc].abs().max()) * gs:.3e}); |dthat| max scaled {float(dthat[0, c, zc].abs().max()) * gs:.3e}; t_rstd {float(S['t_rstd'][0, c, zc]):.3e}")
