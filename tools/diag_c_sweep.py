"""Diagnostics (GPU): the tensor-core path on one fixed problem while only C varies -- separates shape-dependent defects
(ragged tiles) from operand-precision effects.   python tools/diag_c_sweep.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import enf_ref as R          # noqa: E402
from helpers import make_case            # noqa: E402
import gpu_helpers                       # noqa: E402

KEYS = ["out", "dp", "da", "dsigma", "gw_v_w1", "gw_v_b1", "gw_q_w1", "gw_q_b1", "gf_Wp", "gf_bp", "g_U", "g_W3"]
for inv, H, B, Z in (("rel_pos", 1, 1, 49), ("rel_pos_periodic", 2, 4, 4)):
    cfg = R.EnfConfig(num_in=2, num_hidden=128, num_heads=H, num_out=1, latent_dim=16, invariant_type=inv,
                      embedding_freq_multiplier=(0.05, 0.05))
    print(f"== {inv} H={H} B={B} Z={Z}:  " + " ".join(f"{k:>9s}" for k in KEYS))
    for C in (384, 512, 513, 514, 576, 640, 641):
        params, x, p, a, sigma, d_out = make_case(cfg, B, 641, Z, seed=101)
        x, d_out = x[:, :C].contiguous(), d_out[:, :C].contiguous()
        _, errs = gpu_helpers.run_stages(cfg, params, x, p, a, sigma, d_out, precision=1)
        print(f"   C={C:4d}          " + " ".join(f"{errs[k]:9.2e}" for k in KEYS))
