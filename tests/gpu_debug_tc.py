"""Manual diagnostic (not a pytest): per-stage errors of the tensor-core path vs the fp64 folded model."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import enf_ref as R
from helpers import make_case
from gpu_helpers import run_stages

kw = dict(num_in=2, num_hidden=128, num_heads=2, num_out=1, latent_dim=16, invariant_type="rel_pos_periodic",
          embedding_freq_multiplier=(0.05, 0.1))
B, C, Z = int(os.environ.get("B", 2)), int(os.environ.get("C", 75)), int(os.environ.get("Z", 16))
cfg = R.EnfConfig(**kw)
params, x, p, a, sigma, d_out = make_case(cfg, B, C, Z, seed=3)
for prec in (0, 1):
    _, errs = run_stages(cfg, params, x, p, a, sigma, d_out, precision=prec)
    print("precision", prec)
    for k, v in errs.items():
        if v > (1e-5 if prec == 0 else 2e-4):
            print(f"   {k:12s} {v:.3e}")
