"""Diagnostics (GPU): does the two-term split of the relu GEMMs in backward kernels B / C change the value-path / query-path
weight gradients?  Variations of one fuzz problem; ENF_DEBUG_NOSPLIT toggled in-process."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import enf_ref as R          # noqa: E402
from helpers import make_case            # noqa: E402
import gpu_helpers                       # noqa: E402

base = dict(H=1, B=1, C=513, Z=49, O=2, L=16, fv=0.1, inv="rel_pos", seed=101)
variants = [dict(), dict(H=2), dict(B=2), dict(Z=16), dict(O=1), dict(fv=0.05), dict(C=512), dict(seed=1), dict(inv="rel_pos_periodic"),
            dict(seed=1, O=1), dict(seed=1, fv=0.05), dict(seed=1, H=2), dict(seed=1, Z=16)]
K = ["dp", "gw_v_w1", "gw_v_b1", "gw_q_w1", "gw_q_b1"]
for v in variants:
    c = {**base, **v}
    cfg = R.EnfConfig(num_in=2, num_hidden=128, num_heads=c["H"], num_out=c["O"], latent_dim=c["L"], invariant_type=c["inv"],
                      embedding_freq_multiplier=(0.05, c["fv"]))
    case = make_case(cfg, c["B"], c["C"], c["Z"], seed=100 + c["seed"])
    row = []
    for ns in ("0", "1"):
        os.environ["ENF_DEBUG_NOSPLIT"] = ns
        _, errs = gpu_helpers.run_stages(cfg, *case, precision=1)
        row.append(" ".join(f"{errs[k]:8.2e}" for k in K))
    print(f"{str(v):34s} split: {row[0]}   nosplit: {row[1]}")
