# ENF_DEBUG_TRACE=1: dump the clock64 trace of kernel A (CTA 7, tiles 4..7)
import os, sys, types, ctypes
os.environ["ENF_DEBUG_TRACE"] = "1"
sys.path.insert(0, "/root/repo")
import torch
import bench
import enf_pde_b200 as E
from enf_pde_b200 import _lib
cfg = bench.CONFIGS["ns64"]
dev = torch.device("cuda", 0)
inv = E.get_ca_invariant(types.SimpleNamespace(invariant_type=cfg["invariant_type"], num_in=cfg["num_in"]))
nef = E.EquivariantCrossAttentionNeF(cfg["d"], cfg["H"], 0, cfg["O"], cfg["L"], inv, inv, "rff", cfg["freq"], True, cfg["window"], precision="bf16")
B, Z = cfg["B"], cfg["Z"]
p_h, a_h, s_h = E.init_latents(inv, B, Z, cfg["L"], polar_grid=cfg["polar_grid"])
from enf_pde_b200.latents import make_coords
x = make_coords(inv, cfg["grid"]).contiguous().to(dev)
variables = nef.init(0, x[None].cpu(), p_h.to(dev), a_h.to(dev), s_h.to(dev))
p, a, s = (t.to(dev).requires_grad_(True) for t in (p_h, a_h, s_h))
for it in range(3):
    out = nef.apply(variables, x[None].expand(B, -1, -1), p, a, s)
    out.backward(torch.randn_like(out))
torch.cuda.synchronize()

from enf_pde_b200.nef import _XAttnFunction
desc_kw, ws = _XAttnFunction.last_ws
lib = _lib.load()
desc = _lib.EnfDesc(**desc_kw)
n = ctypes.c_int64(0)
off = lib.enf_debug_ws_offset(ctypes.byref(desc), b"dbg", ctypes.byref(n))
t = ws[off:off + 8192 * 4].view(torch.int64).cpu().numpy()
import numpy as np
ep = t[:128].reshape(4, 32); iw = t[512:512 + 128].reshape(4, 32)
base = ep[0, 0]
np.set_printoptions(linewidth=250)
print("epilogue thread 32: slots 0:tile top 1:after gb wait | head h: 2+8h after g4 wait, 3 ld done, 4 gelu pass done, 5 exchange done, 6 dm stored+arrive, 7 colsum+loads | 18 after bar_d, 19 ld done, 20 stored")
for r in range(4):
    print("tile", 4 + r, [int(v - base) if v else None for v in ep[r, :27]])
print("issue warp: 0 top, 1 after DTH sync, 2/6 before head sync, 3/7 after head sync, 4/8 issued")
for r in range(4):
    print("tile", 4 + r, [int(v - base) if v else None for v in iw[r, :9]])

off = lib.enf_debug_ws_offset(ctypes.byref(desc), b"dbg_fwd", ctypes.byref(n))
t = ws[off:off + 8192 * 4].view(torch.int64).cpu().numpy()
ep = t[:128].reshape(4, 32)
base = ep[0, 0]
print("fwd thread 32: 0 top 1 phases ready 2 E0 done 3 sync1 4 next-inv done 5 g1 ready 6 E1 done 7 g2 ready 8 E2 stored 9 sync2 10 softmax done 11 g3 ready 12 gelu done 13 exch 14 LN stored 15 sync3 16 g4_0 ready 17 E4_0 done 18 g4_1 ready 19 E4_1 done 20 end sync")
for r in range(4):
    print("z", 8 + r, [int(v - base) if v else None for v in ep[r, :21]])

off = lib.enf_debug_ws_offset(ctypes.byref(desc), b"dbg", ctypes.byref(n))
t = ws[off:off + 8192 * 4].view(torch.int64).cpu().numpy()
for who, o in (("thread 32 (cq=0: row side work)", 1024), ("thread 160 (cq=1)", 1024 + 128)):
    ep = t[o:o + 128].reshape(4, 32)
    base = t[1024]
    print("Q", who, ": 0 top 1 prev MMAs done 2 phases ready 3 S1 done 4 sync 5 side work done 6 g1 ready 7 E done 8 sync 9 dgrad ready 10 S3 done 11 sync")
    for r in range(4):
        print("tile", 4 + r, [int(v - base) if v else None for v in ep[r, :12]])

for who, o in (("thread 32 (cq=0)", 2048), ("thread 160 (cq=1)", 2048 + 128)):
    ep = t[o:o + 128].reshape(4, 32)
    base = t[2048]
    print("V", who, ": 0 top 1 side/prev done 2 phases 3 S1 done 4 sync 5 inv done 6 g1 7 E2 done 8 sync 9 g2 10 E3 done 11 sync 12 g3 13 E4 ld 14 g3b 15 E4 stored 16 sync 17 g4 18 S3 done 19 sync")
    for r in range(4):
        print("tile", 4 + r, [int(v - base) if v else None for v in ep[r, :20]])
