"""tcgen05 bring-up: one 128-row tile GEMM per operand-major combination the kernels use (`-m gpu`)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("D", [128, 64])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_tcgen05_tile_gemm_matches_fp16_matmul(mode, D):
    # mode 2 at D = 64 is an M = 64 MMA: its accumulator rows sit in lanes 32 (r / 16) + r % 16 (pinned here)
    from enf_pde_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(mode * 7 + D)
    X = torch.randn(128, D, generator=g).cuda()
    Y = torch.randn(128 if mode == 2 else D, D, generator=g).cuda()
    out = torch.full((128, D), float("nan"), device="cuda")
    scratch = torch.empty(D * D * 2, dtype=torch.uint8, device="cuda")
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.enf_debug_tc_gemm(mode, D, ptr(X), ptr(Y), ptr(out), ptr(scratch), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    Xb, Yb = X.half().double(), Y.half().double()      # operands are rounded to fp16 (tc::kOperandFmt)
    want = (Xb @ Yb.T) if mode == 0 else (Xb @ Yb) if mode == 1 else (Xb.T @ Yb)
    got = out[: want.shape[0]].double()
    err = float((got - want).abs().max()) / float(want.abs().max())
    if err >= 1e-5:
        bad = ((got - want).abs() > 1e-3 * want.abs().max()).nonzero()
        print("mismatching (row, col) sample:", bad[:12].tolist(), "count", len(bad), "rows", sorted(set(bad[:, 0].tolist()))[:40])
    assert err < 1e-5, err


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
def test_tcgen05_tile_gemm_num_hidden_32(mode):
    """num_hidden = 32: operand rows keep the 128-byte pitch of the larger kernels and are half used.  Pins which MMA shapes
    work on such rows: K-major with two K-steps (mode 0), MN-major B with N = 32 or 64 (modes 1 / 3), the M = 64 weight-gradient
    shape with N = 32 or 64 (modes 2 / 4)."""
    from enf_pde_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(mode * 5 + 1)
    D = 32
    X = torch.randn(128, D, generator=g).cuda()
    Y = torch.randn(128 if mode in (2, 4) else D, D, generator=g).cuda()
    out = torch.full((128, 64), float("nan"), device="cuda")
    scratch = torch.empty(8192, dtype=torch.uint8, device="cuda")
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.enf_debug_tc_gemm(mode, D, ptr(X), ptr(Y), ptr(out), ptr(scratch), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    Xb, Yb = X.half().double(), Y.half().double()
    want = (Xb @ Yb.T) if mode == 0 else (Xb @ Yb) if mode in (1, 3) else (Xb.T @ Yb)
    got = out[: want.shape[0], :D].double()
    err = float((got - want).abs().max()) / float(want.abs().max())
    print("mode", mode, "err", err)
    assert err < 1e-5, err
    if mode in (3, 4):
        assert float(out[: want.shape[0], D:64].abs().max()) == 0.0          # the zero half of the rows
