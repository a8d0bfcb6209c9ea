"""tf32 tcgen05/TMA stage GEMMs (enf_gemm_tc.cu) against torch fp64 matmul (`-m gpu`).
The product kernel runs a 3-term tf32 split (fp32-level accuracy); the weight-gradient kernel a single tf32 product
on operands rounded to nearest in shared memory."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _trunc_tf32(t):
    return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def _gelu(x):
    return 0.5 * x * (1 + torch.tanh(0.7978845608028654 * (x + 0.044715 * x ** 3)))


def _gelu_grad(x):
    c = 0.7978845608028654
    t = torch.tanh(c * (x + 0.044715 * x ** 3))
    return 0.5 * (1 + t) + 0.5 * x * (1 - t * t) * c * (1 + 3 * 0.044715 * x * x)


def _call(lib, use_tc, M, N, K, A, a_st, B, b_st, C, b_lo=None, bias=None, aux=None, gout=None, acc=0):
    ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    rc = lib.enf_debug_gemm(use_tc, M, N, K, ptr(A), a_st[0], a_st[1], ptr(B), b_st[0], b_st[1], ptr(b_lo), ptr(C), C.stride(0),
                            ptr(bias), ptr(aux), ptr(gout), acc, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 1, rc
    torch.cuda.synchronize()


@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (1000, 256, 256), (4096 + 77, 128, 256), (2048, 256, 128), (640, 64, 128)])
@pytest.mark.parametrize("b_transposed", [False, True])
def test_product_gemm_split_precision(M, N, K, b_transposed):
    from enf_pde_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 1e-5).cuda()                # gradient-sized values: tf32 keeps fp32's range
    W = (torch.randn(N, K, generator=g) if b_transposed else torch.randn(K, N, generator=g)).cuda()
    W_lo = W - _trunc_tf32(W)
    bias = (torch.randn(N, generator=g) * 1e-5).cuda()
    aux = torch.randn(M, N, generator=g).cuda()
    C = torch.full((M, N), float("nan"), device="cuda")
    C2 = torch.full((M, N), float("nan"), device="cuda")
    b_st = (1, K) if b_transposed else (N, 1)
    Bm = W.double().T if b_transposed else W.double()
    _call(lib, 1, M, N, K, A, (K, 1), W, b_st, C, b_lo=W_lo, bias=bias, gout=C2)
    want = A.double() @ Bm + bias.double()
    scale = float(want.abs().max())
    assert float((C.double() - want).abs().max()) / scale < 1e-5
    assert float((C2.double() - _gelu(want)).abs().max()) / scale < 1e-3      # tanh.approx in the epilogue
    _call(lib, 1, M, N, K, A, (K, 1), W, b_st, C, b_lo=W_lo, aux=aux)          # dgrad-style epilogue: * gelu'(aux)
    want = (A.double() @ Bm) * _gelu_grad(aux.double())
    assert float((C.double() - want).abs().max()) / float(want.abs().max()) < 1e-3


@pytest.mark.parametrize("K1,N,Mr", [(128, 128, 4096), (256, 256, 5000), (256, 128, 33000), (128, 256, 130), (128, 64, 2048)])
def test_weight_gradient_gemm(K1, N, Mr):
    from enf_pde_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(K1 + N + Mr)
    A = torch.randn(Mr, K1, generator=g).cuda()
    G = (torch.randn(Mr, N, generator=g) * 1e-6).cuda()
    C0 = torch.randn(K1, N, generator=g).cuda() * 1e-6
    C = C0.clone()
    _call(lib, 1, K1, N, Mr, A, (1, K1), G, (N, 1), C, acc=1)
    want = C0.double() + A.double().T @ G.double()
    # single tf32 product, operands rounded to nearest: relative error of a random-sign sum ~ 2^-11
    assert float((C.double() - want).abs().max()) / float(want.abs().max()) < 1e-3
    # and unbiased: the error does not shrink the result systematically
    assert abs(float(((C.double() - want) * want).sum() / (want * want).sum())) < 5e-5
