"""GPU parity of the tcgen05/TMA GEMM (C ABI: mnx_test_gemm_bf16) against a plain PyTorch fp32
reference of the same op on bf16-rounded operands.  fp32 accumulation on both sides: tolerance
covers summation order only (and bf16 output rounding for the bf16 epilogues)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(M, N, K, epi, bias=True, seed=0):
    from molnextr_b200 import _cabi
    lib = _cabi.load()
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = torch.randn((M, K), generator=g).cuda()
    W = (torch.randn((N, K), generator=g) / K ** 0.5).cuda()
    b = torch.randn((N,), generator=g).cuda() if bias else None
    out = torch.full((M, N), float("nan"), device="cuda")
    rc = lib.mnx_test_gemm_bf16(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()),
                                C.c_void_p(b.data_ptr() if bias else 0), C.c_void_p(out.data_ptr()), M, N, K, epi,
                                C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    ref = A.bfloat16().float() @ W.bfloat16().float().t()
    if bias:
        ref = ref + b
    return out, ref


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 256, 192), (1000, 384, 128), (144, 1024, 4096), (4608, 3072, 1024)])
def test_gemm_f32_epilogue(M, N, K):
    out, ref = _run(M, N, K, epi=3)
    assert torch.isfinite(out).all()
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=2e-4 * K ** 0.5)


def test_gemm_no_bias():
    out, ref = _run(257, 128, 256, epi=3, bias=False)
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=5e-3)


def test_gemm_bf16_and_gelu_epilogues():
    out, ref = _run(515, 512, 128, epi=0)
    torch.testing.assert_close(out, ref.bfloat16().float(), rtol=1.6e-2, atol=1e-2)
    out, ref = _run(515, 512, 128, epi=1)
    torch.testing.assert_close(out, torch.nn.functional.gelu(ref).bfloat16().float(), rtol=1.6e-2, atol=1e-2)


def test_gemm_rejects_bad_shapes():
    from molnextr_b200 import _cabi
    lib = _cabi.load()
    x = torch.zeros(64 * 100, device="cuda")
    p = C.c_void_p(x.data_ptr())
    assert lib.mnx_test_gemm_bf16(p, p, None, p, 10, 100, 64, 3, None) == -1     # N % 128 != 0
    assert lib.mnx_test_gemm_bf16(p, p, None, p, 10, 128, 40, 3, None) == -1     # K % 64 != 0
