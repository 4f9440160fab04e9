"""Bring-up checks of the TMA / tcgen05 / TMEM plumbing through the C ABI (no soft-max arithmetic)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from dinosoft_b200 import _cabi

    return _cabi, _cabi.lib()


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 384, 512), (200, 300, 72), (128, 1000, 768), (256, 256, 64),
                                   (512, 1100, 256)])
def test_selftest_gemm_matches_torch(M, N, K):
    cabi, lib = _lib()
    torch.manual_seed(0)
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    c = torch.full((M, N), float("nan"), device="cuda")
    cabi.check(lib.dsoft_selftest_gemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K,
                                       torch.cuda.current_stream().cuda_stream), "selftest_gemm")
    torch.cuda.synchronize()
    ref = a.double() @ b.double().T
    err = (c.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, err


@pytest.mark.parametrize("M,N,K,F", [(128, 128, 64, 64), (256, 512, 128, 256), (130, 300, 64, 320), (128, 256, 512, 768)])
def test_selftest_chain_matches_torch(M, N, K, F):
    """out = fp16(A.B^T).V : swizzled G store + MN-major fp16 second GEMM + chunked accumulator drain."""
    cabi, lib = _lib()
    torch.manual_seed(1)
    a = (torch.randn(M, K, device="cuda") / K ** 0.5).bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    v = torch.randn(N, F, device="cuda").half()
    out = torch.full((M, F), float("nan"), device="cuda")
    cabi.check(lib.dsoft_selftest_chain(a.data_ptr(), b.data_ptr(), v.data_ptr(), out.data_ptr(), M, N, K, F,
                                        torch.cuda.current_stream().cuda_stream), "selftest_chain")
    torch.cuda.synchronize()
    g = (a.float() @ b.float().T).half().double()
    ref = g @ v.double()
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 3e-4, err  # fp16 re-rounding of G can differ by one ulp where fp32 sums differ in the last bit
