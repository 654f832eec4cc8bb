"""GPU: the tcgen05 / TMEM building block (kind::tf32 UMMA, SWIZZLE_128B K-major operands) that the
DFT-as-GEMM spectral stage is built on, against a torch reference."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _tf32(x):  # what the tensor core sees: low 13 mantissa bits dropped
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("N,K", [(128, 32), (128, 128), (64, 96), (256, 64), (16, 32)])
def test_tcgen05_tf32_gemm_selftest(N, K):
    from openasr_b200 import _capi
    lib = _capi.load()
    g = torch.Generator().manual_seed(N * 1000 + K)
    A = _tf32(torch.randn(128, K, generator=g)).cuda()
    B = _tf32(torch.randn(N, K, generator=g)).cuda()
    D = torch.full((128, N), float("nan"), device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    rc = lib.spl_tc_selftest(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(D.data_ptr()), N, K,
                             C.c_void_p(status.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _capi.check(rc, "spl_tc_selftest")
    torch.cuda.synchronize()
    assert status.item() == 0, "MMA completion barrier timed out"
    ref = A.double() @ B.double().t()
    err = (D.double() - ref).abs().max().item()
    assert err < 1e-4 * max(1.0, ref.abs().max().item()), err
