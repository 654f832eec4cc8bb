"""GPU: the tcgen05 / TMEM building block (kind::tf32 UMMA, SWIZZLE_128B K-major operands) that the
DFT-as-GEMM spectral stage is built on, against a torch reference."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _tf32(x):  # what the tensor core sees: low 13 mantissa bits dropped
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("N,K", [(128, 32), (128, 128), (64, 96), (256, 64), (16, 32),
                                 (128, 8), (128, 24), (64, 40), (256, 72)])  # K % 32 != 0: SWIZZLE_32B tiles
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


def _pad_batch(ws):
    lens = [w.shape[0] for w in ws]
    x = torch.zeros(len(ws), max(lens))
    for i, w in enumerate(ws):
        x[i, :lens[i]] += w
    return x, lens


@pytest.mark.parametrize("sr,D,energy", [(16000, 80, False), (16000, 40, True), (8000, 40, False)])
def test_tcgen05_dft_kernel_matches_oracle(monkeypatch, wavs, sr, D, energy):
    """The DFT-as-GEMM engine of kernel A (FP16 hi/lo split on tcgen05, the default engine) on real speech:
    same tolerance as the FFT kernels (|d| <= 1e-3 + 1e-4 |ref|), exact lengths and zero padding."""
    from oracle import frontend_oracle as fo
    from openasr_b200 import SPLayer
    monkeypatch.setenv("SPL_ENGINE", "umma")
    dec = 16000 // sr
    x, lens = _pad_batch([wavs[0][::dec].contiguous(), wavs[1][::dec].contiguous(), wavs[0][::dec][:7001].contiguous()])
    conf = {"feature_type": "fbank", "sample_rate": sr, "num_mel_bins": D, "use_energy": energy, "dither": 0.0}
    layer = SPLayer(conf).cuda().eval()
    assert layer._handle(torch.device("cuda", 0)).engine_name() == ("fft" if energy else "umma")
    feats, flen = layer(x.cuda(), lens)
    torch.cuda.synchronize()
    assert layer._handle(torch.device("cuda", 0)).debug_status() == 0
    ref, rlen = fo.splayer_forward(x, lens, conf)
    ref64, _ = fo.splayer_forward(x, lens, conf, dtype=torch.float64)
    assert torch.equal(flen.cpu(), rlen)
    f = feats.cpu()
    assert torch.isfinite(f).all()
    d = (f - ref).abs()
    tol = 1e-3 + 1e-4 * ref.abs() + 2.0 * (ref.double() - ref64).abs().float()
    print("tcgen05 DFT: max|d| vs ref32 %.3g, vs fp64 %.3g (ref32 vs fp64 %.3g)" % (
        d.max().item(), (f.double() - ref64).abs().max().item(), (ref.double() - ref64).abs().max().item()))
    assert (d <= tol).all(), "max|d|=%g" % d.max().item()
    for i, m in enumerate(rlen.tolist()):
        assert (f[i, m:] == 0).all()
