"""GPU: round-2 entry points -- several batches per launch (spl_forward_multi), sync-free device lengths, SpecAug
resolved in kernel B from the uploaded uniforms, the device dither generator itself, and the multi-GPU pieces
(NCCL all-reduce of the global-CMVN statistics, DataParallel replicas)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import frontend_oracle as fo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SA = {"freq_mask_num": 2, "freq_mask_width": 27, "time_mask_num": 2, "time_mask_width": 40}


def make_layer(**kw):
    from openasr_b200 import SPLayer
    conf = {"feature_type": "fbank", "sample_rate": 16000, "num_mel_bins": 80, "use_energy": False, "dither": 0.0}
    conf.update(kw)
    return SPLayer(conf).cuda(), conf


@pytest.mark.parametrize("engine", ["fft", "umma"])
def test_forward_multi_equals_consecutive_forwards(monkeypatch, engine):
    """K batches in one launch of each kernel == K forward calls: lengths, padding, CMVN, SpecAug rectangles
    (same host RNG stream).  The FFT engine is bit-identical (batch position does not enter its arithmetic); the
    tcgen05 engine scales / pivots per 8-frame group of the flattened frame list, so it agrees within the tolerance."""
    monkeypatch.setenv("SPL_ENGINE", engine)
    layer, conf = make_layer(cmvn="utterance", spec_aug=SA)
    layer.train()
    assert layer._handle(torch.device("cuda", 0)).engine_name() == engine
    batches = [fo.synth_batch(5 + k, 3000, 60000, 16000, seed=50 + k) for k in range(4)]
    torch.manual_seed(3)
    single = [layer(w.cuda(), l) for w, l in batches]
    torch.manual_seed(3)
    multi = layer.forward_multi([(w.cuda(), l) for w, l in batches])
    torch.cuda.synchronize()
    assert layer._handle(torch.device("cuda", 0)).debug_status() == 0
    for (fs, ls), (fm, lm), (w, l) in zip(single, multi, batches):
        assert torch.equal(ls, lm) and fs.shape == fm.shape
        for i, m in enumerate(lm.tolist()):  # padding rows: zero, or the (tiny) time mean where a time mask spills
            assert (fs[i, m:] - fm[i, m:]).abs().max().item() < 1e-4 if m < fs.shape[1] else True
        if engine == "fft":
            assert torch.equal(fs, fm)
        else:
            # normalised features: errors scale with 1/sigma_d; masks identical (checked through the zero pattern
            # of padding and exact equality of the mask-fill positions below)
            assert (fs - fm).abs().max().item() < 2e-2 and (fs - fm).abs().mean().item() < 1e-4
    # against the oracle with the same uniforms
    torch.manual_seed(3)
    for (fm, lm), (w, l) in zip(multi, batches):
        uni = torch.rand(8, w.shape[0])
        ref, rl = fo.splayer_forward(w, l.tolist(), conf, training=True, specaug_uniforms=uni)
        assert torch.equal(lm.cpu(), rl)
        d = (fm.cpu() - ref).abs()
        assert d.max().item() < 3e-2 and d.mean().item() < 2e-4, (d.max().item(), d.mean().item())


def test_large_launch_ten_full_batches():
    """A launch of the size bench.py times (ten AISHELL-shaped batches, 40 k groups, every CTA's share spans several
    utterances and batches): same features as batch-by-batch calls, bit for bit without CMVN (the arithmetic of a
    group does not depend on which warp runs it), CMVN sums within fp64 rounding, repeated calls agree, padding exact."""
    layer, conf = make_layer()
    layer.eval()
    batches = [fo.synth_batch(32, 56000, 104000, 16000, seed=900 + k) for k in range(10)]
    dev_batches = [(w.cuda(), l) for w, l in batches]
    single = [layer(w, l) for w, l in dev_batches]
    multi = layer.forward_multi(dev_batches)
    for (fs, ls), (fm, lm) in zip(single, multi):
        assert torch.equal(ls, lm) and torch.equal(fs, fm)
    for _ in range(5):
        again = layer.forward_multi(dev_batches)
        for (fm, lm), (fa, la) in zip(multi, again):
            assert torch.equal(fm, fa) and torch.equal(lm, la)
    w0, l0 = batches[9]   # the last batch of the launch against the oracle
    ref, rl = fo.splayer_forward(w0[:4], l0[:4].tolist(), conf)
    d = (multi[9][0][:4, :ref.shape[1]].cpu() - ref).abs()
    assert d.max().item() < 1e-2 and d.mean().item() < 2e-5
    layer_c, conf_c = make_layer(cmvn="utterance")
    layer_c.eval()
    sc = [layer_c(w, l) for w, l in dev_batches]
    mc = layer_c.forward_multi(dev_batches)
    for (fs, ls), (fm, lm) in zip(sc, mc):
        assert torch.equal(ls, lm) and (fs - fm).abs().max().item() < 1e-4
        for i, m in enumerate(lm.tolist()):
            assert (fm[i, m:] == 0).all()


def test_sync_free_device_lengths_match_host_lengths():
    """`sync_free`: CUDA lengths are never read back; T comes from the padded width (== longest utterance, as the
    reference's collate guarantees) and the masks are resolved in kernel B -> identical to the host-length path."""
    layer_h, conf = make_layer(cmvn="utterance", spec_aug=SA)
    layer_d, _ = make_layer(cmvn="utterance", spec_aug=SA, sync_free=True)
    layer_h.train()
    layer_d.train()
    w, l = fo.synth_batch(7, 2000, 40000, 16000, seed=9)
    w = w[:, :int(l.max())].contiguous()
    torch.manual_seed(4)
    fh, lh = layer_h(w.cuda(), l)
    torch.manual_seed(4)
    fd, ld = layer_d(w.cuda(), l.cuda())
    assert torch.equal(lh, ld) and torch.equal(fh, fd)


def test_kernel_b_uniforms_equal_host_rectangles():
    """SpecAug from uniforms (device-resolved against feat_len) == SpecAug from the host-resolved rectangles of
    spl_specaug_rects, including the len < width quirk (negative starts, spill into padding rows)."""
    from openasr_b200 import frontend
    g = torch.Generator().manual_seed(2)
    B, T, V = 9, 120, 80
    flen = torch.tensor([120, 3, 17, 39, 40, 41, 80, 1, 119])
    x = torch.randn(B, T, V, generator=g)
    x = x * (torch.arange(T)[None, :, None] < flen[:, None, None])
    conf = {"freq_mask_num": 2, "freq_mask_width": 27, "time_mask_num": 3, "time_mask_width": 40}
    uni = torch.rand(10, B, generator=g)
    rect = frontend.specaug_rectangles_c(uni.numpy(), flen.numpy(), T, V, conf)
    a, b = x.clone().cuda(), x.clone().cuda()
    fl = flen.cuda()
    st = frontend.column_stats(a, fl)
    frontend.post_inplace(a, fl, utt_stats=st, mask_params=torch.from_numpy(rect).cuda(), n_freq=2, n_time=3)
    frontend.post_inplace(b, fl, utt_stats=st, mask_uniforms=uni.cuda().contiguous(), n_freq=2, n_time=3,
                          freq_width=27.0, time_width=40.0)
    assert torch.equal(a, b)
    ref, _ = fo.spec_aug(x.clone(), flen, conf, uniforms=uni)
    assert torch.allclose(a.cpu(), ref, atol=1e-5)
    assert torch.equal(a.cpu() != x, ref != x)


@pytest.mark.parametrize("engine", ["fft", "umma"])
@pytest.mark.parametrize("dither", [0.7, 1.5, 3.0, -1.0])
def test_device_dither_any_scale_is_finite(monkeypatch, engine, dither):
    """sqrt of a slightly negative argument for dither values that are not powers of two."""
    monkeypatch.setenv("SPL_ENGINE", engine)
    layer, conf = make_layer(dither=dither, cmvn="utterance")
    layer.eval()
    w, l = fo.synth_batch(16, 30000, 60000, 16000, seed=77)
    for it in range(6):
        torch.manual_seed(it)
        f, _ = layer(w.cuda(), l)
        assert torch.isfinite(f).all()


@pytest.mark.parametrize("sr,D,dither", [(16000, 80, 1.0), (8000, 40, 1.0), (20000, 80, -0.5), (14000, 64, 2.0)])
def test_device_dither_noise_itself(sr, D, dither):
    """The device generator, sample by sample (spl_debug_dither_noise): distribution of the reference's one-uniform
    pseudo Box-Muller (kaldi_signal.py:176-177: mean 0.0576, std 1.057, range (-1.21, 5.6)), no correlation along
    the frame, none between the overlapping samples of consecutive frames (the reference draws (m, Nw) fresh
    values), and -- replayed through the host-noise mode -- the same features as the device-RNG mode.  The four
    cases cover the compiled-in windows (13 rows per lane: three Philox calls per frame pair), the generic-window
    template with more than ten rows (20 kHz) and a negative / non-unit dither."""
    from openasr_b200 import _capi
    layer, conf = make_layer(dither=dither, sample_rate=sr, num_mel_bins=D)
    layer.eval()
    dev = torch.device("cuda", 0)
    h = layer._handle(dev)
    assert h.engine_name() == "fft"
    win, shift = h.win, h.shift
    w, l = fo.synth_batch(16, 50 * shift * 6, 80 * shift * 6, sr, seed=5)   # ~1.5 M overlapping pairs: 1 sigma = 8e-4
    frames = [fo.num_frames(int(n), win, shift) for n in l.tolist()]
    T = max(frames)
    seed = 0x1234567812345678
    noise = torch.empty((16, T, win), device=dev)
    _capi.check(h._lib.spl_debug_dither_noise(h._h, C.c_void_p(noise.data_ptr()), 16, T, seed,
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)), "noise")
    g = noise.cpu().double()
    ref = fo.dither_noise((2000, 400), generator=torch.Generator().manual_seed(0)).double()
    assert abs(g.mean().item() - ref.mean().item()) < 5e-3 and abs(g.mean().item() - 0.0576) < 5e-3
    assert abs(g.std().item() - ref.std().item()) < 5e-3 and abs(g.std().item() - 1.057) < 5e-3
    assert -1.22 < g.min().item() < -1.19 and 4.0 < g.max().item() < 5.7
    for q in (0.01, 0.1, 0.5, 0.9, 0.99):  # quantiles of the two samples agree
        assert abs(torch.quantile(g.flatten()[:4_000_000], q).item() - torch.quantile(ref.flatten(), q).item()) < 2e-2
    z = g - g.mean()
    lag1 = (z[:, :, 1:] * z[:, :, :-1]).mean().item() / z.var().item()
    assert abs(lag1) < 3e-3
    # sample s of frame t is sample s - shift of frame t + 1: the two noise values must be independent
    cross = (z[:, :-1, shift:] * z[:, 1:, :win - shift]).mean().item() / z.var().item()
    assert abs(cross) < 3.5e-3
    across_utt = (z[0] * z[1]).mean().item() / z.var().item()
    assert abs(across_utt) < 4e-3
    # rows of one lane (samples j and j + R2) come from different fields of the same Philox word
    r2 = h.padded // 16
    rows = (z[:, :, r2:] * z[:, :, :-r2]).mean().item() / z.var().item()
    assert abs(rows) < 3e-3
    # replay: device RNG == host-noise mode fed with the dumped noise (covers the sign and the scale of `dither`)
    ld = l.cuda()
    f_dev, _ = h.fbank(w.cuda(), ld, T, dither_seed=seed)
    f_host, _ = h.fbank(w.cuda(), ld, T, noise=noise)
    assert (f_dev - f_host).abs().max().item() < 5e-4
    # and the oracle with that noise
    for i in (0, 3):
        r = fo.fbank(w[i, :l[i]], float(sr), D, dither=dither, noise=noise[i, :frames[i]].cpu())
        d = (f_dev[i, :frames[i]].cpu() - r).abs()
        assert (d <= 1e-3 + 1e-4 * r.abs() + 4e-3 * (d > 0)).all() and d.mean().item() < 2e-5


def test_throughput_mode_eight_warp_ctas(monkeypatch):
    """SPL_CTAS_PER_SM=1: the 8-warp variant of the FFT engine (two launches co-resident per SM).  Same features as
    the oracle; its dither evaluates the formula on a 16-bit grid (no room for the table) -- same replay check."""
    from openasr_b200 import _capi
    monkeypatch.setenv("SPL_CTAS_PER_SM", "1")
    layer, conf = make_layer(cmvn="utterance")
    layer.eval()
    x, lens = fo.synth_batch(7, 20000, 60000, 16000, seed=8)
    feats, flen = layer(x.cuda(), lens)
    ref, rlen = fo.splayer_forward(x, lens.tolist(), conf)
    assert torch.equal(flen.cpu(), rlen)
    d = (feats.cpu() - ref).abs()
    assert d.max().item() < 3e-2 and d.mean().item() < 2e-4
    monkeypatch.delenv("SPL_CTAS_PER_SM")
    layer16, _ = make_layer(cmvn="utterance")
    layer16.eval()
    f16, _ = layer16(x.cuda(), lens)
    assert torch.equal(f16, feats)  # same arithmetic per group whatever the CTA shape
    monkeypatch.setenv("SPL_CTAS_PER_SM", "1")
    layer_d, _ = make_layer(dither=1.0)
    layer_d.eval()
    h = layer_d._handle(torch.device("cuda", 0))
    frames = [fo.num_frames(int(n), 400, 160) for n in lens.tolist()]
    T = max(frames)
    seed = 0x0123456789abcdef
    noise = torch.empty((7, T, 400), device="cuda")
    _capi.check(h._lib.spl_debug_dither_noise(h._h, C.c_void_p(noise.data_ptr()), 7, T, seed,
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)), "noise")
    g = noise.double()
    assert abs(g.mean().item() - 0.0576) < 5e-3 and abs(g.std().item() - 1.057) < 5e-3 and g.max().item() > 4.3
    ld = lens.cuda()
    f_dev, _ = h.fbank(x.cuda(), ld, T, dither_seed=seed)
    f_host, _ = h.fbank(x.cuda(), ld, T, noise=noise)
    assert (f_dev - f_host).abs().max().item() < 5e-4


def test_engine_switch_and_defaults(monkeypatch):
    layer, _ = make_layer()
    assert layer._handle(torch.device("cuda", 0)).engine_name() == "fft"   # default: faster by measurement
    monkeypatch.setenv("SPL_ENGINE", "umma")
    layer, _ = make_layer()
    h = layer._handle(torch.device("cuda", 0))
    assert h.engine_name() == "umma" and h.engine_name(torch.int16) == "umma"
    layer, _ = make_layer(use_energy=True)
    assert layer._handle(torch.device("cuda", 0)).engine_name() == "fft"   # energy column: FFT engine


# ------------------------------------------------------------------------------------------------ multi-GPU
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_global_cmvn_nccl_two_ranks(tmp_path):
    """GlobalCmvn under torch.distributed / NCCL on 2 ranks: statistics accumulated in kernel A's epilogue on each
    GPU, one all_reduce, mean / istd and the normalised features against the fp64 oracle."""
    out = tmp_path / "nccl.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tests", "nccl_cmvn_worker.py"), str(out)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    import json
    res = json.load(open(out))
    assert res["world"] == 2 and res["ok"], res
    print("NCCL global CMVN:", res)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dataparallel_many_iterations():
    """nn.DataParallel replicas share the module's attributes (one HostStager): 40 iterations -- more than the 16
    slots of a ring -- on every GPU, results equal to the single-GPU module."""
    n = torch.cuda.device_count()
    layer, conf = make_layer(cmvn="utterance", spec_aug=SA)
    layer.train()

    class Probe(torch.nn.Module):  # the reference wraps the whole MODEL (train.py:134): per-replica outputs are gatherable
        def __init__(self, sp):
            super().__init__()
            self.sp = sp

        def forward(self, wav, lens):
            f, fl = self.sp(wav, lens)
            return f.abs().sum(dim=(1, 2)), fl

    dp = torch.nn.DataParallel(Probe(layer), device_ids=list(range(n)))
    w, l = fo.synth_batch(4 * n, 8000, 30000, 16000, seed=21)
    for it in range(40):
        s_, fl = dp(w.cuda(0), l.cuda(0))
        assert torch.isfinite(s_).all() and fl.shape[0] == 4 * n
    layer.eval()
    s_, fl = dp(w.cuda(0), l.cuda(0))
    fr, flr = layer(w.cuda(0), l)
    assert torch.equal(fl.cpu(), flr.cpu())
    assert torch.allclose(s_.cpu(), fr.abs().sum(dim=(1, 2)).cpu(), rtol=1e-4)
