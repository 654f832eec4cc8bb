"""GPU parity tests: the CUDA path (through the C ABI) against the committed reference vectors
and against the CPU oracle on the same seeded inputs.

Tolerance (BASELINE.json north_star): log-mel max-abs 1e-3 and |d| <= 1e-3 + 1e-4*|ref| in fp32;
frame counts / lengths / mask rectangles / zero padding exact.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import frontend_oracle as fo

pytestmark = pytest.mark.gpu

ATOL, RTOL = 1e-3, 1e-4

# Every comparison that needed the widened bound is recorded (test id, elements outside the plain bound, worst |d|)
# and written to gpurun_out/parity_widened.json at the end of the session; tests/golden/widened_baseline.json holds
# the counts measured when the kernels were last changed -- a regression from 3 to 40 elements fails.
WIDENED = {}
_BASE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "widened_baseline.json")
WIDENED_BASELINE = json.load(open(_BASE)) if os.path.isfile(_BASE) else {}


def _record_widened(n_out, n_tot, worst):
    test = os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0]
    e = WIDENED.setdefault(test, {"comparisons": 0, "elements": 0, "outside_plain_bound": 0, "max_abs_diff": 0.0})
    e["comparisons"] += 1
    e["elements"] += int(n_tot)
    e["outside_plain_bound"] += int(n_out)
    e["max_abs_diff"] = max(e["max_abs_diff"], float(worst))
    base = WIDENED_BASELINE.get(test)
    if base is not None:
        assert e["outside_plain_bound"] <= base["outside_plain_bound"] + 2, \
            "%s: %d elements outside the plain bound (baseline %d)" % (test, e["outside_plain_bound"], base["outside_plain_bound"])


@pytest.fixture(scope="session", autouse=True)
def _dump_widened():
    yield
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        json.dump(WIDENED, open(os.path.join(out, "parity_widened.json"), "w"), indent=1, sort_keys=True)
    except OSError:
        pass


def close(a, b, b64=None, scale=None):
    """|a - b| <= 1e-3 + 1e-4 |b| elementwise.  ``b64`` (the fp64 oracle) widens the bound by twice the
    reference's OWN fp32 error on ill-conditioned elements (log of a mel bin ~1e5 below the frame
    energy: the fp32 reference itself is off by up to 4e-3 there); at most 0.1 % of the
    elements may need that.  ``scale`` multiplies the bound (features normalised by 1/std)."""
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    assert a.shape == b.shape, (a.shape, b.shape)
    d = (a - b).abs()
    assert torch.isfinite(a).all()
    tol = ATOL + RTOL * b.abs()
    if scale is not None:
        tol = tol * scale
    if b64 is not None:
        self_gap = (b.double() - b64.detach().cpu().double()).abs().float()
        out = d > tol
        assert out.float().mean().item() <= 1e-3, "too many ill-conditioned elements"
        _record_widened(int(out.sum()), d.numel(), d.max().item())
        tol = tol + 2.0 * self_gap
    ok = d <= tol
    assert ok.all(), "max|d|=%g at %s" % (d.max().item(), np.unravel_index(int((d - tol).argmax()), tuple(d.shape)))
    return d.max().item()


def make_layer(**kw):
    from openasr_b200 import SPLayer
    conf = {"feature_type": "fbank", "sample_rate": 16000, "num_mel_bins": 80, "use_energy": False, "dither": 0.0}
    conf.update(kw)
    return SPLayer(conf).cuda(), conf


def pad_batch(ws):
    lens = [w.shape[0] for w in ws]
    x = torch.zeros(len(ws), max(lens))
    for i, w in enumerate(ws):
        x[i, :lens[i]] += w
    return x, lens


# ---------------------------------------------------------------------------------------------
GOLDEN_CASES = {
    "w0_d80": (0, 16000, 1, 80, False, "povey"),
    "w1_d80": (1, 16000, 1, 80, False, "povey"),
    "w0_d40": (0, 16000, 1, 40, False, "povey"),
    "w1_d40": (1, 16000, 1, 40, False, "povey"),
    "w0_d80_energy": (0, 16000, 1, 80, True, "povey"),
    "w1_d80_hamming": (1, 16000, 1, 80, False, "hamming"),
    "w1_8k_d40": (1, 8000, 2, 40, False, "povey"),
}


@pytest.mark.parametrize("key", sorted(GOLDEN_CASES))
def test_fbank_matches_reference_vectors(key, wavs, golden_dir):
    wi, sr, dec, D, en, wt = GOLDEN_CASES[key]
    ref = torch.from_numpy(np.load(os.path.join(golden_dir, "fbank_ref.npz"))[key])
    layer, _ = make_layer(sample_rate=sr, num_mel_bins=D, use_energy=en, window_type=wt)
    layer.eval()
    w = wavs[wi][::dec].contiguous()
    feats, flen = layer(w.view(1, -1).cuda(), [w.shape[0]])
    assert flen.dtype == torch.int64 and flen.tolist() == [ref.shape[0]]
    assert feats.shape == (1,) + tuple(ref.shape)
    close(feats[0], ref)


def test_known_answers(wavs, golden_dir):
    ka = json.load(open(os.path.join(golden_dir, "known_answers.json")))
    layer, _ = make_layer()
    layer.eval()
    f, _ = layer(wavs[0].view(1, -1).cuda(), [wavs[0].shape[0]])
    k = ka["w0_d80"]
    assert list(f.shape[1:]) == k["shape"]
    assert abs(f[0, 0, 0].item() - k["first"]) < 1e-3
    assert abs(f[0, 100, 40].item() - k["mid"]) < 1e-3
    assert abs(f[0, -1, -1].item() - k["last"]) < 1e-3
    assert abs(f.double().sum().item() - k["sum"]) < 1e-3 * f.numel() * 0.05
    # all-zero input: every frame is log(eps)
    for n, m in ((400, 1), (559, 1), (560, 2)):
        z, zl = layer(torch.zeros(1, n).cuda(), [n])
        assert zl.tolist() == [m] and z.shape == (1, m, 80)
        assert torch.allclose(z.cpu(), torch.full((1, m, 80), ka["zeros_%d" % n]["first"]), atol=1e-5)


def test_c1_fixture_batch_ragged_padding(wavs):
    """BASELINE configs[0]: batch 4 = [w0, w1, w0, w1], 80-dim, vs the oracle; padding exactly 0."""
    layer, conf = make_layer()
    layer.eval()
    x, lens = pad_batch([wavs[0], wavs[1], wavs[0], wavs[1]])
    feats, flen = layer(x.cuda(), torch.tensor(lens).cuda())
    ref, rlen = fo.splayer_forward(x, lens, conf)
    assert torch.equal(flen.cpu(), rlen)
    assert flen.device.type == "cuda"
    close(feats, ref)
    for i, m in enumerate(rlen.tolist()):
        assert (feats[i, m:] == 0).all()


def test_int16_ingest_equals_fp32(wavs):
    layer, _ = make_layer()
    layer.eval()
    x, lens = pad_batch([wavs[0], wavs[1]])
    a, _ = layer(x.cuda(), lens)
    b, _ = layer(x.to(torch.int16).cuda(), lens)
    assert torch.equal(a, b)


def test_host_dither_stream_parity(wavs, golden_dir):
    """dither=1 with the reference's exact CPU noise stream (seeded) against the reference vectors."""
    g = np.load(os.path.join(golden_dir, "fbank_ref.npz"))
    layer, _ = make_layer(dither=1.0, dither_rng="host")
    layer.eval()
    torch.manual_seed(7)
    f, _ = layer(wavs[0].view(1, -1).cuda(), [wavs[0].shape[0]])
    close(f[0], torch.from_numpy(g["w0_d80_dither_seed7"]))
    layer8, _ = make_layer(dither=1.0, dither_rng="host", sample_rate=8000, num_mel_bins=40, use_energy=True)
    layer8.eval()
    w = wavs[1][::2].contiguous()
    torch.manual_seed(11)
    f, _ = layer8(w.view(1, -1).cuda(), [w.shape[0]])
    close(f[0], torch.from_numpy(g["w1_8k_d40_energy_dither_seed11"]))


def test_host_dither_batch_order(wavs):
    """Batch of two: noise is consumed per utterance in batch order, like the reference loop."""
    layer, conf = make_layer(dither=1.0, dither_rng="host")
    layer.eval()
    x, lens = pad_batch([wavs[1], wavs[0]])
    torch.manual_seed(21)
    f, _ = layer(x.cuda(), lens)
    torch.manual_seed(21)
    ref, _ = fo.splayer_forward(x, lens, conf)
    close(f, ref)


def test_device_dither_statistics():
    """Throughput-mode dither: same distribution (mean 0.057, std 1.057 of the one-uniform
    pseudo Box-Muller), deterministic per seed, different across seeds.  On an all-zero input
    the features are a pure function of the noise, so compare their statistics with the oracle's."""
    layer, conf = make_layer(dither=1.0, dither_rng="device")
    layer.eval()
    x = torch.zeros(8, 16000 * 6)
    lens = [x.shape[1]] * 8
    torch.manual_seed(5)
    a, _ = layer(x.cuda(), lens)
    torch.manual_seed(5)
    b, _ = layer(x.cuda(), lens)
    torch.manual_seed(6)
    c, _ = layer(x.cuda(), lens)
    assert torch.equal(a, b) and not torch.equal(a, c)
    torch.manual_seed(0)
    ref, _ = fo.splayer_forward(x, lens, dict(conf))
    am, rm = a.cpu().mean(dim=(0, 1)), ref.mean(dim=(0, 1))
    assert (am - rm).abs().max() < 0.12, (am - rm).abs().max()  # ~4 sigma of a one-bin filter mean
    asd, rsd = a.cpu().std(dim=(0, 1)), ref.std(dim=(0, 1))
    assert ((asd - rsd).abs() / rsd).max() < 0.15


@pytest.mark.parametrize("sr,D,B,lo,hi", [(16000, 80, 32, 56000, 104000), (8000, 40, 64, 16000, 48000),
                                          (16000, 80, 16, 192000, 320000)])
def test_synthetic_shapes_vs_oracle(sr, D, B, lo, hi):
    """BASELINE configs[1..3] at a batch the oracle finishes in seconds (B/4 utterances)."""
    layer, conf = make_layer(sample_rate=sr, num_mel_bins=D)
    layer.eval()
    x, lens = fo.synth_batch(max(2, B // 4), lo, hi, sr, seed=1234)
    feats, flen = layer(x.cuda(), lens)
    ref, rlen = fo.splayer_forward(x, lens.tolist(), conf)
    ref64, _ = fo.splayer_forward(x, lens.tolist(), conf, dtype=torch.float64)
    assert torch.equal(flen.cpu(), rlen)
    close(feats, ref, ref64)


def test_cmvn_utterance_vs_fp64_oracle(wavs):
    layer, conf = make_layer(cmvn="utterance")
    layer.eval()
    x, lens = pad_batch([wavs[0], wavs[1], wavs[0][:9000]])
    feats, flen = layer(x.cuda(), lens)
    ref, rlen = fo.splayer_forward(x, lens, conf)
    assert torch.equal(flen.cpu(), rlen)
    close(feats, ref)
    for i, m in enumerate(rlen.tolist()):
        assert (feats[i, m:] == 0).all()
        assert feats[i, :m].mean(0).abs().max() < 1e-4
    layer2, conf2 = make_layer(cmvn="utterance", cmvn_norm_vars=False)
    layer2.eval()
    f2, _ = layer2(x.cuda(), lens)
    close(f2, fo.splayer_forward(x, lens, conf2)[0])


def test_specaug_rectangles_bit_exact(wavs, golden_dir):
    """SpecAug on the offline path against the reference's own spec_aug output (seed 0 / seed 3)."""
    from openasr_b200 import SPLayer
    g = np.load(os.path.join(golden_dir, "fbank_ref.npz"))
    s = np.load(os.path.join(golden_dir, "specaug_ref.npz"))
    f0, f1 = torch.from_numpy(g["w0_d80"]), torch.from_numpy(g["w1_d80"])
    pad = torch.zeros(2, 418, 80)
    pad[0, :202] = f0
    pad[1] = f1
    lens = torch.from_numpy(s["lengths"])
    for seed, key, sa in ((0, "aug_seed0", (2, 27, 2, 40)), (3, "aug2_seed3", (1, 15, 3, 300))):
        conf = {"feature_type": "offline",
                "spec_aug": dict(zip(("freq_mask_num", "freq_mask_width", "time_mask_num", "time_mask_width"), sa))}
        layer = SPLayer(conf).train()
        torch.manual_seed(seed)
        x = pad.clone().cuda()
        out, olen = layer(x, lens)
        assert out.data_ptr() == x.data_ptr()  # in place, like the reference
        ref = torch.from_numpy(s[key])
        assert torch.equal(out.cpu() != pad, ref != pad), "mask rectangles differ"
        close(out, ref)
        assert torch.equal(olen.cpu(), lens)


def test_full_training_forward_c2_like():
    """fbank + utterance CMVN + SpecAug (training) vs the oracle with the same host uniforms."""
    sa = {"freq_mask_num": 2, "freq_mask_width": 27, "time_mask_num": 2, "time_mask_width": 40}
    layer, conf = make_layer(cmvn="utterance", spec_aug=sa)
    layer.train()
    x, lens = fo.synth_batch(6, 30000, 90000, 16000, seed=9)
    torch.manual_seed(77)
    feats, flen = layer(x.cuda(), lens)
    torch.manual_seed(77)
    uni = torch.rand(8, 6)
    ref, rlen = fo.splayer_forward(x, lens.tolist(), conf, training=True, specaug_uniforms=uni)
    ref64, _ = fo.splayer_forward(x, lens.tolist(), conf, training=True, specaug_uniforms=uni, dtype=torch.float64)
    assert torch.equal(flen.cpu(), rlen)
    # CMVN divides by the per-dimension std-dev, so the log-mel tolerance scales by 1/std
    raw, _ = fo.splayer_forward(x, lens.tolist(), dict(conf, cmvn="none"), training=False)
    istd = torch.stack([1.0 / raw[i, :m].std(0, unbiased=False) for i, m in enumerate(rlen.tolist())])
    scale = istd.clamp_min(1.0)[:, None, :]
    close(feats, ref, ref64, scale=scale)
    for i, m in enumerate(rlen.tolist()):  # padding rows stay exactly zero (no mask reaches them here)
        assert (feats[i, m:] == 0).all() and (ref[i, m:] == 0).all()
    # eval mode: no SpecAug
    layer.eval()
    e, _ = layer(x.cuda(), lens)
    close(e, fo.splayer_forward(x, lens.tolist(), conf, training=False)[0],
          fo.splayer_forward(x, lens.tolist(), conf, training=False, dtype=torch.float64)[0], scale=scale)


@pytest.mark.parametrize("sr,D,B,lo,hi,tw", [(16000, 80, 32, 56000, 104000, 40), (8000, 40, 64, 16000, 48000, 40),
                                             (16000, 80, 16, 192000, 320000, 100)])
def test_full_batch_training_forward_vs_oracle(sr, D, B, lo, hi, tw):
    """BASELINE configs[1..3] at their FULL batch size: fbank + utterance CMVN + SpecAug (training) against the oracle
    (fp32 and fp64) with the same host uniforms -- every element of the batch, exact lengths, exact zero padding,
    identical mask rectangles."""
    sa = {"freq_mask_num": 2, "freq_mask_width": 27 if D == 80 else 13, "time_mask_num": 2, "time_mask_width": tw}
    layer, conf = make_layer(sample_rate=sr, num_mel_bins=D, cmvn="utterance", spec_aug=sa)
    layer.train()
    x, lens = fo.synth_batch(B, lo, hi, sr, seed=4321)
    torch.manual_seed(11)
    feats, flen = layer(x.cuda(), lens)
    torch.manual_seed(11)
    uni = torch.rand(8, B)
    ref, rlen = fo.splayer_forward(x, lens.tolist(), conf, training=True, specaug_uniforms=uni)
    ref64, _ = fo.splayer_forward(x, lens.tolist(), conf, training=True, specaug_uniforms=uni, dtype=torch.float64)
    assert torch.equal(flen.cpu(), rlen)
    raw, _ = fo.splayer_forward(x, lens.tolist(), dict(conf, cmvn="none"), training=False)
    istd = torch.stack([1.0 / raw[i, :m].std(0, unbiased=False) for i, m in enumerate(rlen.tolist())])
    close(feats, ref, ref64, scale=istd.clamp_min(1.0)[:, None, :])
    fc = feats.cpu()
    for i, m in enumerate(rlen.tolist()):
        assert torch.equal(fc[i, m:] == 0, ref[i, m:] == 0)
    # the masked cells (rows replaced by time means, columns by frequency means) are the same cells
    plain, _ = fo.splayer_forward(x, lens.tolist(), conf, training=False)
    assert torch.equal((fc - plain).abs() > 1e-2, (ref - plain).abs() > 1e-2) or \
        ((fc - plain).abs() > 1e-2).ne((ref - plain).abs() > 1e-2).sum().item() < 1e-5 * fc.numel()


def test_short_time_mask_quirk():
    """len < time_mask_width: negative starts / spill into padding follow Python slice rules."""
    sa = {"freq_mask_num": 1, "freq_mask_width": 10, "time_mask_num": 2, "time_mask_width": 100}
    layer, conf = make_layer(spec_aug=sa, num_mel_bins=40)
    layer.train()
    x, lens = fo.synth_batch(5, 4000, 30000, 16000, seed=2)
    for seed in range(6):
        torch.manual_seed(seed)
        feats, flen = layer(x.cuda(), lens)
        torch.manual_seed(seed)
        uni = torch.rand(6, 5)
        ref, _ = fo.splayer_forward(x, lens.tolist(), conf, training=True, specaug_uniforms=uni)
        close(feats, ref)


def test_errors_and_api():
    from openasr_b200 import SPLayer, WavConv
    with pytest.raises(ValueError):
        SPLayer({"feature_type": "mfcc"})
    layer, _ = make_layer()
    with pytest.raises(AssertionError):
        layer(torch.zeros(1, 399).cuda(), [399])
    with pytest.raises(RuntimeError):
        layer(torch.zeros(1, 1000), [1000])  # CPU tensor: no fallback
    assert len(layer.state_dict()) == 0
    wc = WavConv({"d_model": 8}).cuda()
    y, ly = wc(torch.randn(2, 3200).cuda(), torch.tensor([3200, 1600]).cuda())
    assert y.shape == (2, 20, 8) and ly.tolist() == [20, 10]


# --------------------------------------------------------------------------------------------- scheduling edges
def test_engines_agree_on_ragged_misaligned_batches(monkeypatch, wavs):
    """The three kernel-A engines -- tcgen05 DFT-as-GEMM (default), warp-pipelined FFT, simple one-tile-per-CTA --
    on ragged batches, misaligned rows (every 16-byte shift) and storage offsets."""
    x, lens = fo.synth_batch(9, 500, 70000, 16000, seed=4)
    x = torch.cat([x, torch.zeros(9, 3)], dim=1)          # odd row pitch -> unaligned rows
    outs = {}
    for mode in ("umma", "fft", "simple"):
        monkeypatch.setenv("SPL_ENGINE", mode)
        layer, conf = make_layer()
        layer.eval()
        assert layer._handle(torch.device("cuda", 0)).engine_name() == mode
        xc = x.cuda()
        outs[mode] = [layer(xc, lens)[0]] + [layer(xc[:, k:], (lens - k).clamp_min(400))[0] for k in (1, 2, 3)]
        assert layer._handle(torch.device("cuda", 0)).debug_status() == 0
    for other in ("umma", "fft"):  # different formulation / rounding order only (ill-conditioned synthetic elements
        for a, b in zip(outs[other], outs["simple"]):  # move by a few 1e-3, like the fp32 reference itself does vs fp64)
            assert (a - b).abs().max().item() < 8e-3 and (a - b).abs().mean().item() < 2e-5
    assert torch.equal(outs["umma"][0] == 0, outs["simple"][0] == 0)
    ref, _ = fo.splayer_forward(x, lens.tolist(), conf)
    ref64 = fo.splayer_forward(x, lens.tolist(), conf, dtype=torch.float64)[0]
    close(outs["umma"][0], ref, ref64)
    close(outs["fft"][0], ref, ref64)
    # use_energy: the tcgen05 engine hands the configuration to the FFT engine
    monkeypatch.setenv("SPL_ENGINE", "umma")
    layer, conf = make_layer(use_energy=True)
    layer.eval()
    assert layer._handle(torch.device("cuda", 0)).engine_name() == "fft"
    close(layer(x.cuda(), lens)[0], fo.splayer_forward(x, lens.tolist(), conf)[0],
          fo.splayer_forward(x, lens.tolist(), conf, dtype=torch.float64)[0])


def test_many_short_utterances_and_large_batches():
    """Groups / chunks spanning many 1-3 frame utterances; B = 512 (in-kernel scheduling) and
    B = 600 (simple-kernel fallback)."""
    for B in (512, 600):
        g = torch.Generator().manual_seed(B)
        lens = torch.randint(400, 900, (B,), generator=g)
        x = (1000 * torch.randn(B, 900, generator=g)).round()
        x = x * (torch.arange(900)[None, :] < lens[:, None])
        sub = list(range(0, B, 37))
        layer, conf = make_layer(num_mel_bins=40)
        layer.eval()
        feats, flen = layer(x.cuda(), lens)
        ref, rlen = fo.splayer_forward(x[sub], lens[sub].tolist(), conf)
        ref64, _ = fo.splayer_forward(x[sub], lens[sub].tolist(), conf, dtype=torch.float64)
        assert torch.equal(flen.cpu()[sub], rlen)
        close(feats.cpu()[sub][:, :ref.shape[1]], ref, ref64)
        assert (feats.cpu()[sub][:, ref.shape[1]:] == 0).all()
        # utterance CMVN on tiny utterances: a 1-3 frame utterance has a near-zero variance, so the
        # normalised values amplify the fbank's own fp32 noise without bound; isolate kernel B by
        # applying the fp64 oracle CMVN to the GPU's own un-normalised features
        layer2, conf2 = make_layer(num_mel_bins=40, cmvn="utterance")
        layer2.eval()
        norm, _ = layer2(x.cuda(), lens)
        norm = norm.cpu()
        assert torch.isfinite(norm).all()
        raw_gpu = feats.cpu()
        want = fo.cmvn_apply(raw_gpu[sub].double(), rlen.tolist(), "utterance").float()
        for j, i in enumerate(sub):
            m = rlen[j].item()
            assert (norm[i, m:] == 0).all()
            if m >= 3:
                sd = raw_gpu[i, :m].double().std(0, unbiased=False)
                ok = sd > 0.05  # columns whose variance is not dominated by rounding noise
                assert ((norm[i, :m] - want[j, :m]).abs()[:, ok] < 2e-3).all()
            elif m == 1:
                assert norm[i, :m].abs().max().item() < 1e-3


def test_global_cmvn_single_gpu(wavs):
    """cmvn='global': kernel A's fused fp64 statistics pass + kernel B against the fp64 oracle
    (the cross-GPU all-reduce is a no-op at world size 1; the gloo test covers N = 2)."""
    from openasr_b200.cmvn import GlobalCmvn
    layer, conf = make_layer(cmvn="global")
    layer.eval()
    x, lens = pad_batch([wavs[0], wavs[1], wavs[0][:20000]])
    acc = GlobalCmvn(layer, torch.device("cuda"))
    acc.update(x.cuda(), lens)
    acc.update(x[:2].cuda(), lens[:2])          # statistics accumulate across calls
    mean, istd = acc.finalize()
    raw, rlen = fo.splayer_forward(x, lens, dict(conf, cmvn="none"))
    st = fo.cmvn_stats(raw, rlen.tolist()) + fo.cmvn_stats(raw[:2], rlen[:2].tolist())
    assert abs(acc.stats[-1].item() - st[2, 0].item()) < 0.5
    assert torch.allclose(mean.cpu(), st[0] / st[2], atol=2e-5)
    feats, flen = layer(x.cuda(), lens)
    ref, _ = fo.splayer_forward(x, lens, conf, global_stats=st)
    close(feats, ref, scale=(istd.cpu().float().clamp_min(1.0))[None, None, :])
    assert len(layer.state_dict()) == 0


def test_concurrent_host_threads_same_device(wavs):
    """DataParallel calls the module from one Python thread per GPU (train.py:134); here several
    threads share ONE device and one cached handle, each on its own stream: results must be identical
    to the serial ones."""
    import threading
    layer, conf = make_layer(cmvn="utterance")
    layer.eval()
    batches = []
    for s in range(4):
        x, lens = fo.synth_batch(5, 8000, 60000, 16000, seed=40 + s)
        batches.append((x.cuda(), lens))
    serial = [layer(x, l)[0].clone() for x, l in batches]
    out = [None] * 4

    def work(i):
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for _ in range(10):
                out[i] = layer(batches[i][0], batches[i][1])[0]
        st.synchronize()

    ths = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for a, b in zip(serial, out):
        assert torch.equal(a, b)


# --------------------------------------------------------------------------------------------- full-size properties
@pytest.mark.parametrize("sr,D,B,lo,hi", [(16000, 80, 32, 56000, 104000), (8000, 40, 64, 16000, 48000)])
def test_full_size_invariances(sr, D, B, lo, hi):
    """Size-independent properties at the BASELINE shapes (no oracle needed):
    * a frame's features do not depend on where the utterance sits in the batch (bit-exact under a
      batch permutation and against the utterance processed alone);
    * time-shift equivariance: dropping the first 4 hops of samples drops the first 4 frames, every
      other frame keeps its slot in its 4-frame group, so the values are bit-identical;
    * dropping 1 hop moves frames to the other half of their packed FFT: equal to rounding error;
    * padding rows are exact zeros and lengths exact at full size."""
    x, lens = fo.synth_batch(B, lo, hi, sr, seed=11)
    layer, conf = make_layer(sample_rate=sr, num_mel_bins=D)
    layer.eval()
    xc = x.cuda()
    feats, flen = layer(xc, lens)
    shift, win = layer._handle(xc.device).shift, layer._handle(xc.device).win
    frames = 1 + (lens - win) // shift
    assert torch.equal(flen.cpu(), frames)
    for i in range(B):
        assert (feats[i, frames[i]:] == 0).all()
    # batch permutation
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(3))
    fp, lp = layer(xc[perm.cuda()], lens[perm])
    assert torch.equal(lp.cpu(), frames[perm])
    assert torch.equal(fp, feats[perm.cuda()])
    # one utterance alone
    for i in (0, B // 2, B - 1):
        n = int(lens[i])
        fa, la = layer(xc[i:i + 1, :n].contiguous(), [n])
        assert int(la[0]) == int(frames[i])
        assert torch.equal(fa[0], feats[i, :int(frames[i])])
    # shift by one group of 4 frames: bit-exact
    f4, l4 = layer(xc[:, 4 * shift:].contiguous(), lens - 4 * shift)
    assert torch.equal(l4.cpu(), frames - 4)
    for i in range(B):
        m = int(frames[i]) - 4
        assert torch.equal(f4[i, :m], feats[i, 4:4 + m])
    # shift by one frame: other half of the complex FFT, rounding-level agreement only
    f1, l1 = layer(xc[:, shift:].contiguous(), lens - shift)
    assert torch.equal(l1.cpu(), frames - 1)
    diffs = []
    for i in range(B):
        m = int(frames[i]) - 1
        diffs.append((f1[i, :m] - feats[i, 1:1 + m]).abs().flatten())
    d = torch.cat(diffs)
    # fp32 rounding only; the handful of ill-conditioned elements reach a few 1e-3
    assert d.max().item() < 1e-2 and d.mean().item() < 1e-5 and (d > 1e-3).float().mean().item() < 1e-4, \
        (d.max().item(), d.mean().item())


def test_pinned_collator_int16_and_fp32(wavs):
    """Row f4: pad into pinned staging, asynchronous H2D on a side stream, then SPLayer.forward --
    identical features for the int16 and the fp32 staging paths and for the plain padded batch."""
    from openasr_b200.batching import PinnedWaveCollator
    layer, conf = make_layer()
    layer.eval()
    ws16 = [w.numpy().astype(np.int16) for w in (wavs[0], wavs[1], wavs[0][:30000])]
    x, lens = pad_batch([torch.from_numpy(w.astype(np.float32)) for w in ws16])
    ref, rlen = layer(x.cuda(), lens)
    for int16 in (True, False):
        col = PinnedWaveCollator("cuda:0", max_batch=4, max_len=max(lens) + 100, slots=2, int16=int16)
        for _ in range(3):  # cycles through the slots (reuse after completion)
            dev, l, ev = col(ws16)
            col.wait(ev, dev)
            f, fl = layer(dev, l)
            assert dev.dtype == (torch.int16 if int16 else torch.float32)
            assert torch.equal(fl, rlen) and torch.equal(f, ref)


@pytest.mark.parametrize("seed", list(range(10)))
def test_fuzz_configs_vs_oracle(seed):
    """Randomised geometry: batch size, ragged lengths down to a single window (1-, 2-, 3-frame groups and
    utterances), odd row pitch and storage offset (unaligned TMA heads), sample rate, mel count, energy
    column, int16 / fp32 ingest -- each against the CPU oracle, dither 0."""
    g = torch.Generator().manual_seed(1000 + seed)
    sr = [16000, 8000][int(torch.randint(0, 2, (1,), generator=g))]
    win = 400 if sr == 16000 else 200
    D = [23, 40, 64, 80, 128][int(torch.randint(0, 5, (1,), generator=g))]
    B = int(torch.randint(1, 24, (1,), generator=g))
    use_energy = bool(torch.randint(0, 2, (1,), generator=g))
    as_int16 = bool(torch.randint(0, 2, (1,), generator=g))
    hi = [win + 3, win + 700, 9000, 30000][int(torch.randint(0, 4, (1,), generator=g))]
    lens = torch.randint(win, hi + 1, (B,), generator=g)
    pitch = int(lens.max()) + int(torch.randint(0, 7, (1,), generator=g))
    off = int(torch.randint(0, 4, (1,), generator=g))
    x = (1500.0 * torch.randn(B, pitch, generator=g)).round().clamp(-32768, 32767)
    x = x * (torch.arange(pitch)[None, :] < lens[:, None])
    layer, conf = make_layer(sample_rate=sr, num_mel_bins=D, use_energy=use_energy)
    layer.eval()
    store = torch.zeros(B * pitch + off)
    store[off:] = x.flatten()
    xd = store.cuda()[off:].view(B, pitch)  # storage offset: rows start at 4-byte, not 16-byte, boundaries
    if as_int16:
        xd = xd.to(torch.int16)
    feats, flen = layer(xd, lens)
    ref, rlen = fo.splayer_forward(x, lens.tolist(), conf)
    ref64, _ = fo.splayer_forward(x, lens.tolist(), conf, dtype=torch.float64)
    assert torch.equal(flen.cpu(), rlen)
    assert feats.shape == ref.shape
    close(feats, ref, ref64)
    for i in range(B):
        assert (feats[i, int(rlen[i]):] == 0).all()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_dataparallel_two_gpus(wavs):
    """train.py:134 wraps the model in torch.nn.DataParallel: SPLayer replicas run concurrently in per-GPU
    threads (shared Python attributes, one handle per device)."""
    from openasr_b200 import SPLayer

    class Model(torch.nn.Module):
        def __init__(self, conf):
            super().__init__()
            self.splayer = SPLayer(conf)

        def forward(self, wav, lens):
            feats, flen = self.splayer(wav, lens)
            return feats, flen

    conf = {"feature_type": "fbank", "sample_rate": 16000, "num_mel_bins": 80, "use_energy": False, "dither": 0.0,
            "cmvn": "utterance"}
    ws = [wavs[0], wavs[1][:wavs[0].shape[0]], wavs[0].flip(0), wavs[1][1000:1000 + wavs[0].shape[0]]]
    x, lens = pad_batch(ws)  # equal lengths: DataParallel's gather needs the same T on every replica
    model = torch.nn.DataParallel(Model(conf).cuda().eval(), device_ids=[0, 1])
    single = SPLayer(conf).cuda().eval()
    ref, rlen = single(x.cuda(), lens)
    for _ in range(5):
        feats, flen = model(x.cuda(), torch.tensor(lens).cuda())
        assert torch.equal(flen.cpu(), rlen.cpu())
        assert torch.equal(feats.cpu(), ref.cpu())


@pytest.mark.parametrize("sr", [6000, 10000, 11025, 12000, 20000])
def test_other_sample_rates_generic_window(sr, monkeypatch):
    """Sample rates whose window is not the 400 / 200-sample special case run the generic-window
    instantiations (dynamic row validity, 16 un-pruned FFT rows) of every kernel-A engine."""
    S, Nw, Nfft = fo.frame_params(float(sr))
    assert Nfft in (256, 512)
    g = torch.Generator().manual_seed(sr)
    B = 7
    lens = torch.randint(Nw, 6 * sr // 10, (B,), generator=g)
    lens[0] = Nw  # exactly one frame
    x = (1800.0 * torch.randn(B, int(lens.max()), generator=g)).round()
    x = x * (torch.arange(x.shape[1])[None, :] < lens[:, None])
    outs = {}
    for mode in ("umma", "fft", "simple"):
        monkeypatch.setenv("SPL_ENGINE", mode)
        layer, conf = make_layer(sample_rate=sr, num_mel_bins=40)
        layer.eval()
        outs[mode], flen = layer(x.cuda(), lens)
    ref, rlen = fo.splayer_forward(x, lens.tolist(), conf)
    ref64, _ = fo.splayer_forward(x, lens.tolist(), conf, dtype=torch.float64)
    assert torch.equal(flen.cpu(), rlen)
    for mode in outs:
        close(outs[mode], ref, ref64)
    # host-stream dither (parity mode) on the generic path
    monkeypatch.delenv("SPL_ENGINE")
    layer, conf = make_layer(sample_rate=sr, num_mel_bins=40, dither=1.0, dither_rng="host")
    layer.eval()
    torch.manual_seed(5)
    f, _ = layer(x[:3].cuda(), lens[:3])
    torch.manual_seed(5)
    r, _ = fo.splayer_forward(x[:3], lens[:3].tolist(), conf)
    close(f, r)
