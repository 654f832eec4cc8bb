"""Soak run (GPU box): many forward / forward_multi calls with random geometry and all modes interleaved (sample rate,
mel bins, CMVN, SpecAug, int16 ingest, device dither, 1-3 batches per call); every 10th call is checked against
the CPU oracle and the engine's device status word must stay 0.  SPL_ENGINE selects the kernel-A engine.
Prints a summary; non-zero exit on any mismatch.  (compute-sanitizer is closed on the GPU pool: this is the
long-run evidence that the kernels neither hang nor corrupt their neighbours' buffers.)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from openasr_b200 import SPLayer
from oracle import frontend_oracle as fo

N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
g = torch.Generator().manual_seed(77)
layers = {}
worst = 0.0
checked = 0
multi_calls = 0
worst_multi = 0.0
ENGINE = os.environ.get("SPL_ENGINE", "fft")
t0 = time.time()
mem0 = None
for it in range(N):
    sr = [16000, 8000][int(torch.randint(0, 2, (1,), generator=g))]
    win = 400 if sr == 16000 else 200
    D = [40, 80][int(torch.randint(0, 2, (1,), generator=g))]
    cmvn = ["none", "utterance"][int(torch.randint(0, 2, (1,), generator=g))]
    train = bool(torch.randint(0, 2, (1,), generator=g))
    dith = [0.0, 1.0][int(torch.randint(0, 2, (1,), generator=g))]
    key = (sr, D, cmvn, dith)
    if key not in layers:
        conf = {"feature_type": "fbank", "sample_rate": sr, "num_mel_bins": D, "use_energy": False, "dither": dith,
                "cmvn": cmvn, "spec_aug": {"freq_mask_num": 2, "freq_mask_width": 10, "time_mask_num": 2, "time_mask_width": 20}}
        layers[key] = (SPLayer(conf).cuda(), conf)
    layer, conf = layers[key]
    layer.train(train)
    B = int(torch.randint(1, 48, (1,), generator=g))
    hi = [win + 50, 5000, 60000][int(torch.randint(0, 3, (1,), generator=g))]
    lens = torch.randint(win, hi + 1, (B,), generator=g)
    L = int(lens.max())
    x = (2000.0 * torch.randn(B, L, generator=g)).round()
    x = x * (torch.arange(L)[None, :] < lens[:, None])
    xd = x.cuda()
    if it % 3 == 0:
        xd = xd.to(torch.int16)
    check = (it % 10 == 0) and not train and dith == 0.0
    nb = 1 + int(torch.randint(0, 3, (1,), generator=g))
    if nb == 1:
        feats, flen = layer(xd, lens)
    else:  # the same batch nb times in one library call: every copy must equal the single call (dither off)
        outs = layer.forward_multi([(xd, lens)] * nb)
        feats, flen = outs[0]
        if dith == 0.0 and not train:
            for f2, l2 in outs[1:]:
                dm = (f2 - feats).abs()
                assert torch.equal(l2, flen)
                if ENGINE == "fft":   # batch position does not enter the FFT engine's arithmetic
                    assert torch.equal(f2, feats)
                else:                 # tcgen05: scale / pivot per 8 frames of the flattened list (tests/test_gpu_multi.py)
                    assert dm.max().item() < 3e-2 and dm.mean().item() < 1e-4, (it, dm.max().item(), dm.mean().item())
                worst_multi = max(worst_multi, dm.max().item())
        multi_calls += 1
    assert torch.isfinite(feats).all()
    for i, m in enumerate(flen.tolist()):
        if not train:
            assert (feats[i, m:] == 0).all()
    if check:
        ref, rlen = fo.splayer_forward(x, lens.tolist(), conf, training=False)
        assert torch.equal(flen.cpu(), rlen)
        if cmvn == "none":
            d = (feats.cpu() - ref).abs()
            tol = 1e-3 + 1e-4 * ref.abs()
            frac_bad = (d > tol).float().mean().item()
            assert frac_bad < 1e-3 and d.max().item() < 2e-2, (it, frac_bad, d.max().item())
            worst = max(worst, d.max().item())
        checked += 1
    if it == 50:
        torch.cuda.synchronize()
        mem0 = torch.cuda.memory_reserved()
torch.cuda.synchronize()
status = [l._handle(torch.device("cuda", 0)).debug_status() for l, _ in layers.values()]
assert all(st == 0 for st in status), status
engines = sorted({l._handle(torch.device("cuda", 0)).engine_name() for l, _ in layers.values()})
print("soak ok [%s]: %d calls (%d forward_multi, worst copy-to-copy |d| %.2e), %d checked vs oracle, worst |d| %.2e, "
      "device status 0, %.1f s, reserved memory %d -> %d MB" % ("/".join(engines), N, multi_calls, worst_multi, checked, worst, time.time() - t0,
                                       (mem0 or 0) >> 20, torch.cuda.memory_reserved() >> 20))
