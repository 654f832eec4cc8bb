"""Soak run (GPU box): many forward calls with random geometry and all modes interleaved; every 10th
call is checked against the CPU oracle.  Prints a summary; non-zero exit on any mismatch."""
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
t0 = time.time()
mem0 = None
for it in range(N):
    sr = [16000, 8000][int(torch.randint(0, 2, (1,), generator=g))]
    win = 400 if sr == 16000 else 200
    D = [40, 80][int(torch.randint(0, 2, (1,), generator=g))]
    cmvn = ["none", "utterance"][int(torch.randint(0, 2, (1,), generator=g))]
    train = bool(torch.randint(0, 2, (1,), generator=g))
    key = (sr, D, cmvn)
    if key not in layers:
        conf = {"feature_type": "fbank", "sample_rate": sr, "num_mel_bins": D, "use_energy": False, "dither": 0.0,
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
    check = (it % 10 == 0) and not train
    feats, flen = layer(xd, lens)
    assert torch.isfinite(feats).all()
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
print("soak ok: %d calls, %d checked vs oracle, worst |d| %.2e, %.1f s, reserved memory %d -> %d MB" %
      (N, checked, worst, time.time() - t0, (mem0 or 0) >> 20, torch.cuda.memory_reserved() >> 20))
