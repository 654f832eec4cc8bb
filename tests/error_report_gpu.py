"""Diagnostic (GPU box): error of the CUDA path vs the fp32 oracle and vs the fp64 oracle."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import frontend_oracle as fo
from openasr_b200 import SPLayer

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
wavs = [torch.from_numpy(np.load(os.path.join(G, "wav%d.npy" % i)).astype(np.float32)) for i in (0, 1)]
cases = [("wav0", wavs[0], 16000, 80), ("wav1", wavs[1], 16000, 80), ("wav1@8k", wavs[1][::2].contiguous(), 8000, 40)]
x, lens = fo.synth_batch(2, 60000, 90000, 16000, seed=1)
cases.append(("synth", x[0, :lens[0]], 16000, 80))
for name, w, sr, D in cases:
    conf = {"feature_type": "fbank", "sample_rate": sr, "num_mel_bins": D, "use_energy": False, "dither": 0.0}
    layer = SPLayer(conf).cuda().eval()
    g, _ = layer(w.view(1, -1).cuda(), [w.shape[0]])
    g = g[0].cpu().double()
    o32 = fo.fbank(w, sample_rate=sr, num_mel_bins=D, dither=0.0).double()
    o64 = fo.fbank(w, sample_rate=sr, num_mel_bins=D, dither=0.0, dtype=torch.float64)
    print("%-8s gpu-vs-ref32 max %.2e mean %.2e | gpu-vs-fp64 max %.2e mean %.2e | ref32-vs-fp64 max %.2e mean %.2e" % (
        name, (g - o32).abs().max(), (g - o32).abs().mean(), (g - o64).abs().max(), (g - o64).abs().mean(),
        (o32 - o64).abs().max(), (o32 - o64).abs().mean()))
