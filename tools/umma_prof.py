"""A handful of multi-batch launches of kernel A on the AISHELL shape (for ncu / timing of one engine).
    python tools/umma_prof.py [K] [dither] [engine] [launches]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dither = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
os.environ["SPL_ENGINE"] = sys.argv[3] if len(sys.argv) > 3 else "umma"
n_launch = int(sys.argv[4]) if len(sys.argv) > 4 else 6
import torch
from openasr_b200 import SPLayer, tables
from openasr_b200.synth import synth_batch
dev = torch.device("cuda", 0)
layer = SPLayer({"feature_type": "fbank", "sample_rate": 16000, "num_mel_bins": 80, "use_energy": False, "dither": dither}).cuda().eval()
h = layer._handle(dev)
items = []
for k in range(K):
    wav, lens = synth_batch(32, 56000, 104000, 16000, seed=1234 + k)
    frames = [tables.frame_count(int(n), h.win, h.shift) for n in lens.tolist()]
    T = max(frames)
    items.append({"wav": wav.to(dev), "lens": lens.to(dev), "T": T, "feats": torch.empty((32, T, 80), device=dev),
                  "flen": torch.zeros(32, dtype=torch.int64, device=dev), "stats": torch.empty((32, 2, 80), dtype=torch.float64, device=dev)})
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(n_launch):
    e0.record()
    h.fbank_multi(items, dither_seed=7 + i)
    e1.record()
    torch.cuda.synchronize()
    print("launch %d: %.1f us for %d batches (%s, dither %g) status 0x%x" % (i, 1e3 * e0.elapsed_time(e1), K, h.engine_name(), dither, h.debug_status()), flush=True)
