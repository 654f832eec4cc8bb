"""Smallest run that exercises every kernel of the library once, for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_case.py
    compute-sanitizer --tool racecheck python tools/sanitize_case.py
The fixture batch (two real utterances + a truncated one, odd row pitch) through both kernel-A engines (FFT default,
tcgen05), kernel B with utterance CMVN + SpecAug, the multi-batch entry, int16 ingest, conv0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from openasr_b200 import SPLayer
from openasr_b200.blocks.conv_layers import Conv2dSubsampleV2

G = os.path.join(ROOT, "tests", "golden")
ws = [torch.from_numpy(np.load(os.path.join(G, "wav%d.npy" % i)).astype(np.float32)) for i in (0, 1)]
ws = [ws[0][:16000], ws[1][:24000], ws[0][:7001]]
lens = [w.shape[0] for w in ws]
x = torch.zeros(len(ws), max(lens) + 1)
for i, w in enumerate(ws):
    x[i, :lens[i]] = w
sa = {"freq_mask_num": 2, "freq_mask_width": 27, "time_mask_num": 2, "time_mask_width": 40}
for eng in ("fft", "umma"):
    os.environ["SPL_ENGINE"] = eng
    layer = SPLayer({"feature_type": "fbank", "sample_rate": 16000, "num_mel_bins": 80, "use_energy": False, "dither": 1.0,
                     "cmvn": "utterance", "spec_aug": sa}).cuda().train()
    h = layer._handle(torch.device("cuda", 0))
    f, fl = layer(x.cuda(), lens)
    outs = layer.forward_multi([(x.cuda(), lens), (x[:2, :24000].contiguous().cuda(), lens[:2])])
    f16, _ = layer(x.to(torch.int16).cuda(), lens)
    torch.cuda.synchronize()
    print(eng, h.engine_name(), "status 0x%x" % h.debug_status(), tuple(f.shape), fl.tolist(), bool(torch.isfinite(f).all()),
          bool(torch.isfinite(f16).all()), [tuple(o[0].shape) for o in outs], flush=True)
conv = Conv2dSubsampleV2(80, 64, 2).cuda().eval()
with torch.no_grad():
    y, ly = conv(f, fl)
torch.cuda.synchronize()
print("conv0", tuple(y.shape), flush=True)
