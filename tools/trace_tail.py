"""Tail of a multi-batch launch of kernel A (diagnostic build: make -C openasr_b200/csrc trace): when each warp leaves
the group loop and when each CTA ends, relative to the earliest CTA start -- how much of the launch is idle tail."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import openasr_b200._capi as cap

cap.LIB_PATH = os.path.join(os.path.dirname(cap.LIB_PATH), "libspl_b200_trace.so")
import torch
from openasr_b200 import SPLayer
from openasr_b200.synth import synth_batch

dither = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
conf = {"feature_type": "fbank", "sample_rate": 16000, "num_mel_bins": 80, "use_energy": False, "dither": dither, "cmvn": "utterance"}
layer = SPLayer(conf).cuda().eval()
batches = []
for k in range(K):
    x, lens = synth_batch(32, 56000, 104000, 16000, seed=1234 + k)
    batches.append((x.cuda(), lens))
for _ in range(4):
    layer.forward_multi(batches)
torch.cuda.synchronize()
lib = cap.load()
NCTA, WPC = 148, 16
n = 296 * 8 * 32
buf = (ctypes.c_ulonglong * n)()
lib.spl_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.spl_debug_trace(buf, n) == 0
t = np.frombuffer(buf, dtype=np.uint64)[:NCTA * WPC * 32].reshape(NCTA, WPC, 32).astype(np.int64)
gt = t[:, :, 1]                       # globaltimer (ns) at warp start
start_ns = gt - gt.min()
print("warp starts spread over %.1f us" % (start_ns.max() / 1e3))
rel = t - t[:, :, 0:1]                # cycles since the warp's own start
loop_end, kern_end = rel[:, :, 30], rel[:, :, 31]
cyc = lambda a: "min %.0f p10 %.0f median %.0f p90 %.0f max %.0f" % (a.min(), np.percentile(a, 10), np.median(a), np.percentile(a, 90), a.max())
print("loop start (tables landed), cycles:", cyc(rel[:, :, 3]))
print("per-warp loop end, cycles:  ", cyc(loop_end))
print("per-warp kernel end, cycles:", cyc(kern_end))
cta_end = kern_end.max(axis=1)
print("per-CTA end, cycles:        ", cyc(cta_end))
total = cta_end.max()
busy = (loop_end - rel[:, :, 3]).mean()
print("launch = %.0f cycles (%.1f us at 1.965 GHz); mean warp busy in the loop %.0f cycles = %.3f of the launch" % (total, total / 1965.0, busy, busy / total))
print("  prologue share %.3f, tail after the warp's last group %.3f (intra-CTA %.3f + inter-CTA %.3f)" % (
    rel[:, :, 3].mean() / total, (total - loop_end).mean() / total,
    (cta_end[:, None] - loop_end).mean() / total, (total - cta_end).mean() / total))
order = np.argsort(cta_end)
np.save(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "trace_tail_K%d_%s.npy" % (K, os.environ.get("TAG", "a"))),
        np.stack([np.arange(NCTA), cta_end, loop_end.mean(axis=1)]))
print("correlation of CTA end with blockIdx: %.3f" % np.corrcoef(np.arange(NCTA), cta_end)[0, 1])
print("fastest CTAs (blockIdx, end):", [(int(c), int(cta_end[c])) for c in order[:5]])
print("slowest CTAs (blockIdx, end):", [(int(c), int(cta_end[c])) for c in order[-5:]])
