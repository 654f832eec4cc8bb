"""Per-warp timeline of the warp-pipelined kernel A (diagnostic build: make -C openasr_b200/csrc trace).

Prints, in SM clock cycles relative to each warp's kernel entry: prologue end, and for every group
iteration the TMA wait, stage 1, stage 2 + power, mel and store phases; then the distribution of
per-warp end times, which shows the tail the kernel's latency is made of.
"""
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

dither = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
WPC = 8 if os.environ.get("SPL_CTAS_PER_SM") == "1" else 16  # warps per CTA of the variant being traced
NCTA = 148 if True else 296
conf = {"feature_type": "fbank", "sample_rate": 16000, "num_mel_bins": 80, "use_energy": False, "dither": dither,
        "cmvn": "utterance"}
layer = SPLayer(conf).cuda().eval()
x, lens = synth_batch(32, 56000, 104000, 16000, seed=1234)
xc = x.cuda()
for _ in range(6):
    layer(xc, lens)
torch.cuda.synchronize()
lib = cap.load()
n = 296 * 8 * 32
buf = (ctypes.c_ulonglong * n)()
lib.spl_debug_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.spl_debug_trace(buf, n) == 0
t = np.frombuffer(buf, dtype=np.uint64)[:NCTA * WPC * 32].reshape(NCTA, WPC, 32).astype(np.int64)
gt = t[:, :, 1]
print("globaltimer spread of warp starts (ns): min %d max %d" % (0, int(gt.max() - gt.min())))
rel = t - t[:, :, 0:1]
print("prologue: tables wait begins %.0f, ends %.0f cycles (median)" % (np.median(rel[:, :, 2]), np.median(rel[:, :, 3])))
print("prologue detail (median cycles): own setup done %.0f, prefix barrier %.0f, shares barrier %.0f, first TMA issued %.0f, zero-pad done %.0f" %
      tuple(np.median(rel[:, :, k]) for k in (22, 23, 24, 25, 2)))
print("  warp 0 only: setup done %.0f" % np.median(rel[:, 0, 22]))
names = ["wait", "stage1", "stage2", "mel", "store"]
for it in range(4):
    base = 4 + 6 * it
    ok = (t[:, :, base] > t[:, :, 0]) & (t[:, :, base + 5] > t[:, :, base])
    ok &= (rel[:, :, base] < rel[:, :, 30])
    if it > 0:
        ok &= t[:, :, base] >= t[:, :, base - 1]
    if ok.sum() == 0:
        break
    d = np.diff(t[:, :, base:base + 6], axis=2)[ok]
    print("iter %d: %4d warps, starts at %6.0f; " % (it, ok.sum(), np.median(rel[:, :, base][ok])) +
          ", ".join("%s %5.0f" % (nm, np.median(d[:, i])) for i, nm in enumerate(names)) +
          " | total %5.0f (p90 %5.0f)" % (np.median(d.sum(1)), np.percentile(d.sum(1), 90)))
end = rel[:, :, 30]
print("loop end per warp: median %.0f  p10 %.0f  p90 %.0f  max %.0f cycles" %
      (np.median(end), np.percentile(end, 10), np.percentile(end, 90), end.max()))
print("kernel end (slot 31): median %.0f max %.0f" % (np.median(rel[:, :, 31]), rel[:, :, 31].max()))
cta_end = rel[:, :, 31].max(axis=1)
print("per-CTA end: min %.0f median %.0f max %.0f" % (cta_end.min(), np.median(cta_end), cta_end.max()))
its = np.zeros((NCTA, WPC), dtype=int)
for it in range(4):
    base = 4 + 6 * it
    its += ((t[:, :, base] > t[:, :, 0]) & (rel[:, :, base] < rel[:, :, 30])).astype(int)
print("groups per warp histogram:", np.bincount(its.ravel()))

# per-SM view: which CTAs share an SM, when each SM finishes, and how many groups it ran
smid = t[:, 0, 26]
gt0 = t[:, :, 1].min()
cta_start_ns = t[:, :, 1].min(axis=1) - gt0
cta_groups = its.sum(axis=1)
sm_end = {}
for c in range(NCTA):
    sm_end.setdefault(int(smid[c]), []).append((int(cta_end[c]), int(cta_groups[c]), int(cta_start_ns[c]), c))
ends = sorted(((max(e for e, _, _, _ in v), sum(g for _, g, _, _ in v), len(v), [x[3] for x in v]) for v in sm_end.values()))
print("SMs used: %d; CTAs per SM histogram: %s" % (len(sm_end), np.bincount([len(v) for v in sm_end.values()])))
print("fastest SMs (end cycles, groups, ctas):", ends[:4])
print("slowest SMs (end cycles, groups, ctas):", ends[-6:])
print("late-starting CTAs (start ns > 1000):", [(c, int(cta_start_ns[c]), int(smid[c])) for c in range(NCTA) if cta_start_ns[c] > 1000][:20])
g_by_sm = np.array([e[1] for e in ends]); e_by_sm = np.array([e[0] for e in ends])
for g in sorted(set(g_by_sm)):
    print("  SMs with %d groups: %d, median end %.0f" % (g, (g_by_sm == g).sum(), np.median(e_by_sm[g_by_sm == g])))

# spread of the first iteration: by phase, by warp index, by co-resident CTA
base = 4
d0 = np.diff(t[:, :, base:base + 6], axis=2)  # [cta, warp, phase]
print("iteration 0 phase percentiles (p50 / p90 / p99):")
for i, nm in enumerate(names):
    v = d0[:, :, i].ravel()
    print("  %-7s %6.0f %6.0f %6.0f" % (nm, np.percentile(v, 50), np.percentile(v, 90), np.percentile(v, 99)))
tot0 = d0.sum(axis=2)
print("iteration 0 total by warp index (median):", [int(np.median(tot0[:, w])) for w in range(WPC)])
print("iteration 0 start by warp index (median):", [int(np.median(rel[:, w, base])) for w in range(WPC)])
print("iteration 0 total, warps 0..7 vs 8..15 of the CTA (median): %.0f vs %.0f" % (np.median(tot0[:, :WPC // 2]), np.median(tot0[:, WPC // 2:])))
per_sm = {}
for c in range(NCTA):
    per_sm.setdefault(int(smid[c]), []).append(float(np.median(tot0[c])))
sm_med = np.array([np.mean(v) for v in per_sm.values()])
print("iteration 0 per-SM median total: p10 %.0f p50 %.0f p90 %.0f max %.0f" %
      (np.percentile(sm_med, 10), np.percentile(sm_med, 50), np.percentile(sm_med, 90), sm_med.max()))
gpc = np.array([k // 18 for k in per_sm.keys()])
for gidx in sorted(set(gpc)):
    sel = gpc == gidx
    print("  smid %3d..%3d: median iteration-0 total %.0f, SM end %.0f" %
          (gidx * 18, gidx * 18 + 17, np.median(sm_med[sel]),
           np.median([max(e for e, _, _, _ in sm_end[k]) for k in per_sm.keys() if k // 18 == gidx])))
