"""Per-warp clock64 timeline of CTA 0 of the tcgen05 engine (diagnostic build: make -C openasr_b200/csrc trace).
    python tools/trace_umma.py [K] [dither]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dither = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
os.environ["SPL_ENGINE"] = "umma"
import numpy as np, torch
from openasr_b200 import _capi
_capi.LIB_PATH = os.path.join(ROOT, "openasr_b200", "lib", "libspl_b200_trace.so")
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
for i in range(3):
    h.fbank_multi(items, dither_seed=7 + i)
torch.cuda.synchronize()
lib = _capi.load()
lib.spl_debug_utrace.argtypes = [C.c_void_p, C.c_int]
buf = np.zeros(23 * 512, np.uint64)
assert lib.spl_debug_utrace(buf.ctypes.data, buf.size) == 0
buf = buf.reshape(23, 512)
t00 = min(int(buf[w, 0]) & 0xffffffffffff for w in range(23))
names = {0: "start", 1: "P ready", 2: "P samples", 3: "P pass1", 40: "M tmem-empty", 100: "S built", 101: "S samples-free",
         102: "S tma-issued", 103: "S published", 110: "E ready", 111: "E tmem-full", 130: "E tmem-released", 131: "E tile-end", 200: "done"}
for w in (0, 12, 16):
    n = int(buf[w, 511])
    print("---- warp %d (%d stamps)" % (w, n))
    prev = None
    line = []
    for i in range(min(n, 510)):
        v = int(buf[w, i]); tag = v >> 48; t = (v & 0xffffffffffff) - t00
        nm = names.get(tag, ("P c%d start" % (tag - 10) if 10 <= tag < 20 else "P c%d full" % (tag - 20) if 20 <= tag < 30 else
                             "M c%d afull" % (tag - 50) if 50 <= tag < 60 else "M c%d issued" % (tag - 70) if 70 <= tag < 80 else
                             "T hs" if tag == 90 else {120: "E it", 121: "E ld", 122: "E acc", 123: "F in", 124: "F rows", 125: "F out"}.get(tag, str(tag))))
        line.append("%s@%d" % (nm, t))
        if len(line) == 6:
            print("   " + "  ".join(line)); line = []
        if i > 130: break
    if line: print("   " + "  ".join(line))
