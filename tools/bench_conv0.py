"""Measurement of row f2 (conv0 + ReLU on the front-end's features), same rules as bench.py:
CUDA events, >= 3 warm-ups, outputs cycled over a pool larger than L2, one JSON line.
`value` = audio-seconds/s through the layer at the AISHELL shape; roofline = HBM (the layer writes
C*D1/(2D) = 15.6x its input); `library` = the same op through torch / cuDNN; `cpu_baseline` = oracle."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from openasr_b200.blocks.conv_layers import conv0_relu_forward
from oracle import conv_oracle as co

dev = torch.device("cuda", 0)
B, T, D, C = 32, 649, 80, 32
audio_s = 32 * 5.0  # the AISHELL batch these features come from (mean 5 s per utterance)
gen = torch.Generator().manual_seed(0)
x = (4.0 * torch.randn(B, T, D, generator=gen) + 8.0).to(dev)
w = (0.3 * torch.randn(C, 1, 3, 3, generator=gen)).to(dev)
b = (0.1 * torch.randn(C, generator=gen)).to(dev)
T1, D1 = (T - 3) // 2 + 1, D - 2
alg_bytes = 4 * B * T * D + 4 * B * C * T1 * D1
from openasr_b200 import _capi
import ctypes as Ct
lib = _capi.load()
outs = [torch.empty((B, C, T1, D1), device=dev) for _ in range(3)]  # 3 x 103 MB > L2
stream = torch.cuda.Stream(device=dev)

def ours(i):
    o = outs[i % len(outs)]
    _capi.check(lib.spl_conv0_relu(None, Ct.c_void_p(x.data_ptr()), B, T, D, Ct.c_void_p(w.data_ptr()), Ct.c_void_p(b.data_ptr()),
                                   C, Ct.c_void_p(o.data_ptr()), Ct.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))

def library(i):
    torch.relu_(torch.nn.functional.conv2d(x.unsqueeze(1), w, b, stride=(2, 1)))

def timed(fn, K=60):
    with torch.cuda.stream(stream):
        for i in range(5):
            fn(i)
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            for i in range(K):
                fn(i)
        g.replay()
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):
            e0.record(stream)
            g.replay()
            e1.record(stream)
            stream.synchronize()
            best = min(best, e0.elapsed_time(e1))
    return 1e3 * best / K  # us per launch

us_ours = timed(ours)
us_lib = timed(library)
peak, src = bench.measured_peaks()
xc, wc, bc = x.cpu(), w.cpu(), b.cpu()
torch.set_num_threads(os.cpu_count() or 1)
co.conv0_relu(xc, wc, bc)
t0 = time.perf_counter()
co.conv0_relu(xc, wc, bc)
cpu_s = time.perf_counter() - t0
print(json.dumps({
    "metric": "audio-sec/sec through conv0+ReLU (row f2)", "value": audio_s / (us_ours * 1e-6), "unit": "audio-s/s",
    "n_gpus": 1, "us_per_launch": us_ours, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
    "config": {"workload": "aishell features 32x649x80 -> 32x32x324x78", "pool": "3 outputs of 103 MB (> L2)"},
    "roofline": {"bound": "hbm", "achieved": alg_bytes / (us_ours * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
                 "frac": alg_bytes / (us_ours * 1e-6) / 1e9 / peak, "alg_bytes_per_launch": alg_bytes, "peak_source": src},
    "library": {"impl": "torch conv2d + relu_ (cuDNN)", "us_per_launch": us_lib},
    "cpu_baseline": {"value": audio_s / cpu_s, "unit": "audio-s/s", "cores": os.cpu_count(), "kind": "port",
                     "sample": "the whole batch once after one warm-up"}}))
