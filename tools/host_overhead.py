"""Host-side cost of one SPLayer.forward call (no copies, no sync): Python + ctypes + launch overhead.
Run on a GPU box: python tools/host_overhead.py"""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from openasr_b200 import SPLayer
from openasr_b200.synth import synth_batch

dev = torch.device("cuda", 0)
conf = bench.workload_config("aishell", 1.0)
layer = SPLayer(conf).to(dev).train()
x, lens = synth_batch(32, 56000, 104000, 16000, seed=1)
xd = x.to(dev)
xi = x.to(torch.int16).to(dev)
lens_list = lens.tolist()
for name, inp in (("fp32", xd), ("int16", xi)):
    for _ in range(20):
        layer(inp, lens_list)
    torch.cuda.synchronize()
    N = 300
    t0 = time.perf_counter()
    for _ in range(N):
        layer(inp, lens_list)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("%s: host %.1f us/call (issue only), %.1f us/call incl. drain" % (name, 1e6 * (t1 - t0) / N, 1e6 * (t2 - t0) / N))
# forward_multi: 8 batches per call (the per-call Python cost is paid once)
batches = [(xd, lens_list)] * 8
for _ in range(10):
    layer.forward_multi(batches)
torch.cuda.synchronize()
N = 100
t0 = time.perf_counter()
for _ in range(N):
    layer.forward_multi(batches)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("forward_multi(8 batches): host %.1f us/call = %.1f us/batch (issue only)" % (1e6 * (t1 - t0) / N, 1e6 * (t1 - t0) / N / 8))
# sync-free: CUDA lengths, T from the padded width
conf_sf = dict(conf, sync_free=True)
layer_sf = SPLayer(conf_sf).to(dev).train()
xw = xd[:, :int(lens.max())].contiguous()
ld = lens.to(dev)
for _ in range(20):
    layer_sf(xw, ld)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300):
    layer_sf(xw, ld)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("sync_free forward (device lengths, no D2H): host %.1f us/call" % (1e6 * (t1 - t0) / 300))
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    layer(xd, lens_list)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
