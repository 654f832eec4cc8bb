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
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    layer(xd, lens_list)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
