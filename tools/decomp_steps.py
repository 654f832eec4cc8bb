import os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import bench
from openasr_b200 import SPLayer, frontend
dev = torch.device("cuda", 0)
wl = "aishell"
dither = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
conf = bench.workload_config(wl, dither)
layer = SPLayer(conf).to(dev).train()
h, items = bench.build_pool(layer, conf, wl, 16, dev, seed=1234)
stream = torch.cuda.Stream(device=dev)
sa = conf.get("spec_aug")
def mk(mode):
    def step(i):
        it = items[i % len(items)]
        if mode in ("A", "AB"):
            h.fbank(it["wav"], it["lens"], it["T"], dither_seed=1 + i, utt_stats=it["stats"], out=it["feats"], feat_len=it["flen"])
        if mode == "A0":
            h.fbank(it["wav"], it["lens"], it["T"], dither_seed=1 + i, utt_stats=None, out=it["feats"], feat_len=it["flen"])
        if mode in ("B", "AB"):
            frontend.post_inplace(it["feats"], it["flen"], cmvn_mode=conf["cmvn"], utt_stats=it["stats"], mask_params=it["rect"],
                                  n_freq=sa["freq_mask_num"], n_time=sa["time_mask_num"])
    return step
with torch.cuda.stream(stream):
    for i in range(4):
        mk("AB")(i)
stream.synchronize()
K = 256
for mode in ("A0", "A", "B", "AB"):
    for ns in (1, 2, 4, 8):
        side = [torch.cuda.Stream(device=dev) for _ in range(ns - 1)]
        graphs, reps, rem = bench.time_graphed(mk(mode), K, 64, stream, side)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            with torch.cuda.stream(stream):
                e0.record(stream)
                for _r in range(reps):
                    graphs[0][0].replay()
                e1.record(stream)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print("mode %-3s streams %d: %.2f us/step" % (mode, ns, 1e3 * best / K), flush=True)
