import os, sys
sys.path.insert(0, os.getcwd())
import torch, bench
from openasr_b200 import SPLayer
dev = torch.device("cuda", 0)
for dither in (0.0, 1.0):
    conf = bench.workload_config("aishell", dither)
    layer = SPLayer(conf).to(dev).train()
    h, items = bench.build_pool(layer, conf, "aishell", 16, dev, seed=1234)
    for it in items:
        it["wav16"] = it["wav"].to(torch.int16)
    stream = torch.cuda.Stream(device=dev)
    for name in ("wav", "wav16"):
        def step(i):
            it = items[i % len(items)]
            h.fbank(it[name], it["lens"], it["T"], dither_seed=1 + i, utt_stats=it["stats"], out=it["feats"], feat_len=it["flen"])
        with torch.cuda.stream(stream):
            for i in range(4): step(i)
        stream.synchronize()
        graphs, reps, rem = bench.time_graphed(step, 256, 64, stream, ())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize()
            with torch.cuda.stream(stream):
                e0.record(stream)
                for _r in range(reps): graphs[0][0].replay()
                e1.record(stream)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print("dither %.0f input %-5s: kernel A %.2f us" % (dither, name, 1e3 * best / 256), flush=True)
