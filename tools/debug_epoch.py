import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
from openasr_b200 import SPLayer
conf = bench.workload_config("epoch", 1.0)
layer = SPLayer(conf).to(dev).train()
h, items = bench.build_pool(layer, "epoch", int(os.environ.get("DBG_POOL", "16")), dev, seed=1234 + 100003 * rank)
gstats = torch.zeros(2 * h.d_out + 1, dtype=torch.float64, device=dev)
layer.set_global_cmvn(torch.zeros(h.d_out, device=dev), torch.ones(h.d_out, device=dev))
stream = torch.cuda.Stream(device=dev)
sp = C.c_void_p(stream.cuda_stream)
for mode in ("stats", "full"):
    groups = bench.make_groups(h, items, conf, layer, 8, int(os.environ.get("DBG_POOL", "16")), 0, mode=mode, global_stats=gstats if mode == "stats" else None)
    for i, g in enumerate(groups):
        with torch.cuda.stream(stream):
            g.run(i, sp)
        torch.cuda.synchronize(dev)
        print("rank", rank, mode, "group", i, "ok", flush=True)
if world > 1:
    dist.all_reduce(gstats.clone())
    torch.cuda.synchronize(dev)
    print("rank", rank, "allreduce ok", flush=True)
    # graphs + collective as the bench does
    g1 = bench.graph_of(bench.make_groups(h, items, conf, layer, 8, int(os.environ.get("DBG_POOL", "16")), 0, mode="stats", global_stats=gstats), stream)
    g2 = bench.graph_of(bench.make_groups(h, items, conf, layer, 8, 16, 0), stream)
    with torch.cuda.stream(stream):
        g1.replay(); g2.replay()
    torch.cuda.synchronize(dev)
    print("rank", rank, "graphs ok", flush=True)
    with torch.cuda.stream(stream):
        g1.replay()
        dist.all_reduce(gstats)
        g2.replay()
    torch.cuda.synchronize(dev)
    print("rank", rank, "graph+collective ok", gstats[-1].item(), flush=True)
    dist.destroy_process_group()
