"""torchrun worker of tests/test_gpu_multi.py::test_global_cmvn_nccl_two_ranks (one process per GPU, NCCL)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def main():
    from openasr_b200 import SPLayer
    from openasr_b200.cmvn import GlobalCmvn, shard_utterances
    from oracle import frontend_oracle as fo
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    conf = {"feature_type": "fbank", "sample_rate": 16000, "num_mel_bins": 80, "use_energy": False, "dither": 0.0,
            "cmvn": "global"}
    layer = SPLayer(conf).to(dev).eval()
    batches = [fo.synth_batch(8, 4000, 50000, 16000, seed=300 + k) for k in range(3)]  # the same on every rank
    gc = GlobalCmvn(layer, dev)
    mine = []
    for w, l in batches:
        idx = shard_utterances(l.tolist(), world, rank)
        ws, ls = w[idx][:, :int(l[idx].max())].contiguous(), l[idx]
        mine.append((ws, ls, idx))
        gc.update(ws.to(dev), ls)
    mean, istd = gc.finalize()          # ONE all_reduce(SUM) of 2D+1 doubles over NCCL
    torch.cuda.synchronize()
    # oracle: fp64 statistics over ALL utterances of all batches
    feats = []
    for w, l in batches:
        for i in range(w.shape[0]):
            feats.append(fo.fbank(w[i, :l[i]], 16000.0, 80, dither=0.0, dtype=torch.float64))
    allf = torch.cat(feats)
    m64, s64 = allf.mean(0), allf.std(0, unbiased=False)
    e_mean = (mean.cpu() - m64).abs().max().item()
    e_istd = ((istd.cpu() - 1.0 / s64).abs() * s64).max().item()
    count = int(gc.stats[-1].item())
    ok = e_mean < 2e-4 and e_istd < 2e-4 and count == allf.shape[0]
    # features of this rank's shard, normalised with the global statistics
    ws, ls, idx = mine[0]
    f, fl = layer(ws.to(dev), ls)
    w0, l0 = batches[0]
    e_feat = 0.0
    for j, i in enumerate(idx):
        ref = (fo.fbank(w0[i, :l0[i]], 16000.0, 80, dither=0.0, dtype=torch.float64) - m64) / s64
        e_feat = max(e_feat, (f[j, :ref.shape[0]].cpu().double() - ref).abs().max().item())
    t = torch.tensor([e_feat], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and t.item() < 5e-3
    if rank == 0:
        json.dump({"world": world, "ok": bool(ok), "mean_err": e_mean, "istd_rel_err": e_istd, "frames": count,
                   "feat_err_max_over_ranks": float(t.item())}, open(sys.argv[1], "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
