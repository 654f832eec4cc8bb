"""CPU, world_size 2 over gloo: the N>1 host logic (utterance sharding, global-CMVN all-reduce)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openasr_b200 import cmvn
from oracle import frontend_oracle as fo


def test_shard_utterances_partition_and_balance():
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(56000, 104000, (32,), generator=g).tolist()
    for world in (1, 2, 4, 8):
        parts = [cmvn.shard_utterances(lens, world, r) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(32))
        assert {len(p) for p in parts} == {32 // world}
        loads = [sum(lens[i] for i in p) for p in parts]
        assert max(loads) / (sum(loads) / world) < 1.05
    parts = [cmvn.shard_utterances([5, 4, 3], 2, r) for r in range(2)]
    assert sorted(map(len, parts)) == [1, 2]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    x, lens = fo.synth_batch(6, 8000, 20000, 16000, seed=5)
    lens = lens.tolist()
    mine = cmvn.shard_utterances(lens, world, rank)
    conf = {"feature_type": "fbank", "sample_rate": 16000, "num_mel_bins": 40, "use_energy": False, "dither": 0.0}
    feats, flen = fo.splayer_forward(x[mine], [lens[i] for i in mine], conf)
    st = fo.cmvn_stats(feats, flen.tolist())          # [3, D]
    packed = torch.cat([st[0], st[1], st[2, :1]])      # product layout [2*D+1]
    cmvn.all_reduce_stats(packed)
    mean, istd = cmvn.finalize_stats(packed)
    q.put((rank, mine, mean, istd))
    dist.barrier()
    dist.destroy_process_group()


def test_global_cmvn_allreduce_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process truth
    x, lens = fo.synth_batch(6, 8000, 20000, 16000, seed=5)
    conf = {"feature_type": "fbank", "sample_rate": 16000, "num_mel_bins": 40, "use_energy": False, "dither": 0.0}
    feats, flen = fo.splayer_forward(x, lens.tolist(), conf)
    st = fo.cmvn_stats(feats, flen.tolist())
    mean = st[0] / st[2]
    istd = (st[1] / st[2] - mean * mean).clamp_min(1e-20).rsqrt()
    for rank, mine, m, s in res:
        assert torch.allclose(m, mean, rtol=1e-12, atol=1e-12)
        assert torch.allclose(s, istd, rtol=1e-10, atol=1e-12)
    assert sorted(i for r in res for i in r[1]) == list(range(6))
