#!/usr/bin/env python
"""bench.py -- audio-seconds/second of the online speech front-end (fbank + CMVN + SpecAug).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload aishell|hkust|libri|epoch] [--shard] [--batches-per-launch 8]

A "step" is one pass of the hot path (kernel A: fused fbank, kernel B: CMVN + SpecAug) over ONE synthetic batch
of the named shape; `--batches-per-launch` consecutive steps share one persistent launch of each kernel
(spl_forward_multi, the queue-of-batches form of the reference's per-utterance loop).  Default workload =
BASELINE.json configs[1], the AISHELL-1 shape: 32 x ~5 s, 16 kHz, 80-dim fbank, reference-default dither (1.0,
device RNG), utterance CMVN, SpecAug 2x27 F + 2x40 T, training mode.

  value        device-resident throughput: inputs already in HBM, a pool of distinct batches larger than L2
               cycled under CUDA graphs, CUDA events, max over ranks.
  e2e          the same metric through the public module API (SPLayer.forward_multi) with HOST buffers:
               pinned fp32 wav -> H2D, forward (host RNG draws, uploads, both kernels), D2H of features + lengths.
  roofline     kernel A alone over the same pool; algorithmic bytes = 4*sum n_i + 4*B*T*D + 16*B per batch
                against MEASURED_PEAKS hbm_gbs.
  cpu_baseline the oracle port of the reference's CPU path on this box's host cores (N = 1 only).
  secondary    HKUST and LibriSpeech shapes measured by the same code in the same run (N = 1 only).
  copy_control the e2e staging alone (same pinned buffers and streams, no kernels).

--workload epoch: BASELINE configs[4] -- 1000 h of AISHELL-shaped batches split over the ranks, two passes: pass 1
accumulates (sum x, sum x^2, count) in kernel A's epilogue, ONE NCCL all-reduce of 2D+1 doubles, pass 2
featurizes with global CMVN + SpecAug.  --workload libri --shard: configs[3], every 16-utterance batch split over
the ranks by sample count (strong scaling).

With --impl reference the oracle port (the reference is pure Python, there is no oracle/_ref binary) is timed on
the host cores with all threads; each step is a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

# dram__bytes_read.sum + dram__bytes_write.sum of kernel A per launch (ncu --set full), per workload; filled from
# profiles/r2_traffic.json when present (written by tools/ncu_traffic.py from the committed ncu captures)
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r2_traffic.json")

WORKLOADS = {
    # name: (B, n_lo, n_hi, sample_rate, D, cmvn, spec_aug)
    "aishell": (32, 56000, 104000, 16000, 80, "utterance",
                {"freq_mask_num": 2, "freq_mask_width": 27, "time_mask_num": 2, "time_mask_width": 40}),
    "hkust": (64, 16000, 48000, 8000, 40, "utterance", None),
    "libri": (16, 192000, 320000, 16000, 80, "none",
              {"freq_mask_num": 2, "freq_mask_width": 27, "time_mask_num": 2, "time_mask_width": 100}),
}
WORKLOADS["epoch"] = WORKLOADS["aishell"][:5] + ("global", WORKLOADS["aishell"][6])


def workload_config(name, dither):
    B, lo, hi, sr, D, cmvn, sa = WORKLOADS[name]
    conf = {"feature_type": "fbank", "sample_rate": sr, "num_mel_bins": D, "use_energy": False,
            "dither": dither, "cmvn": cmvn, "dither_rng": "device", "specaug_rng": "host"}
    if sa is not None:
        conf["spec_aug"] = dict(sa)
    return conf


def public_config(name, dither):
    """The workload description both arms print (identical keys and values)."""
    B, lo, hi, sr, D, cmvn, sa = WORKLOADS[name]
    return {"workload": name, "batch": B, "samples_per_utt": [lo, hi], "sample_rate": sr, "num_mel_bins": D,
            "cmvn": cmvn, "spec_aug": sa, "dither": dither, "training": True}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


def ncu_traffic(name):
    try:
        return json.load(open(TRAFFIC_FILE)).get(name)
    except Exception:
        return None


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """Reference arm: the oracle port of the reference's CPU path, all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import frontend_oracle as fo
    wl = "aishell" if args.workload == "epoch" else args.workload
    B, lo, hi, sr, D, cmvn, sa = WORKLOADS[wl]
    conf = workload_config(wl, args.dither)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wav, lens = fo.synth_batch(B, lo, hi, sr, seed=1234)
    lens_l = lens.tolist()
    # calibrate the per-step sample so that (K + W) steps end within ~150 s
    t0 = time.perf_counter()
    fo.splayer_forward(wav[:1, :lens_l[0]], lens_l[:1], conf, training=True)
    t_utt = max(time.perf_counter() - t0, 1e-3)
    n_steps = args.steps + args.warmup
    per_step = max(1, min(B, int(150.0 / n_steps / t_utt)))
    sub = wav[:per_step, :max(lens_l[:per_step])].contiguous()
    sub_l = lens_l[:per_step]
    audio_s = sum(sub_l) / sr
    for _ in range(args.warmup):
        fo.splayer_forward(sub, sub_l, conf, training=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fo.splayer_forward(sub, sub_l, conf, training=True)
    dt = time.perf_counter() - t0
    value = audio_s * args.steps / dt
    sample = "%d of %d utterances (%.1f audio-s) per step" % (per_step, B, audio_s)
    line = {
        "impl": "reference", "metric": "audio-sec/sec featurized (fbank+CMVN+SpecAug)", "value": value,
        "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": public_config(wl, args.dither),
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def build_pool(layer, wl, pool, dev, seed, shard=None):
    """Device-resident pool of distinct batches + everything a step needs (lengths, SpecAug uniforms, output
    buffers), so the timed region holds only the hot-path launches.  ``shard = (world, rank)``: every batch is
    generated whole (same seed on every rank) and this rank keeps its utterances (cmvn.shard_utterances)."""
    from openasr_b200 import frontend, tables
    from openasr_b200.cmvn import shard_utterances
    from openasr_b200.synth import synth_batch
    B, lo, hi, sr, D, cmvn, sa = WORKLOADS[wl]
    h = layer._handle(dev)
    items = []
    for i in range(pool):
        wav, lens = synth_batch(B, lo, hi, sr, seed=seed + 7919 * i)
        if shard is not None:
            idx = shard_utterances(lens.tolist(), shard[0], shard[1])
            lens = lens[idx]
            wav = wav[idx][:, :int(lens.max())].contiguous()
        nb = wav.shape[0]
        frames = [tables.frame_count(int(n), h.win, h.shift) for n in lens.tolist()]
        T = max(frames)
        it = {
            "wav_host": wav, "lens_host": lens, "wav": wav.to(dev), "lens": lens.to(dev), "T": T, "frames": frames,
            "feats": torch.empty((nb, T, h.d_out), dtype=torch.float32, device=dev),
            "flen": torch.empty((nb,), dtype=torch.int64, device=dev),
            "stats": torch.empty((nb, 2, h.d_out), dtype=torch.float64, device=dev),
            "audio_s": float(lens.sum()) / sr,
            "alg_bytes": 4 * int(lens.sum()) + 4 * nb * T * h.d_out + 16 * nb,
            "uni": None,
        }
        if sa is not None:
            it["uni"] = frontend.specaug_uniforms(nb, sa["freq_mask_num"], sa["time_mask_num"]).to(dev)
        items.append(it)
    return h, items


class Group:
    """Pre-built argument arrays of one spl_forward_multi call over consecutive pool batches."""

    def __init__(self, h, items, conf, layer, mode="full", global_stats=None):
        from openasr_b200 import _capi
        self.h, self.n = h, len(items)
        sa = conf.get("spec_aug")
        need_stats = conf["cmvn"] == "utterance" or (sa is not None and sa["time_mask_num"] > 0)
        self.fa = (_capi.SplFbankArgs * self.n)()
        self.pa = (_capi.SplPostArgs * self.n)()
        self.keep = []
        self.post = mode == "full" and (conf["cmvn"] != "none" or sa is not None)
        # one contiguous fp64 buffer for the group's per-utterance sums: the library zeroes it with ONE memset
        tot_b = sum(it["wav"].shape[0] for it in items)
        stats_all = torch.empty((tot_b, 2, h.d_out), dtype=torch.float64, device=items[0]["wav"].device)
        self.keep.append(stats_all)
        b0 = 0
        for k, it in enumerate(items):
            nb = it["wav"].shape[0]
            st = stats_all[b0:b0 + nb]
            b0 += nb
            self.keep.append(h._fill_args(self.fa[k], it["wav"], it["lens"], it["T"], None, 0,
                                          st if (need_stats and mode != "stats") else None,
                                          global_stats if mode == "stats" else None, it["feats"], it["flen"]))
            a = self.pa[k]
            a.Dm = h.d_out
            a.cmvn_mode = _capi.CMVN_MODES[conf["cmvn"]]
            a.norm_vars = 1
            if conf["cmvn"] == "global":
                a.global_mean = layer._gmean.data_ptr()
                a.global_istd = layer._gistd.data_ptr()
            if sa is not None:
                a.n_freq_masks, a.n_time_masks = sa["freq_mask_num"], sa["time_mask_num"]
                a.mask_uniforms = it["uni"].data_ptr()
                a.freq_mask_width, a.time_mask_width = float(sa["freq_mask_width"]), float(sa["time_mask_width"])
        self.launches = 1 + (1 if self.post else 0)

    def run(self, seed, stream_ptr):
        from openasr_b200 import _capi
        for k in range(self.n):
            self.fa[k].dither_seed = seed & 0xFFFFFFFFFFFFFFFF
        _capi.check(self.h._lib.spl_forward_multi(self.h._h, self.fa, self.pa if self.post else None, self.n, stream_ptr),
                    "spl_forward_multi")


def make_groups(h, items, conf, layer, kb, n_steps, start, mode="full", global_stats=None):
    """Groups covering n_steps consecutive pool batches beginning at pool position `start`."""
    groups, done = [], 0
    while done < n_steps:
        n = min(kb, n_steps - done)
        sel = [items[(start + done + j) % len(items)] for j in range(n)]
        groups.append(Group(h, sel, conf, layer, mode, global_stats))
        done += n
    return groups


def graph_of(groups, stream, seed0=0x9E3779B97F4A7C15, side=()):
    """One CUDA graph of all groups.  side: extra streams forked inside the capture -- group i runs (kernel A then
    kernel B) on lane i % (1 + len(side)), so that the launch of one group fills the tail of the previous one."""
    g = torch.cuda.CUDAGraph()
    lanes = [stream] + list(side)
    with torch.cuda.stream(stream):
        with torch.cuda.graph(g, stream=stream):
            if side:
                fork = torch.cuda.Event()
                fork.record(stream)
                for s in side:
                    s.wait_event(fork)
            for i, gr in enumerate(groups):
                ln = lanes[i % len(lanes)]
                with torch.cuda.stream(ln):
                    gr.run(seed0 + i, C.c_void_p(ln.cuda_stream))
            for s in side:
                join = torch.cuda.Event()
                join.record(s)
                stream.wait_event(join)
    return g


def time_graph(g, stream, repeats, barrier):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = None
    for _ in range(repeats):
        barrier()
        with torch.cuda.stream(stream):
            e0.record(stream)
            g.replay()
            e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


def measure_workload(args, wl, dev, rank, world, dist, want_e2e=True, shard=None, sampler=None):
    """value / roofline / e2e of one workload on this rank (max over ranks inside)."""
    from openasr_b200 import SPLayer
    B, lo, hi, sr, D, cmvn, sa = WORKLOADS[wl]
    conf = workload_config(wl, args.dither)
    layer = SPLayer(conf).to(dev).train()
    h, items = build_pool(layer, wl, args.pool, dev, seed=1234 + (0 if shard else 100003 * rank), shard=shard)
    K, W, kb = args.steps, max(args.warmup, 3), args.batches_per_launch
    stream = torch.cuda.Stream(device=dev)
    sp = C.c_void_p(stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def allmax(v):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return v

    def allsum(v):
        if world > 1:
            t = torch.tensor([v], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return float(t.item())
        return v

    # warm-up (eager; also configures the kernels' shared-memory attributes), on pool positions [0, W)
    with torch.cuda.stream(stream):
        for i, gr in enumerate(make_groups(h, items, conf, layer, kb, W, 0)):
            gr.run(i, sp)
    stream.synchronize()
    # timed graph on the pool positions that FOLLOW the warm-up (nothing of it is L2-warm: the pool is several
    # times the 126 MB L2 and is walked in order)
    groups = make_groups(h, items, conf, layer, kb, K, W)
    side = [torch.cuda.Stream(device=dev) for _ in range(max(1, args.streams) - 1)]
    g = graph_of(groups, stream, side=side)
    with torch.cuda.stream(stream):
        g.replay()  # graph warm-up (its inputs are evicted again by the time the replay wraps around the pool)
    stream.synchronize()
    ms_total = allmax(time_graph(g, stream, args.repeats, barrier))
    if sampler is not None:  # same load, untimed, so that nvidia-smi (100 ms period) sees the clocks under load
        t_end = time.perf_counter() + 0.7
        while time.perf_counter() < t_end:
            with torch.cuda.stream(stream):
                for _r in range(8):
                    g.replay()
            stream.synchronize()
    audio_total = allsum(sum(items[(W + i) % len(items)]["audio_s"] for i in range(K)))
    value = audio_total / (ms_total * 1e-3)
    launches = sum(gr.launches for gr in groups)

    # ---- roofline: kernel A alone over the same pool positions ----
    groups_a = make_groups(h, items, conf, layer, kb, K, W, mode="a")
    ga = graph_of(groups_a, stream)
    with torch.cuda.stream(stream):
        ga.replay()
    stream.synchronize()
    ms_a = time_graph(ga, stream, args.repeats, lambda: torch.cuda.synchronize(dev))
    alg_bytes = sum(items[(W + i) % len(items)]["alg_bytes"] for i in range(K))
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (ms_a * 1e-3) / 1e9
    tr = ncu_traffic(wl)
    roofline = {"bound": "hbm", "kernel": "fbank_umma_kernel" if h.engine_name() == "umma" else "fbank_warp_kernel",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": tr["bytes_per_launch"] if tr else None, "traffic_source": tr["source"] if tr else None,
                "peak_source": peak_src, "us_per_launch": 1e3 * ms_a / len(groups_a),
                "us_per_batch": 1e3 * ms_a / K, "batches_per_launch": kb,
                "alg_bytes_per_launch": alg_bytes / len(groups_a),
                "step_share": ms_a / ms_total if world == 1 else None}
    # secondary ceiling: algorithmic fp32 flops of the path against the CUDA-core peak
    nfft = 512 if sr > 8000 else 256
    nw = int(sr * 0.025)
    flops_per_frame = 2.5 * nfft * __import__("math").log2(nfft) + 5 * nw + 1.5 * nfft + 2.0 * nfft
    frames = sum(float(items[(W + i) % len(items)]["flen"].sum().item()) for i in range(min(K, len(items)))) * (K / min(K, len(items)))
    fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12  # 148 SMs x 128 FMA lanes x 2 flops x 1.965 GHz = 74.4 TFLOP/s
    roofline["fp32_ceiling"] = {"flops_per_frame": float(flops_per_frame), "achieved": frames * flops_per_frame / (ms_a * 1e-3) / 1e12,
                                "peak": fp32_peak, "unit": "TFLOP/s",
                                "frac": frames * flops_per_frame / (ms_a * 1e-3) / 1e12 / fp32_peak,
                                "note": "real FFT 2.5 N log2 N + framing 5 Nw + power 1.5 N + mel 2 N flops per frame; nominal CUDA-core peak"}
    out = {"value": value, "ms_total": ms_total, "launches": launches, "roofline": roofline,
           "pool_bytes": sum(it["wav"].numel() * 4 + it["feats"].numel() * 4 for it in items),
           "engine": h.engine_name()}
    if want_e2e and world == 1 and kb > 1:  # one batch per launch: the per-step figure that is NOT amortised over kb batches
        n1 = min(K, 64)
        groups_1 = make_groups(h, items, conf, layer, 1, n1, W)
        g1 = graph_of(groups_1, stream)
        with torch.cuda.stream(stream):
            g1.replay()
        stream.synchronize()
        ms_1 = time_graph(g1, stream, args.repeats, lambda: torch.cuda.synchronize(dev))
        out["single_batch_launch"] = {"us_per_step": 1e3 * ms_1 / n1, "batches_per_launch": 1, "steps": n1,
                                      "note": "kernel A + kernel B launched once per batch, same graph / pool hygiene"}
    if want_e2e:
        n_e2e = max(kb, min(K, args.e2e_steps) // kb * kb)
        out["copy_control"] = run_e2e(layer, items, dev, n_e2e, kb, world, dist, copy_only=True)
        out["e2e"] = run_e2e(layer, items, dev, n_e2e, kb, world, dist)
        out["e2e_int16"] = run_e2e(layer, items, dev, n_e2e, kb, world, dist, int16=True)
    out["items"] = items
    out["conf"] = conf
    return out


def run_e2e(layer, items, dev, steps, kb, world, dist, int16=False, copy_only=False):
    """SPLayer.forward_multi from pinned host buffers, `kb` batches per call: H2D of the wav batches, forward (host
    RNG draws, pinned upload of lengths + uniforms, both kernels), D2H of features and lengths; 3 calls in flight on
    3 streams.  copy_only: the same staging without the kernels (what the host side of the box sustains)."""
    nslot = 3
    streams = [torch.cuda.Stream(device=dev) for _ in range(nslot)]
    maxT = max(it["T"] for it in items)
    B = max(it["wav_host"].shape[0] for it in items)
    d_out = items[0]["feats"].shape[2]
    pin_in = [(it["wav_host"].to(torch.int16) if int16 else it["wav_host"]).pin_memory() for it in items]
    pin_out = [[torch.empty((B * maxT * d_out,), dtype=torch.float32).pin_memory() for _ in range(kb)] for _ in range(nslot)]
    pin_len = [[torch.empty((B,), dtype=torch.int64).pin_memory() for _ in range(kb)] for _ in range(nslot)]
    dev_in = [[torch.empty_like(max(pin_in, key=lambda t: t.numel()), device=dev) for _ in range(kb)] for _ in range(nslot)]
    dev_out = [[torch.empty((B * maxT * d_out,), dtype=torch.float32, device=dev) for _ in range(kb)] for _ in range(nslot)]
    done = [None] * nslot
    h2d = sum(p.numel() * p.element_size() for p in pin_in) / len(pin_in) + 8 * B
    d2h = sum(it["feats"].numel() * 4 for it in items) / len(items) + 8 * B

    def one(call):
        s = call % nslot
        if done[s] is not None:
            done[s].synchronize()
        with torch.cuda.stream(streams[s]):
            batch = []
            for j in range(kb):
                i = (call * kb + j) % len(items)
                src = pin_in[i]
                wav = dev_in[s][j].view(-1)[:src.numel()].view(src.shape)
                wav.copy_(src, non_blocking=True)
                batch.append((wav, items[i]["lens_host"]))
            if copy_only:
                for j in range(kb):
                    i = (call * kb + j) % len(items)
                    n = items[i]["feats"].numel()
                    pin_out[s][j][:n].copy_(dev_out[s][j][:n], non_blocking=True)
                    pin_len[s][j][:items[i]["flen"].numel()].copy_(items[i]["flen"], non_blocking=True)
            else:
                outs = layer.forward_multi(batch)
                for j, (feats, flen) in enumerate(outs):
                    pin_out[s][j][:feats.numel()].view_as(feats).copy_(feats, non_blocking=True)
                    pin_len[s][j][:flen.numel()].copy_(flen, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(streams[s])
            done[s] = ev

    ncall = steps // kb
    for c in range(3):
        one(c)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for c in range(ncall):
        one(c)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    audio = sum(items[i % len(items)]["audio_s"] for i in range(ncall * kb))
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        t = torch.tensor([audio], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        audio = float(t.item())
    api = "copies only (no kernels)" if copy_only else "SPLayer.forward_multi(%d batches per call, host lengths)" % kb
    return {"value": audio / dt, "unit": "audio-s/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "steps": ncall * kb, "api": "pinned %s wav -> cuda, %s, D2H of feats + lengths; 3 calls in flight"
                                        % ("int16 PCM" if int16 else "fp32", api)}


def run_cpu_baseline(conf, items, sr):
    """Oracle port of the reference CPU path on a bounded sample (about 10-30 s of CPU work)."""
    from oracle import frontend_oracle as fo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    it = items[0]
    lens = it["lens_host"].tolist()
    t0 = time.perf_counter()
    fo.splayer_forward(it["wav_host"][:1, :lens[0]], lens[:1], conf, training=True)
    t_utt = max(time.perf_counter() - t0, 1e-3)
    n = max(1, min(len(lens), int(8.0 / t_utt)))
    sub = it["wav_host"][:n, :max(lens[:n])].contiguous()
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        fo.splayer_forward(sub, lens[:n], conf, training=True)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    audio = sum(lens[:n]) / sr
    return {"value": audio / best, "unit": "audio-s/s", "cores": cores, "kind": "port",
            "sample": "%d of %d utterances of one batch (%.1f audio-s), best of 2 after 1 warm-up utterance" % (n, len(lens), audio)}


# ------------------------------------------------------------------------------------------------
def run_epoch(args, dev, rank, world, dist):
    """BASELINE configs[4]: `epoch_hours` of AISHELL-shaped batches, split evenly over the ranks; two passes with
    ONE all-reduce (NCCL) of the 2D+1 fp64 statistics in between.  Strong scaling (the epoch is fixed)."""
    from openasr_b200 import SPLayer
    from openasr_b200.cmvn import finalize_stats
    wl = "epoch"
    B, lo, hi, sr, D, cmvn, sa = WORKLOADS[wl]
    conf = workload_config(wl, args.dither)
    layer = SPLayer(conf).to(dev).train()
    h, items = build_pool(layer, wl, max(args.pool, 64), dev, seed=1234 + 100003 * rank)
    mean_audio = sum(it["audio_s"] for it in items) / len(items)
    total_batches = int(round(args.epoch_hours * 3600.0 / mean_audio))
    mine = total_batches // world + (1 if rank < total_batches % world else 0)
    kb = args.batches_per_launch
    stream = torch.cuda.Stream(device=dev)
    sp = C.c_void_p(stream.cuda_stream)
    gstats = torch.zeros(2 * h.d_out + 1, dtype=torch.float64, device=dev)
    layer.set_global_cmvn(torch.zeros(h.d_out, device=dev), torch.ones(h.d_out, device=dev))  # buffers the pass-2 arguments point at
    # eager warm-up of both passes and of the collective (first launches configure the kernels, the first collective
    # sets up NCCL's channels) BEFORE anything is captured; everything on `stream`
    with torch.cuda.stream(stream):
        for i, gr in enumerate(make_groups(h, items, conf, layer, kb, 2 * kb, 0, mode="stats", global_stats=gstats)):
            gr.run(i, sp)
        for i, gr in enumerate(make_groups(h, items, conf, layer, kb, 2 * kb, 0)):
            gr.run(i, sp)
    stream.synchronize()
    with torch.cuda.stream(stream):  # also the torch ops of finalize (their first call loads / tunes kernels)
        tmp = gstats.clone()
        if world > 1:
            dist.all_reduce(tmp)
        m0, i0 = finalize_stats(tmp)
        layer._gmean.copy_(m0.float())
        layer._gistd.copy_(i0.float())
    stream.synchronize()
    # one graph = one walk over the pool; the passes replay it mine / pool times (+ a remainder graph)
    full, rem = divmod(mine, len(items))
    alive = []  # the groups own the per-utterance statistics buffers the captured launches write: keep them

    def captured(n, mode):
        alive.append(make_groups(h, items, conf, layer, kb, n, 0, mode=mode, global_stats=gstats if mode == "stats" else None))
        return graph_of(alive[-1], stream)

    g1, g2 = captured(len(items), "stats"), captured(len(items), "full")
    g1r = captured(rem, "stats") if rem else None
    g2r = captured(rem, "full") if rem else None
    with torch.cuda.stream(stream):
        g1.replay()
        g2.replay()
    torch.cuda.synchronize(dev)
    gstats.zero_()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for _ in range(full):
            g1.replay()
        if g1r is not None:
            g1r.replay()
        ev[1].record(stream)
        if world > 1:
            dist.all_reduce(gstats)  # NCCL, on `stream` (the current stream): 2D+1 doubles over NVLink
        ev[4].record(stream)
        mean, istd = finalize_stats(gstats)
        layer._gmean.copy_(mean.float())
        layer._gistd.copy_(istd.float())
        ev[2].record(stream)
        for _ in range(full):
            g2.replay()
        if g2r is not None:
            g2r.replay()
        ev[3].record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ms = [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])]
    ms_ar = ev[1].elapsed_time(ev[4])
    audio = sum(items[i % len(items)]["audio_s"] for i in range(mine))
    t = torch.tensor([sum(ms), ms[0], ms[1], ms[2], ms_ar], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ta = torch.tensor([audio], dtype=torch.float64, device=dev)
        dist.all_reduce(ta)
        audio = float(ta.item())
    tot, p1, ar, p2, ar_only = [float(x) for x in t[:5]]
    count = float(gstats[2 * h.d_out].item())
    if rank != 0:
        return
    peak, peak_src = measured_peaks()
    line = {
        "metric": "audio-sec/sec featurized (fbank+CMVN+SpecAug)", "value": audio / (tot * 1e-3), "unit": "audio-s/s",
        "n_gpus": world, "steps": mine, "warmup": len(items), "ms_per_step": tot / max(mine, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(public_config(wl, args.dither), epoch_hours=args.epoch_hours, total_batches=total_batches),
        "harness": {"pool_batches": len(items), "batches_per_launch": kb, "engine": h.engine_name(),
                    "partition": "epoch split evenly by batch over %d rank(s)" % world,
                    "collective": "one all_reduce(SUM) of %d fp64 (NCCL) between the passes" % (2 * h.d_out + 1) if world > 1 else "none (1 rank)"},
        "epoch": {"audio_hours": audio / 3600.0, "pass1_stats_ms": p1, "allreduce_ms": ar_only, "allreduce_finalize_ms": ar, "pass2_features_ms": p2,
                  "total_ms": tot, "global_frames": count,
                  "pass1_audio_s_per_s": audio / (p1 * 1e-3), "pass2_audio_s_per_s": audio / (p2 * 1e-3)},
        "gpu_launches": (mine + kb - 1) // kb * 3, "peak_source": peak_src,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch.distributed as dist
    from openasr_b200 import _capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = None
    if world > 1 and not args.no_numa_bind:
        from openasr_b200.batching import bind_to_gpu_numa_node
        numa_cpus = bind_to_gpu_numa_node(local)  # pinned staging buffers land on the GPU's NUMA node
    if args.workload == "epoch":
        run_epoch(args, dev, rank, world, dist)
        if world > 1:
            dist.destroy_process_group()
        return
    wl = args.workload
    B, lo, hi, sr, D, cmvn, sa = WORKLOADS[wl]
    shard = (world, rank) if (args.shard and world > 1) else None
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = _capi.launch_count()
    res = measure_workload(args, wl, dev, rank, world, dist, shard=shard, sampler=sampler)
    clocks = sampler.stop() if sampler else None
    secondary = None
    if world == 1 and wl == "aishell" and not args.no_secondary:
        secondary = {}
        for other in ("hkust", "libri"):
            r = measure_workload(args, other, dev, rank, world, dist)
            secondary[other] = {"config": public_config(other, args.dither), "value": r["value"], "unit": "audio-s/s",
                                "ms_per_step": r["ms_total"] / args.steps, "roofline": r["roofline"], "e2e": r["e2e"],
                                "e2e_int16": r["e2e_int16"]}
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu_baseline = run_cpu_baseline(res["conf"], res["items"], sr) if (world == 1 and not args.no_cpu_baseline) else None
    K = args.steps
    line = {
        "metric": "audio-sec/sec featurized (fbank+CMVN+SpecAug)", "value": res["value"], "unit": "audio-s/s",
        "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": res["ms_total"] / K,
        "higher_is_better": True, "scaling": "strong" if shard else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": public_config(wl, args.dither),
        "harness": {"engine": res["engine"], "dither_rng": "device", "pool_batches": args.pool, "pool_bytes": res["pool_bytes"],
                    "l2_flush": "pool larger than L2 (126 MB), walked in order; the timed steps follow the warm-up steps in the pool",
                    "batches_per_launch": args.batches_per_launch, "cuda_graph": True, "streams": args.streams,
                    "timing": "best of %d regions of K steps, CUDA events, amortised over %d batches per launch" % (args.repeats, args.batches_per_launch),
                    "partition": ("every batch sharded by utterance over %d ranks (strong scaling)" % world) if shard
                                 else "by utterance, %d rank(s), own pool per rank, no data-path collective" % world,
                    "numa_bound_cpus": len(numa_cpus) if numa_cpus else None},
        "roofline": res["roofline"], "cpu_baseline": cpu_baseline, "e2e": res["e2e"],
        "e2e_int16": res["e2e_int16"], "copy_control": res["copy_control"],
        "single_batch_launch": res.get("single_batch_launch"), "secondary": secondary,
        "gpu_launches": res["launches"], "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_conv0(args):
    """Row f2 (conv0 + ReLU on the front-end's features), same rules as the main arm: CUDA events,
    >= 3 warm-ups, outputs cycled over a pool larger than L2, one JSON line.  `value` = audio-seconds/s
    through the layer at the AISHELL shape; roofline = HBM (the layer writes C*D1/(2D) = 15.6x its
    input); `library` = the same op through torch / cuDNN; `cpu_baseline` = the oracle port."""
    from openasr_b200 import _capi
    dev = torch.device("cuda", 0)
    B, T, D, Cc = 32, 649, 80, 32
    audio_s = 32 * 5.0  # the AISHELL batch these features come from (mean 5 s per utterance)
    gen = torch.Generator().manual_seed(0)
    x = (4.0 * torch.randn(B, T, D, generator=gen) + 8.0).to(dev)
    w = (0.3 * torch.randn(Cc, 1, 3, 3, generator=gen)).to(dev)
    b = (0.1 * torch.randn(Cc, generator=gen)).to(dev)
    T1, D1 = (T - 3) // 2 + 1, D - 2
    alg_bytes = 4 * B * T * D + 4 * B * Cc * T1 * D1
    lib = _capi.load()
    outs = [torch.empty((B, Cc, T1, D1), device=dev) for _ in range(3)]  # 3 x 103 MB > L2
    stream = torch.cuda.Stream(device=dev)

    def ours(i):
        o = outs[i % len(outs)]
        _capi.check(lib.spl_conv0_relu(None, C.c_void_p(x.data_ptr()), B, T, D, C.c_void_p(w.data_ptr()),
                                       C.c_void_p(b.data_ptr()), Cc, C.c_void_p(o.data_ptr()),
                                       C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))

    def library(i):
        torch.relu_(torch.nn.functional.conv2d(x.unsqueeze(1), w, b, stride=(2, 1)))

    def timed(fn, K):
        with torch.cuda.stream(stream):
            for i in range(max(3, args.warmup)):
                fn(i)
            stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for i in range(K):
                    fn(i)
            g.replay()
            stream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = None
            for _ in range(args.repeats):
                e0.record(stream)
                g.replay()
                e1.record(stream)
                stream.synchronize()
                ms = e0.elapsed_time(e1)
                best = ms if best is None else min(best, ms)
        return 1e3 * best / K  # us per launch

    K = max(8, min(args.steps, 64))
    launches0 = _capi.launch_count()
    us_ours = timed(ours, K)
    n_launch = _capi.launch_count() - launches0
    us_lib = timed(library, K)
    peak, src = measured_peaks()
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import conv_oracle as co
        xc, wc, bc = x.cpu(), w.cpu(), b.cpu()
        torch.set_num_threads(os.cpu_count() or 1)
        co.conv0_relu(xc, wc, bc)
        t0 = time.perf_counter()
        co.conv0_relu(xc, wc, bc)
        cpu = {"value": audio_s / (time.perf_counter() - t0), "unit": "audio-s/s", "cores": os.cpu_count(), "kind": "port",
               "sample": "the whole batch once after one warm-up"}
    print(json.dumps({
        "metric": "audio-sec/sec through conv0+ReLU (row f2)", "value": audio_s / (us_ours * 1e-6), "unit": "audio-s/s",
        "n_gpus": 1, "steps": K, "warmup": max(3, args.warmup), "ms_per_step": us_ours * 1e-3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "aishell features 32x649x80 -> 32x32x324x78", "pool": "3 outputs of 103 MB (> L2)"},
        "roofline": {"bound": "hbm", "kernel": "conv0_relu_kernel", "achieved": alg_bytes / (us_ours * 1e-6) / 1e9, "peak": peak,
                     "unit": "GB/s", "frac": alg_bytes / (us_ours * 1e-6) / 1e9 / peak, "traffic": None,
                     "alg_bytes_per_launch": alg_bytes, "peak_source": src, "us_per_launch": us_ours},
        "library": {"impl": "torch conv2d + relu_ (cuDNN)", "us_per_launch": us_lib},
        "cpu_baseline": cpu, "gpu_launches": int(n_launch)}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=640)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--stage", default="frontend", choices=["frontend", "conv0"],
                    help="frontend: the SPLayer hot path (default, the contract line); conv0: row f2 on its features")
    ap.add_argument("--workload", default="aishell", choices=sorted(WORKLOADS))
    ap.add_argument("--shard", action="store_true", help="multi-GPU: split every batch over the ranks (BASELINE configs[3])")
    ap.add_argument("--epoch-hours", type=float, default=1000.0)
    ap.add_argument("--dither", type=float, default=1.0)
    ap.add_argument("--pool", type=int, default=32)
    ap.add_argument("--batches-per-launch", type=int, default=8)
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--no-numa-bind", action="store_true", help="multi-GPU: do not bind ranks to their GPU's NUMA node")
    ap.add_argument("--e2e-steps", type=int, default=96)
    ap.add_argument("--streams", type=int, default=1, help="launch lanes of the `value` graph (1 = strictly serial launches)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    if args.stage == "conv0" and args.impl == "ours":
        run_conv0(args)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
