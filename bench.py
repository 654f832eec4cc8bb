#!/usr/bin/env python
"""bench.py -- audio-seconds/second of the online speech front-end (fbank + CMVN + SpecAug).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload aishell|hkust|libri]

A "step" is one pass of the hot path (kernel A fused fbank + kernel B CMVN/SpecAug) over one
synthetic batch of the named shape.  Default workload = BASELINE.json configs[1], the AISHELL-1
shape: 32 x ~5 s, 16 kHz, 80-dim fbank, reference-default dither (1.0, device RNG), utterance
CMVN, SpecAug 2x27 F + 2x40 T, training mode.

  value     device-resident throughput: inputs already in HBM, a pool of distinct batches larger
            than L2 cycled under CUDA graphs, timed with CUDA events, max over ranks.
  e2e       the same metric through the public module API (SPLayer.forward) with HOST buffers:
            pinned wav -> H2D, forward (host RNG draws included), D2H of the features.
  roofline  the dominant kernel (fbank_kernel) timed alone over the same pool; algorithmic bytes
            = 4*sum n_i + 4*B*T*D + 16*B per launch against MEASURED_PEAKS hbm_gbs.
  cpu_baseline  the oracle port of the reference's CPU path on this box's host cores (N=1 only).

With --impl reference the oracle port (the reference is pure Python, there is no oracle/_ref
binary) is timed on the host cores with all threads; each step is a bounded sample of the workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

# dram__bytes_read.sum + dram__bytes_write.sum of fbank_warp_kernel per launch (ncu --set full,
# profiles/r1_summary.md); the feature writes are still L2-resident when the kernel retires
NCU_TRAFFIC = {"aishell": 9.88e6}  # prof_w5_d1: 9.878 MB read + 0 written (features leave L2 after the launch)

WORKLOADS = {
    # name: (B, n_lo, n_hi, sample_rate, D, cmvn, spec_aug)
    "aishell": (32, 56000, 104000, 16000, 80, "utterance",
                {"freq_mask_num": 2, "freq_mask_width": 27, "time_mask_num": 2, "time_mask_width": 40}),
    "hkust": (64, 16000, 48000, 8000, 40, "utterance", None),
    "libri": (16, 192000, 320000, 16000, 80, "none",
              {"freq_mask_num": 2, "freq_mask_width": 27, "time_mask_num": 2, "time_mask_width": 100}),
}


def workload_config(name, dither):
    B, lo, hi, sr, D, cmvn, sa = WORKLOADS[name]
    conf = {"feature_type": "fbank", "sample_rate": sr, "num_mel_bins": D, "use_energy": False,
            "dither": dither, "cmvn": cmvn, "dither_rng": "device", "specaug_rng": "host"}
    if sa is not None:
        conf["spec_aug"] = dict(sa)
    return conf


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


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
    B, lo, hi, sr, D, cmvn, sa = WORKLOADS[args.workload]
    conf = workload_config(args.workload, args.dither)
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
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "batch": B, "sample_rate": sr, "num_mel_bins": D, "cmvn": cmvn,
                   "spec_aug": sa, "dither": args.dither, "training": True},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def build_pool(layer, conf, wl, pool, dev, seed):
    """Device-resident pool of distinct batches + everything a step needs (lengths, mask rectangles,
    output buffers), so the timed region holds only the hot-path launches."""
    from openasr_b200 import frontend, tables
    from openasr_b200.synth import synth_batch
    B, lo, hi, sr, D, cmvn, sa = WORKLOADS[wl]
    h = layer._handle(dev)
    items = []
    for i in range(pool):
        wav, lens = synth_batch(B, lo, hi, sr, seed=seed + 7919 * i)
        frames = [tables.frame_count(int(n), h.win, h.shift) for n in lens.tolist()]
        T = max(frames)
        it = {
            "wav_host": wav, "lens_host": lens, "wav": wav.to(dev), "lens": lens.to(dev), "T": T, "frames": frames,
            "feats": torch.empty((B, T, h.d_out), dtype=torch.float32, device=dev),
            "flen": torch.empty((B,), dtype=torch.int64, device=dev),
            "stats": torch.empty((B, 2, h.d_out), dtype=torch.float64, device=dev),
            "audio_s": float(lens.sum()) / sr,
            "alg_bytes": 4 * int(lens.sum()) + 4 * B * T * h.d_out + 16 * B,
            "rect": None,
        }
        if sa is not None:
            u = frontend.specaug_uniforms(B, sa["freq_mask_num"], sa["time_mask_num"])
            it["rect"] = frontend.specaug_rectangles(u, torch.tensor(frames), T, h.d_out, sa).to(dev)
        items.append(it)
    return h, items


def step_resident(h, it, conf, seed, only_a=False):
    from openasr_b200 import frontend
    sa = conf.get("spec_aug")
    need_stats = conf["cmvn"] == "utterance" or (sa is not None and sa["time_mask_num"] > 0)
    h.fbank(it["wav"], it["lens"], it["T"], dither_seed=seed, utt_stats=it["stats"] if need_stats else None,
            out=it["feats"], feat_len=it["flen"])
    if only_a:
        return 1
    if conf["cmvn"] != "none" or sa is not None:
        frontend.post_inplace(it["feats"], it["flen"], cmvn_mode=conf["cmvn"], utt_stats=it["stats"] if need_stats else None,
                              mask_params=it["rect"], n_freq=sa["freq_mask_num"] if sa else 0,
                              n_time=sa["time_mask_num"] if sa else 0)
        return 2
    return 1


def time_graphed(fn_step, n_steps, chunk, stream, side_streams=()):
    """Capture `chunk` consecutive steps into a CUDA graph, replay to cover n_steps, time with events.
    With side streams, consecutive steps (independent batches) alternate over the streams inside the
    graph (fork/join), so the tail of one batch overlaps the head of the next."""
    chunk = max(1, min(chunk, n_steps))
    reps, rem = divmod(n_steps, chunk)
    graphs = []
    lanes = [stream] + list(side_streams)
    with torch.cuda.stream(stream):
        for count in ([chunk] if reps else []) + ([rem] if rem else []):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for s in lanes[1:]:
                    s.wait_stream(stream)
                for i in range(count):
                    with torch.cuda.stream(lanes[i % len(lanes)]):
                        fn_step(i)
                for s in lanes[1:]:
                    stream.wait_stream(s)
            graphs.append((g, count))
    return graphs, reps, rem


def run_ours(args):
    import torch.distributed as dist
    from openasr_b200 import SPLayer, _capi

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
    wl = args.workload
    B, lo, hi, sr, D, cmvn, sa = WORKLOADS[wl]
    conf = workload_config(wl, args.dither)
    layer = SPLayer(conf).to(dev).train()
    h, items = build_pool(layer, conf, wl, args.pool, dev, seed=1234 + 100003 * rank)
    pool_bytes = sum(it["wav"].numel() * 4 + it["feats"].numel() * 4 for it in items)
    K, W = args.steps, args.warmup
    stream = torch.cuda.Stream(device=dev)
    launches_per_step = [0]

    def step(i, only_a=False):
        launches_per_step[0] = step_resident(h, items[i % len(items)], conf, 0x9E3779B97F4A7C15 + i, only_a)

    # ---- warm-up (eager, also configures the kernels' shared-memory attributes) ----
    with torch.cuda.stream(stream):
        for i in range(max(W, 3)):
            step(i)
    stream.synchronize()
    side = [torch.cuda.Stream(device=dev) for _ in range(max(0, args.streams - 1))]
    graphs, reps, rem = time_graphed(step, K, args.graph_chunk, stream, side)
    lps = launches_per_step[0]
    with torch.cuda.stream(stream):
        for g, _ in graphs:  # graph warm-up
            g.replay()
    stream.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = None
    for _ in range(args.repeats):
        barrier()
        with torch.cuda.stream(stream):
            e0.record(stream)
            if reps:
                for _r in range(reps):
                    graphs[0][0].replay()
            if rem:
                graphs[-1][0].replay()
            e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    ms_total = best
    if sampler is not None:  # same load, untimed, so that nvidia-smi (100 ms period) sees the clocks under load
        t_end = time.perf_counter() + 0.7
        while time.perf_counter() < t_end:
            with torch.cuda.stream(stream):
                for _r in range(8):
                    graphs[0][0].replay()
            stream.synchronize()
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    audio_total = sum(items[i % len(items)]["audio_s"] for i in range(K))
    if world > 1:
        t = torch.tensor([audio_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        audio_total = float(t.item())
    value = audio_total / (ms_total * 1e-3)

    # ---- roofline: kernel A alone over the same pool ----
    def step_a(i):
        step(i, only_a=True)
    ga, reps_a, rem_a = time_graphed(step_a, K, args.graph_chunk, stream)
    with torch.cuda.stream(stream):
        ga[0][0].replay()
    stream.synchronize()
    torch.cuda.synchronize(dev)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _r in range(reps_a):
            ga[0][0].replay()
        if rem_a:
            ga[-1][0].replay()
        e1.record(stream)
    torch.cuda.synchronize(dev)
    ms_a = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    alg_bytes = sum(items[i % len(items)]["alg_bytes"] for i in range(K)) / K
    peak, peak_src = measured_peaks()
    achieved = alg_bytes / (ms_a * 1e-3 / K) / 1e9
    roofline = {"bound": "hbm", "kernel": "fbank_warp_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": NCU_TRAFFIC.get(wl), "peak_source": peak_src,
                "us_per_launch": 1e3 * ms_a / K, "alg_bytes_per_launch": alg_bytes,
                "step_share": ms_a / ms_total if world == 1 else None}

    # ---- e2e through SPLayer.forward with host buffers (every rank; max over ranks) ----
    e2e = run_e2e(layer, items, dev, max(8, min(K, args.e2e_steps)), world)
    e2e_i16 = run_e2e(layer, items, dev, max(8, min(K, args.e2e_steps)), world, int16=True)

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu_baseline = run_cpu_baseline(conf, items, sr) if (world == 1 and not args.no_cpu_baseline) else None
    line = {
        "metric": "audio-sec/sec featurized (fbank+CMVN+SpecAug)", "value": value, "unit": "audio-s/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl, "batch": B, "sample_rate": sr, "num_mel_bins": D, "cmvn": cmvn, "spec_aug": sa,
                   "dither": args.dither, "dither_rng": "device", "training": True,
                   "pool_batches": len(items), "pool_bytes": pool_bytes, "l2_flush": "pool larger than L2 (126 MB)",
                   "cuda_graph_chunk": args.graph_chunk, "batches_in_flight": args.streams, "timing": "best of %d regions of K steps, CUDA events" % args.repeats, "partition": "by utterance, %d rank(s), no data-path collective" % world,
                   "numa_bound_cpus": len(numa_cpus) if numa_cpus else None},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "e2e_int16_ingest": e2e_i16,
        "gpu_launches": lps * K, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(layer, items, dev, steps, world, int16=False):
    """SPLayer.forward from pinned host buffers: H2D of the fp32 wav batch, forward (host RNG draws,
    lengths/mask uploads, both kernels), D2H of the features; 3 slots in flight on 3 streams."""
    import torch.distributed as dist
    nslot = 3
    streams = [torch.cuda.Stream(device=dev) for _ in range(nslot)]
    maxL = max(it["wav_host"].shape[1] for it in items)
    maxT = max(it["T"] for it in items)
    B = items[0]["wav_host"].shape[0]
    d_out = items[0]["feats"].shape[2]
    pin_in = [(it["wav_host"].to(torch.int16) if int16 else it["wav_host"]).pin_memory() for it in items]
    pin_out = [torch.empty((B * maxT * d_out,), dtype=torch.float32).pin_memory() for _ in range(nslot)]
    pin_len = [torch.empty((B,), dtype=torch.int64).pin_memory() for _ in range(nslot)]
    done = [None] * nslot
    h2d = sum(p.numel() * p.element_size() for p in pin_in) / len(pin_in) + 8 * B
    d2h = sum(it["feats"].numel() * 4 for it in items) / len(items) + 8 * B

    def one(i):
        s = i % nslot
        it = items[i % len(items)]
        if done[s] is not None:
            done[s].synchronize()
        with torch.cuda.stream(streams[s]):
            wav = pin_in[i % len(items)].to(dev, non_blocking=True)
            feats, flen = layer(wav, it["lens_host"])
            pin_out[s][:feats.numel()].view_as(feats).copy_(feats, non_blocking=True)
            pin_len[s].copy_(flen, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(streams[s])
            done[s] = ev

    for i in range(4):
        one(i)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        one(i)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    audio = sum(items[i % len(items)]["audio_s"] for i in range(steps))
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        t = torch.tensor([audio], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        audio = float(t.item())
    return {"value": audio / dt, "unit": "audio-s/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "steps": steps, "api": "SPLayer.forward(pinned %s wav -> cuda, host lengths) + D2H of feats"
                                   % ("int16 PCM" if int16 else "fp32")}


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


def run_conv0(args):
    """Row f2 (conv0 + ReLU on the front-end's features), same rules as the main arm: CUDA events,
    >= 3 warm-ups, outputs cycled over a pool larger than L2, one JSON line.  `value` = audio-seconds/s
    through the layer at the AISHELL shape; roofline = HBM (the layer writes C*D1/(2D) = 15.6x its
    input); `library` = the same op through torch / cuDNN; `cpu_baseline` = the oracle port."""
    import ctypes as Ct
    from openasr_b200 import _capi
    dev = torch.device("cuda", 0)
    B, T, D, C = 32, 649, 80, 32
    audio_s = 32 * 5.0  # the AISHELL batch these features come from (mean 5 s per utterance)
    gen = torch.Generator().manual_seed(0)
    x = (4.0 * torch.randn(B, T, D, generator=gen) + 8.0).to(dev)
    w = (0.3 * torch.randn(C, 1, 3, 3, generator=gen)).to(dev)
    b = (0.1 * torch.randn(C, generator=gen)).to(dev)
    T1, D1 = (T - 3) // 2 + 1, D - 2
    alg_bytes = 4 * B * T * D + 4 * B * C * T1 * D1
    lib = _capi.load()
    outs = [torch.empty((B, C, T1, D1), device=dev) for _ in range(3)]  # 3 x 103 MB > L2
    stream = torch.cuda.Stream(device=dev)

    def ours(i):
        o = outs[i % len(outs)]
        _capi.check(lib.spl_conv0_relu(None, Ct.c_void_p(x.data_ptr()), B, T, D, Ct.c_void_p(w.data_ptr()),
                                       Ct.c_void_p(b.data_ptr()), C, Ct.c_void_p(o.data_ptr()),
                                       Ct.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))

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
    ap.add_argument("--dither", type=float, default=1.0)
    ap.add_argument("--pool", type=int, default=16)
    ap.add_argument("--graph-chunk", type=int, default=64)
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--no-numa-bind", action="store_true", help="multi-GPU: do not bind ranks to their GPU's NUMA node")
    ap.add_argument("--streams", type=int, default=4, help="independent batches in flight inside the graph")
    ap.add_argument("--e2e-steps", type=int, default=96)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
