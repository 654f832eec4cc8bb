"""Staged bring-up of the tcgen05 engine on a GPU box: each stage in its own process (a CUDA fault must not
take the later stages with it).  Prints everything needed to localise a bug from one run.

    python tools/umma_debug.py            # all stages
    python tools/umma_debug.py acc|batch|modes|multi|time
"""
import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
G = os.path.join(ROOT, "tests", "golden")


def _layer(**kw):
    from openasr_b200 import SPLayer
    conf = {"feature_type": "fbank", "sample_rate": 16000, "num_mel_bins": 80, "use_energy": False, "dither": 0.0}
    conf.update(kw)
    return SPLayer(conf).cuda().eval(), conf


def _cmp(name, got, ref, ref64=None):
    import torch
    got, ref = got.detach().cpu().float(), ref.float()
    fin = torch.isfinite(got).all().item()
    d = (got - ref).abs()
    tol = 1e-3 + 1e-4 * ref.abs()
    if ref64 is not None:
        tol = tol + 2 * (ref.double() - ref64.double()).abs().float()
    bad = (d > tol)
    print("[%s] finite=%s max|d|=%.3e mean|d|=%.3e violations=%d/%d" % (name, fin, d[torch.isfinite(d)].max().item() if fin or torch.isfinite(d).any() else float("nan"),
                                                                     d[torch.isfinite(d)].mean().item() if torch.isfinite(d).any() else float("nan"), int(bad.sum()), d.numel()), flush=True)
    if bad.any():
        idx = bad.nonzero()[:6]
        for i in idx.tolist():
            print("    at", i, "got", got[tuple(i)].item(), "ref", ref[tuple(i)].item())
    return int(bad.sum()) == 0 and fin


def stage_acc():
    """One utterance, one CTA: raw accumulators of the first tile against the CPU model."""
    os.environ["SPL_UMMA_DEBUG"] = "1"
    os.environ["SPL_UMMA_CTAS"] = "1"
    os.environ["SPL_ENGINE"] = "umma"
    import numpy as np
    import torch
    from openasr_b200 import _capi
    from oracle import frontend_oracle as fo
    import emulate_umma as em
    wav = torch.from_numpy(np.load(os.path.join(G, "wav0.npy")).astype(np.float32))
    layer, conf = _layer()
    h = layer._handle(torch.device("cuda", 0))
    print("engine:", h.engine_name(), flush=True)
    x = wav[None, :].cuda().contiguous()
    t0 = time.time()
    feats, flen = layer(x, [wav.shape[0]])
    torch.cuda.synchronize()
    print("ran in %.3f s, status=0x%x, flen=%s" % (time.time() - t0, h.debug_status(), flen.tolist()), flush=True)
    ref = fo.fbank(wav, 16000.0, 80, dither=0.0)
    ok = _cmp("acc-stage features", feats[0], ref, fo.fbank(wav, 16000.0, 80, dither=0.0, dtype=torch.float64))
    T = em.load_tables(16000, 80)
    hsh = (x.data_ptr() % 16) // 4
    out_m, acc_m, inv2_m = em.emulate(wav.numpy(), T, h=hsh, return_acc=True)
    buf = np.zeros((128, 513), np.float32)
    _capi.check(h._lib.spl_debug_umma_acc(h._h, buf.ctypes.data, buf.size), "acc")
    acc_g, inv2_g = buf[:, :512], buf[:, 512]
    print("shift h=%d; 1/s^2 model %s gpu %s" % (hsh, inv2_m[:3], inv2_g[:3]))
    names = ("ce", "co", "se", "so")
    for b in range(4):
        dm = np.abs(acc_g[:, 128 * b:128 * b + 128] - acc_m[:128, 128 * b:128 * b + 128])
        sc = np.abs(acc_m[:128, 128 * b:128 * b + 128]).max()
        print("  block %s: max|d|=%.4g (model max %.4g) rows with error>1e-3*max: %d, cols: %d" % (
            names[b], dm.max(), sc, int((dm.max(1) > 1e-3 * sc).sum()), int((dm.max(0) > 1e-3 * sc).sum())), flush=True)
        if dm.max() > 1e-3 * sc:
            r, c = np.unravel_index(dm.argmax(), dm.shape)
            print("    worst at row %d col %d: gpu %.5g model %.5g" % (r, c, acc_g[r, 128 * b + c], acc_m[r, 128 * b + c]))
            print("    gpu row0[:8]  ", acc_g[0, 128 * b:128 * b + 8])
            print("    model row0[:8]", acc_m[0, 128 * b:128 * b + 8])
            print("    gpu row5[:8]  ", acc_g[5, 128 * b:128 * b + 8])
            print("    model row5[:8]", acc_m[5, 128 * b:128 * b + 8])
    if not ok:
        d = (feats[0].cpu() - ref).abs()
        print("  feature error by row (first 16):", d.max(1).values[:16].tolist())
        print("  feature error by col (first 16):", d.max(0).values[:16].tolist())
        dmod = np.abs(feats[0].cpu().numpy() - out_m)
        print("  vs CPU model: max|d| %.3e" % dmod.max())
    return ok


def stage_batch():
    """Ragged batch, misaligned rows, full grid; FFT engine beside it."""
    import numpy as np
    import torch
    from oracle import frontend_oracle as fo
    ws = [torch.from_numpy(np.load(os.path.join(G, "wav%d.npy" % i)).astype(np.float32)) for i in (0, 1)]
    ws = [ws[0], ws[1], ws[0][:7001].contiguous(), ws[1][:400].contiguous(), ws[1][100:900].contiguous()]
    lens = [w.shape[0] for w in ws]
    ok = True
    for pad in (0, 1, 2, 3):
        x = torch.zeros(len(ws), max(lens) + pad)
        for i, w in enumerate(ws):
            x[i, :lens[i]] = w
        for eng in ("umma", "fft"):
            os.environ["SPL_ENGINE"] = eng
            layer, conf = _layer()
            h = layer._handle(torch.device("cuda", 0))
            f, fl = layer(x.cuda(), lens)
            torch.cuda.synchronize()
            ref, rl = fo.splayer_forward(x, lens, conf)
            r64, _ = fo.splayer_forward(x, lens, conf, dtype=torch.float64)
            assert fl.cpu().tolist() == rl.tolist(), (fl, rl)
            ok &= _cmp("batch pad=%d %s status=0x%x" % (pad, h.engine_name(), h.debug_status()), f, ref, r64)
            pz = all((f[i, m:] == 0).all().item() for i, m in enumerate(rl.tolist()))
            print("    padding rows exactly zero:", pz)
            ok &= pz
    return ok


def stage_modes():
    """Host-noise parity mode, device dither statistics, int16 ingest, 8 kHz, CMVN stats."""
    import numpy as np
    import torch
    from oracle import frontend_oracle as fo
    os.environ["SPL_ENGINE"] = "umma"
    ws = [torch.from_numpy(np.load(os.path.join(G, "wav%d.npy" % i)).astype(np.float32)) for i in (0, 1)]
    lens = [w.shape[0] for w in ws]
    x = torch.zeros(2, max(lens) + 1)
    for i, w in enumerate(ws):
        x[i, :lens[i]] = w
    ok = True
    layer, conf = _layer(dither=1.0, dither_rng="host")
    torch.manual_seed(5)
    f, _ = layer(x.cuda(), lens)
    torch.manual_seed(5)
    r, _ = fo.splayer_forward(x, lens, conf)
    ok &= _cmp("host-noise dither", f, r)
    # int16
    layer, conf = _layer()
    f16, _ = layer(x.to(torch.int16).cuda(), lens)
    f32_, _ = layer(x.cuda(), lens)
    r, _ = fo.splayer_forward(x, lens, conf)
    ok &= _cmp("int16 ingest (%s)" % layer._handle(torch.device("cuda", 0)).engine_name(torch.int16), f16, r)
    print("    int16 vs fp32 path max|d| %.3e" % (f16 - f32_).abs().max().item())
    # 8 kHz
    x8 = x[:, ::2].contiguous()
    l8 = [(n + 1) // 2 for n in lens]
    layer, conf = _layer(sample_rate=8000, num_mel_bins=40)
    f, _ = layer(x8.cuda(), l8)
    r, _ = fo.splayer_forward(x8, l8, conf)
    r64, _ = fo.splayer_forward(x8, l8, conf, dtype=torch.float64)
    ok &= _cmp("8 kHz 40 mel", f, r, r64)
    # device dither: features of silence, umma vs fft
    z = torch.zeros(4, 48000)
    st = {}
    for eng in ("umma", "fft"):
        os.environ["SPL_ENGINE"] = eng
        layer, conf = _layer(dither=1.0)
        torch.manual_seed(1)
        f, _ = layer(z.cuda(), [48000] * 4)
        st[eng] = (f.mean(dim=(0, 1)).cpu(), f.std(dim=(0, 1)).cpu())
        print("    device dither on silence [%s]: mean[:4] %s std[:4] %s finite %s" % (
            eng, st[eng][0][:4].tolist(), st[eng][1][:4].tolist(), torch.isfinite(f).all().item()))
    dm = (st["umma"][0] - st["fft"][0]).abs().max().item()
    ds = (st["umma"][1] - st["fft"][1]).abs().max().item()
    print("    umma vs fft: max mean diff %.3f, max std diff %.3f" % (dm, ds))
    ok &= dm < 0.1 and ds < 0.1
    # CMVN + SpecAug through the module
    os.environ["SPL_ENGINE"] = "umma"
    sa = {"freq_mask_num": 2, "freq_mask_width": 27, "time_mask_num": 2, "time_mask_width": 40}
    layer, conf = _layer(cmvn="utterance", spec_aug=sa)
    layer.train()
    torch.manual_seed(11)
    f, fl = layer(x.cuda(), lens)
    torch.manual_seed(11)
    uni = torch.rand(8, 2)
    r, rl = fo.splayer_forward(x, lens, conf, training=True, specaug_uniforms=uni)
    d = (f.cpu() - r).abs().max().item()
    print("[cmvn+specaug] max|d| %.3e" % d)
    ok &= d < 5e-3
    return ok


def stage_multi():
    import torch
    from oracle import frontend_oracle as fo
    from openasr_b200 import tables
    os.environ["SPL_ENGINE"] = os.environ.get("MULTI_ENGINE", "umma")
    layer, conf = _layer()
    dev = torch.device("cuda", 0)
    h = layer._handle(dev)
    print("engine", h.engine_name())
    items, singles = [], []
    for k in range(5):
        wav, lens = fo.synth_batch(6 + k, 3000, 50000, 16000, seed=100 + k)
        frames = [tables.frame_count(int(n), h.win, h.shift) for n in lens.tolist()]
        T = max(frames)
        it = {"wav": wav.cuda(), "lens": lens.cuda(), "T": T,
              "feats": torch.full((wav.shape[0], T, 80), float("nan"), device=dev),
              "flen": torch.zeros(wav.shape[0], dtype=torch.int64, device=dev),
              "stats": torch.empty((wav.shape[0], 2, 80), dtype=torch.float64, device=dev)}
        items.append(it)
        singles.append(layer(wav.cuda(), lens)[0])
    gst = torch.zeros(161, dtype=torch.float64, device=dev)
    h.fbank_multi(items, global_stats=gst)
    torch.cuda.synchronize()
    print("multi status=0x%x" % h.debug_status())
    ok = True
    tot = 0
    s1 = torch.zeros(80, dtype=torch.float64)
    for k, it in enumerate(items):
        same = torch.equal(it["feats"], singles[k])
        md = (it["feats"] - singles[k]).abs().max().item()
        fl = it["flen"].cpu()
        st = it["stats"].cpu()
        f = it["feats"].cpu().double()
        want = torch.stack([torch.stack([f[i, :fl[i]].sum(0), (f[i, :fl[i]] ** 2).sum(0)]) for i in range(f.shape[0])])
        es = ((st - want).abs() / (1 + want.abs())).max().item()
        tot += int(fl.sum())
        s1 += want[:, 0].sum(0)
        print("  batch %d: identical to single launch: %s (max|d| %.2e); utt_stats rel err %.2e" % (k, same, md, es))
        ok &= md < 1e-5 and es < 1e-9
    g = gst.cpu()
    print("  global: count %d (want %d), sum rel err %.2e" % (int(g[160]), tot, ((g[:80] - s1).abs() / (1 + s1.abs())).max().item()))
    ok &= int(g[160]) == tot
    return ok


def stage_time():
    import torch
    from oracle import frontend_oracle as fo
    from openasr_b200 import tables
    dev = torch.device("cuda", 0)
    for dither in (0.0, 1.0):
        for eng in ("umma", "fft"):
            os.environ["SPL_ENGINE"] = eng
            layer, conf = _layer(dither=dither)
            h = layer._handle(dev)
            items = []
            for k in range(16):
                wav, lens = fo.synth_batch(32, 56000, 104000, 16000, seed=1234 + k)
                frames = [tables.frame_count(int(n), h.win, h.shift) for n in lens.tolist()]
                T = max(frames)
                items.append({"wav": wav.cuda(), "lens": lens.cuda(), "T": T,
                              "feats": torch.empty((32, T, 80), device=dev), "flen": torch.zeros(32, dtype=torch.int64, device=dev),
                              "stats": torch.empty((32, 2, 80), dtype=torch.float64, device=dev), "audio": float(lens.sum()) / 16000})
            for K in (1, 2, 4, 8):
                def run():
                    for i in range(0, 16, K):
                        h.fbank_multi(items[i:i + K], dither_seed=7)
                for _ in range(3):
                    run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                best = 1e9
                for _ in range(5):
                    e0.record()
                    run()
                    e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                print("dither %.0f engine %-5s K=%d: %.2f us per batch (status 0x%x)" % (dither, h.engine_name(), K, 1e3 * best / 16, h.debug_status()), flush=True)


STAGES = {"acc": stage_acc, "batch": stage_batch, "modes": stage_modes, "multi": stage_multi, "time": stage_time}

if __name__ == "__main__":
    if len(sys.argv) > 1:
        r = STAGES[sys.argv[1]]()
        print("STAGE %s -> %s" % (sys.argv[1], r), flush=True)
        sys.exit(0 if r is not False else 1)
    for name in STAGES:
        print("=" * 30, name, flush=True)
        p = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=600)
        print("stage %s exit code %d" % (name, p.returncode), flush=True)
