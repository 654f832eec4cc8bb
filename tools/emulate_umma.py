"""CPU model (numpy) of the tcgen05 DFT-as-GEMM engine (openasr_b200/csrc/fbank_umma.cu), driven by the SAME
host-built tables the kernel uses (spl_debug_umma_tables: pre-swizzled FP16 hi/lo twiddle images, correction
chunk, delayed windows, streaming mel program).  Test / diagnostic infrastructure: tests/test_umma_model.py
checks it against the oracle, which pins the tables and the arithmetic of the kernel without a GPU.

Mirrors the kernel step by step:
  producer : pivot + power-of-two scale per 8-frame group, x~ = s (x - mu_c) (+ s d g), z' = w^(h) (x~_j - c x~_{j-1}),
             fold a = z'_j + z'_{N-j}, b = z'_j - z'_{N-j}, FP16 hi/lo split
  GEMM     : 3 products per block and K step, fp32 accumulation; correction chunk (-(1-c) mean, z'_{N/2})
  epilogue : power of bins 1..N/2-1 ascending, streaming mel program, 1/s^2, log
"""
import ctypes as C
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import frontend_oracle as fo  # noqa: E402  (the checker; this file is test infrastructure)

f32 = np.float32


def f16split(v):
    hi = v.astype(np.float16)
    lo = (v - hi.astype(f32)).astype(np.float16)
    return hi.astype(f32), lo.astype(f32)


def load_tables(sr, D, fmt=0, window="povey"):
    """(info, twiddle bytes, table floats) from the library's host-side builder."""
    from openasr_b200 import _capi, tables
    lib = _capi.load()
    S, Nw, N = tables.frame_geometry(sr)
    win = tables.window_table(window, Nw).to(torch.float32).contiguous()
    mel = tables.mel_table(D, N, sr).to(torch.float32).contiguous()
    info = (C.c_int32 * 16)()
    _capi.check(lib.spl_debug_umma_tables(N, Nw, D, C.c_void_p(win.data_ptr()), C.c_void_p(mel.data_ptr()), fmt,
                                          None, 0, None, 0, info), "spl_debug_umma_tables")
    if not info[0]:
        return None
    tw = np.zeros(info[1], np.uint8)
    tab = np.zeros(info[2], f32)
    _capi.check(lib.spl_debug_umma_tables(N, Nw, D, C.c_void_p(win.data_ptr()), C.c_void_p(mel.data_ptr()), fmt,
                                          tw.ctypes.data, tw.nbytes, tab.ctypes.data, tab.size, info),
                "spl_debug_umma_tables")
    return {"N": N, "Nw": Nw, "S": S, "D": D, "nshift": 4 if fmt == 0 else 8, "tw": tw, "tab": tab,
            "off_melw": info[3], "off_melc": info[4], "nflush": info[5], "nparts": info[6],
            "part_f0": [info[7 + i] for i in range(4)], "part_s0": [info[11 + i] for i in range(5)]}


def _tile(tw, off, rows):
    """De-swizzle one SWIZZLE_32B K-major FP16 tile [rows x 16] starting at byte `off`."""
    raw = tw[off:off + rows * 32].view(np.float16).reshape(rows, 16)
    out = np.empty((rows, 16), f32)
    for n in range(rows):
        sw = (n >> 2) & 1
        out[n, 0:8] = raw[n, 8 * sw:8 * sw + 8]
        out[n, 8:16] = raw[n, 8 * (1 - sw):8 * (1 - sw) + 8]
    return out


def emulate(wave, T, h=0, noise=None, dither=0.0, c=0.97, remove_dc=True, return_acc=False):
    """Log-mel features (m, D) of one utterance as the kernel computes them (shift h of every row)."""
    N, Nw, S, D = T["N"], T["Nw"], T["S"], T["D"]
    HALF, NB, NCH = N // 4, N // 2, N // 64
    nshift = T["nshift"]
    KX = (3 * nshift + 2 + 15) // 16
    tab, tw = T["tab"], T["tw"]
    wA = tab[h * N:(h + 1) * N]
    b_tile = HALF * 32
    b_stage = 4 * b_tile
    n = wave.shape[0]
    m = fo.num_frames(n, Nw, S)
    idx = np.arange(m)[:, None] * S + np.arange(Nw)[None, :]
    x = wave.astype(f32)[idx]
    # pivot / scale per 8-row group (the kernel scans the group's sample span)
    piv = np.zeros((m, 1), f32)
    sc = np.ones((m, 1), f32)
    for r0 in range(0, m, 8):
        r1 = min(r0 + 8, m)
        seg = wave[r0 * S:(r1 - 1) * S + Nw].astype(f32)
        pv = f32(seg.sum(dtype=f32) / f32(seg.size)) if remove_dc else f32(0)
        mx = float(np.abs(seg).max())
        bound = 4.0 * (2.0 * mx + 6.0 * abs(dither))
        e = 13 - (int(math.ceil(math.log2(bound))) if bound > 0 else 0)
        e = max(-60, min(60, e))
        piv[r0:r1] = pv
        sc[r0:r1] = f32(2.0 ** e)
    xt = (x * sc + (-piv * sc)).astype(f32)
    if dither != 0.0:
        xt = (xt + (sc * f32(dither)) * noise.astype(f32)).astype(f32)
    xs = np.zeros((m, N), f32)
    xs[:, h:h + Nw] = xt
    rowsum = xt.sum(axis=1, dtype=f32)
    prev = np.concatenate([xs[:, :1], xs[:, :-1]], axis=1)
    prev[:, h] = xs[:, h]
    z = (wA[None, :] * (xs - f32(c) * prev).astype(f32)).astype(f32)
    zj = z[:, :NB]
    zm = np.zeros_like(zj)
    zm[:, 1:] = z[:, :NB:-1][:, :NB - 1]
    a = (zj + zm).astype(f32)
    b = (zj - zm).astype(f32)
    acc = [np.zeros((m, HALF), f32) for _ in range(4)]  # ce co se so
    for ch in range(NCH):
        for hf in range(2):
            src = a if hf == 0 else b
            for pp in range(2):
                cols = 32 * ch + 2 * np.arange(16) + pp
                hi, lo = f16split(src[:, cols])
                t_hi = _tile(tw, (ch * 2 + hf) * b_stage + (2 * pp) * b_tile, HALF)
                t_lo = _tile(tw, (ch * 2 + hf) * b_stage + (2 * pp + 1) * b_tile, HALF)
                acc[2 * hf + pp] += (hi @ t_hi.T + lo @ t_hi.T + hi @ t_lo.T).astype(f32)
    # correction chunk
    g = (-(f32(1) - f32(c)) * rowsum / f32(Nw)).astype(f32) if remove_dc else np.zeros(m, f32)
    zh = z[:, NB]
    ghi, glo = f16split(g)
    zhi, zlo = f16split(zh)
    ax = np.zeros((m, 16 * KX), f32)
    ax[:, 3 * h + 0] = ghi
    ax[:, 3 * h + 1] = glo
    ax[:, 3 * h + 2] = ghi
    ax[:, 3 * nshift] = zhi
    ax[:, 3 * nshift + 1] = zlo
    for hf in range(2):
        for pp in range(2):
            for kx in range(KX):
                bx = _tile(tw, (NCH * 2 + hf) * b_stage + (kx * 2 + pp) * b_tile, HALF)
                acc[2 * hf + pp] += (ax[:, 16 * kx:16 * kx + 16] @ bx.T).astype(f32)
    ce, co, se, so = acc
    # epilogue: steps 0..HALF-1 -> bin step+1 (column = step); step HALF+i -> column HALF-1-i
    melw = tab[T["off_melw"]:T["off_melw"] + 2 * NB].reshape(NB, 2)
    melc = tab[T["off_melc"]:T["off_melc"] + NB // 16].view(np.uint32)
    out = np.zeros((m, D), f32)
    accA = np.zeros(m, f32)
    accB = np.zeros(m, f32)
    col = 0
    inv2 = (1.0 / (sc[:, 0].astype(np.float64) ** 2)).astype(f32)

    def emit():
        nonlocal accA, accB, col
        out[:, col] = np.log(np.maximum(accA * inv2, f32(fo.EPS)))
        accA, accB = accB, np.zeros(m, f32)
        col += 1

    for step in range(2 * HALF):
        if step < HALF:
            colm = step
            re, im = ce[:, colm] + co[:, colm], se[:, colm] + so[:, colm]
        else:
            colm = 2 * HALF - 1 - step
            re, im = ce[:, colm] - co[:, colm], so[:, colm] - se[:, colm]
        pw = (re * re + im * im).astype(f32)
        ns = (int(melc[step >> 4]) >> (2 * (step & 15))) & 3
        for _ in range(ns):
            emit()
        accA = (accA + melw[step, 0] * pw).astype(f32)
        accB = (accB + melw[step, 1] * pw).astype(f32)
    for _ in range(T["nflush"]):
        emit()
    assert col == D, (col, D)
    if return_acc:
        return out, np.concatenate(acc, axis=1), inv2
    return out


def report(name, got, ref32, ref64):
    d = np.abs(got - ref32)
    tol = 1e-3 + 1e-4 * np.abs(ref32)
    bad = d > tol
    wide = 2 * np.abs(ref32 - ref64)
    bad2 = d > tol + wide
    print("%-34s max|d| vs ref32 %.3e vs ref64 %.3e | ref32-ref64 %.3e | viol %d (widened %d) of %d" % (
        name, d.max(), np.abs(got - ref64).max(), np.abs(ref32 - ref64).max(), bad.sum(), bad2.sum(), d.size))
    return int(bad2.sum()), float(d.max())


if __name__ == "__main__":
    torch.manual_seed(0)
    G = os.path.join(ROOT, "tests", "golden")
    T16 = load_tables(16000, 80)
    for wi in (0, 1):
        for dc in (0.0, 3000.0):
            wav = np.load(os.path.join(G, "wav%d.npy" % wi)).astype(f32) + f32(dc)
            tw_ = torch.from_numpy(wav)
            r32 = fo.fbank(tw_, 16000.0, 80, dither=0.0).numpy()
            r64 = fo.fbank(tw_, 16000.0, 80, dither=0.0, dtype=torch.float64).numpy()
            for h in (0, 3):
                report("wav%d dc%d h=%d" % (wi, dc, h), emulate(wav, T16, h=h), r32, r64)
        wav = np.load(os.path.join(G, "wav%d.npy" % wi)).astype(f32)
        tw_ = torch.from_numpy(wav)
        m = fo.num_frames(wav.shape[0], 400, 160)
        noise = fo.dither_noise((m, 400))
        r32 = fo.fbank(tw_, 16000.0, 80, dither=1.0, noise=noise).numpy()
        r64 = fo.fbank(tw_, 16000.0, 80, dither=1.0, noise=noise, dtype=torch.float64).numpy()
        report("wav%d dither1 h=2" % wi, emulate(wav, T16, h=2, noise=noise.numpy(), dither=1.0), r32, r64)
    T8 = load_tables(8000, 40, fmt=1)
    w8 = np.load(os.path.join(G, "wav1.npy")).astype(f32)[::2]
    tw_ = torch.from_numpy(w8)
    r32 = fo.fbank(tw_, 8000.0, 40, dither=0.0).numpy()
    r64 = fo.fbank(tw_, 8000.0, 40, dither=0.0, dtype=torch.float64).numpy()
    report("8 kHz int16 tables h=5", emulate(w8, T8, h=5), r32, r64)
