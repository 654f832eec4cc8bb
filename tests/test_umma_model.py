"""CPU: the tcgen05 engine's host-built tables and arithmetic, through the numpy model of the kernel
(tools/emulate_umma.py) against the oracle.  Pins -- without a GPU -- the pre-swizzled FP16 hi/lo twiddle images,
the DC / Nyquist correction chunk, the delayed windows, the streaming mel program and the epilogue part table
that csrc/umma_tables.cu builds and csrc/fbank_umma.cu consumes."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle import frontend_oracle as fo  # noqa: E402


def _check(got, ref32, ref64):
    d = np.abs(got - ref32)
    tol = 1e-3 + 1e-4 * np.abs(ref32) + 2 * np.abs(ref32 - ref64)
    assert (d <= tol).all(), "max|d| %.3g" % d.max()
    return d.max()


@pytest.mark.parametrize("sr,D,fmt,h", [(16000, 80, 0, 0), (16000, 80, 0, 3), (8000, 40, 1, 5), (16000, 40, 1, 7)])
def test_model_matches_oracle(wavs, sr, D, fmt, h):
    import emulate_umma as em
    T = em.load_tables(sr, D, fmt)
    assert T is not None
    wav = wavs[0][::16000 // sr][:12000].contiguous()
    ref32 = fo.fbank(wav, float(sr), D, dither=0.0).numpy()
    ref64 = fo.fbank(wav, float(sr), D, dither=0.0, dtype=torch.float64).numpy()
    _check(em.emulate(wav.numpy(), T, h=h), ref32, ref64)


def test_model_dc_offset_and_dither(wavs):
    """DC removal folded into the GEMM (mean of the NOISY frame) and a large DC offset (pivot per 8 frames)."""
    import emulate_umma as em
    T = em.load_tables(16000, 80, 0)
    wav = (wavs[1][:16000] + 3000.0).contiguous()
    m = fo.num_frames(wav.shape[0], 400, 160)
    g = torch.Generator().manual_seed(3)
    noise = fo.dither_transform(torch.rand((m, 400), generator=g))
    ref32 = fo.fbank(wav, 16000.0, 80, dither=1.0, noise=noise).numpy()
    ref64 = fo.fbank(wav, 16000.0, 80, dither=1.0, noise=noise, dtype=torch.float64).numpy()
    _check(em.emulate(wav.numpy(), T, h=1, noise=noise.numpy(), dither=1.0), ref32, ref64)


def test_mel_program_and_parts():
    """The streaming mel program emits every filter exactly once, in order, and the epilogue parts start on the
    filter the sequential program is on at their first step."""
    import emulate_umma as em
    for sr, D in ((16000, 80), (8000, 40), (16000, 23), (16000, 128), (11025, 40)):
        T = em.load_tables(sr, D, 0)
        if T is None:
            continue
        NB = T["N"] // 2
        melc = T["tab"][T["off_melc"]:T["off_melc"] + NB // 16].view(np.uint32)
        shifts = [(int(melc[s >> 4]) >> (2 * (s & 15))) & 3 for s in range(NB)]
        assert sum(shifts) + T["nflush"] == D
        assert T["nparts"] in (1, 2, 4)
        s0 = T["part_s0"]
        assert s0[0] == 0 and s0[T["nparts"]] == NB and all(v % 8 == 0 for v in s0[:T["nparts"] + 1])
        f = 0
        for pt in range(T["nparts"]):
            assert T["part_f0"][pt] == f and s0[pt + 1] > s0[pt]
            e = sum(shifts[s0[pt]:s0[pt + 1]])
            assert pt == 0 or e >= 2
            f += e


def test_unsupported_configurations_fall_back():
    """use_energy and windows that leave no room for the alignment shift stay on the FFT engine."""
    import emulate_umma as em
    assert em.load_tables(20400, 40, 0) is None      # Nw = 510: 510 + 3 > 512, no room for the alignment shift
    assert em.load_tables(20000, 40, 1) is not None  # Nw = 500: even the 7 int16 shifts fit
