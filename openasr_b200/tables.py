"""Host-side constant tables of the front-end, built with torch on the CPU.

The window and the mel bank are produced with the same torch operations, in the same
order and dtype as the reference (``src/third_party/kaldi_signal.py:109-128`` and
``:389-455``), so the uploaded tables are bit-identical to what the reference
rebuilds on every call; the CUDA handle keeps them resident instead.
"""
from __future__ import annotations

import math
from typing import Tuple

import torch

WINDOW_TYPES = ("povey", "hamming", "hanning", "rectangular", "blackman")


def frame_geometry(sample_rate: float) -> Tuple[int, int, int]:
    """(shift, window, padded window) -- kaldi_signal.py:150-152 with 10 ms / 25 ms frames."""
    shift = int(sample_rate * 10.0 * 0.001)
    win = int(sample_rate * 25.0 * 0.001)
    padded = 1 if win == 0 else 2 ** (win - 1).bit_length()
    return shift, win, padded


def frame_count(num_samples: int, win: int, shift: int) -> int:
    """kaldi_signal.py:86-90 (snip_edges=True)."""
    return 0 if num_samples < win else 1 + (num_samples - win) // shift


def window_table(window_type: str, win: int) -> torch.Tensor:
    """kaldi_signal.py:109-128."""
    if window_type == "povey":
        return torch.hann_window(win, periodic=False).pow(0.85)
    if window_type == "hamming":
        return torch.hamming_window(win, periodic=False, alpha=0.54, beta=0.46)
    if window_type == "hanning":
        return torch.hann_window(win, periodic=False)
    if window_type == "rectangular":
        return torch.ones(win)
    if window_type == "blackman":
        a = 2 * math.pi / (win - 1)
        n = torch.arange(win, dtype=torch.float32)
        return 0.42 - 0.5 * torch.cos(a * n) + (0.5 - 0.42) * torch.cos(2 * a * n)
    raise ValueError("Invalid window type " + str(window_type))


def mel_table(num_bins: int, padded: int, sample_rate: float, low_freq: float = 20.0,
              high_freq: float = 0.0) -> torch.Tensor:
    """Dense (num_bins, padded // 2) triangular bank -- kaldi_signal.py:389-455, vtln_warp = 1."""
    if num_bins <= 3:
        raise ValueError("Must have at least 3 mel bins")
    nyquist = 0.5 * sample_rate
    if high_freq <= 0.0:
        high_freq += nyquist
    if not (0.0 <= low_freq < nyquist and 0.0 < high_freq <= nyquist and low_freq < high_freq):
        raise ValueError("Bad values in options: low-freq %f and high-freq %f vs. nyquist %f"
                         % (low_freq, high_freq, nyquist))

    def mel_s(f):
        return 1127.0 * math.log(1.0 + f / 700.0)

    bin_width = sample_rate / padded
    lo, hi = mel_s(low_freq), mel_s(high_freq)
    delta = (hi - lo) / (num_bins + 1)
    idx = torch.arange(num_bins, dtype=torch.float32).unsqueeze(1)
    left = lo + idx * delta
    center = lo + (idx + 1.0) * delta
    right = lo + (idx + 2.0) * delta
    mel = (1127.0 * (1.0 + (bin_width * torch.arange(padded // 2, dtype=torch.float32)) / 700.0).log()).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    return torch.max(torch.zeros(1), torch.min(up, down)).contiguous()
