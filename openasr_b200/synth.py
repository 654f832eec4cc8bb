"""Synthetic int16-scaled waveform batches of the named benchmark shapes."""
from __future__ import annotations

import math

import torch


def synth_batch(B: int, n_lo: int, n_hi: int, sample_rate: int, seed: int = 1234):
    """[B, Lmax] fp32 (integer valued, int16 range): 3000*randn + 2000*sin(2 pi f0 t), f0 ~ U[80, 400] Hz
    per utterance so spectra are not flat; ragged lengths n_i ~ U[n_lo, n_hi]; zero padded."""
    g = torch.Generator().manual_seed(seed)
    lengths = torch.randint(n_lo, n_hi + 1, (B,), generator=g)
    L = int(lengths.max())
    f0 = 80.0 + 320.0 * torch.rand(B, 1, generator=g)
    t = torch.arange(L, dtype=torch.float32).unsqueeze(0) / sample_rate
    x = 3000.0 * torch.randn(B, L, generator=g) + 2000.0 * torch.sin(2 * math.pi * f0 * t)
    x = x.clamp(-32768, 32767).round()
    mask = torch.arange(L).unsqueeze(0) < lengths.unsqueeze(1)
    return (x * mask).contiguous(), lengths
