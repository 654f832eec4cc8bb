"""CPU oracle for the OpenASR online speech front-end (TEST INFRASTRUCTURE ONLY).

This module is a plain-torch CPU restatement of the reference's
``SPLayer.forward`` -> ``kaldi_signal.fbank`` -> ``SPLayer.spec_aug`` path.  It is
the *checker* for the CUDA product in ``openasr_b200/``; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product never routes through this file.

Pinning status
--------------
* fbank / SpecAug: **pinned** against the unmodified reference source run in the
  build container (``oracle/ref_shim.py`` imports ``/root/reference`` with an
  external ``torch.rfft`` shim) -- see ``oracle/make_golden.py`` and the committed
  vectors under ``tests/golden/``.  The reference's own tests hold no golden
  vectors or assertions for this path (``test/sp_layers_test.py`` only prints).
* CMVN: **parity unpinned** -- the reference has no CMVN code at all (only the
  ``subtract_mean`` hook, ``kaldi_signal.py:214-220``).  The semantics are the
  extension fixed in SURVEY.md section 5; the oracle for it is fp64 torch.

Every function cites the reference ``file:line`` it restates
(paths relative to ``/root/reference/src``).  The op sequence deliberately
mirrors the reference (including the broadcast multiply + sum mel projection) so
that timing this module is a fair stand-in for the reference's CPU path.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch

# third_party/kaldi_signal.py:48 -- numeric_limits<float>::epsilon()
EPS = float(torch.finfo(torch.float32).eps)

WINDOW_TYPES = ("povey", "hamming", "hanning", "rectangular", "blackman")


# --------------------------------------------------------------------------- sizes
def frame_params(sample_rate: float, frame_length_ms: float = 25.0,
                 frame_shift_ms: float = 10.0) -> Tuple[int, int, int]:
    """(shift S, window Nw, padded Nfft).  third_party/kaldi_signal.py:150-152, :61-64.

    Evaluated in Python floats exactly like the reference:
    ``int(sample_frequency * frame_shift * 0.001)``.
    """
    shift = int(sample_rate * frame_shift_ms * 0.001)
    win = int(sample_rate * frame_length_ms * 0.001)
    nfft = 1 if win == 0 else 2 ** (win - 1).bit_length()
    return shift, win, nfft


def num_frames(num_samples: int, win: int, shift: int) -> int:
    """snip_edges=True frame count.  third_party/kaldi_signal.py:86-90."""
    if num_samples < win:
        return 0
    return 1 + (num_samples - win) // shift


# --------------------------------------------------------------------------- tables
def window_function(window_type: str, win: int, blackman_coeff: float = 0.42,
                    dtype=torch.float32) -> torch.Tensor:
    """third_party/kaldi_signal.py:109-128."""
    if window_type == "hanning":
        return torch.hann_window(win, periodic=False, dtype=dtype)
    if window_type == "hamming":
        return torch.hamming_window(win, periodic=False, alpha=0.54, beta=0.46, dtype=dtype)
    if window_type == "povey":
        return torch.hann_window(win, periodic=False, dtype=dtype).pow(0.85)
    if window_type == "rectangular":
        return torch.ones(win, dtype=dtype)
    if window_type == "blackman":
        a = 2 * math.pi / (win - 1)
        n = torch.arange(win, dtype=dtype)
        return blackman_coeff - 0.5 * torch.cos(a * n) + (0.5 - blackman_coeff) * torch.cos(2 * a * n)
    raise ValueError("Invalid window type " + str(window_type))


def _mel(freq):
    """third_party/kaldi_signal.py:293-299."""
    if isinstance(freq, torch.Tensor):
        return 1127.0 * (1.0 + freq / 700.0).log()
    return 1127.0 * math.log(1.0 + freq / 700.0)


def mel_banks(num_bins: int, nfft: int, sample_rate: float, low_freq: float = 20.0,
              high_freq: float = 0.0, dtype=torch.float32) -> torch.Tensor:
    """Triangular mel filters, shape (num_bins, nfft // 2); vtln_warp == 1.0 branch.

    third_party/kaldi_signal.py:389-455 (get_mel_banks) with the arguments
    fbank passes at :530-531.
    """
    assert num_bins > 3, "Must have at least 3 mel bins"
    assert nfft % 2 == 0
    n_fft_bins = nfft // 2
    nyquist = 0.5 * sample_rate
    if high_freq <= 0.0:
        high_freq += nyquist
    assert 0.0 <= low_freq < nyquist and 0.0 < high_freq <= nyquist and low_freq < high_freq
    bin_width = sample_rate / nfft
    mel_lo = _mel(low_freq)
    mel_hi = _mel(high_freq)
    delta = (mel_hi - mel_lo) / (num_bins + 1)
    b = torch.arange(num_bins, dtype=dtype).unsqueeze(1)
    left = mel_lo + b * delta
    center = mel_lo + (b + 1.0) * delta
    right = mel_lo + (b + 2.0) * delta
    mel = _mel(bin_width * torch.arange(n_fft_bins, dtype=dtype)).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    return torch.max(torch.zeros(1, dtype=dtype), torch.min(up, down))


# --------------------------------------------------------------------------- dither
def dither_noise(shape, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """The reference's one-uniform pseudo Box-Muller noise, drawn on the CPU generator.

    third_party/kaldi_signal.py:176-177: ``x = max(eps, rand(shape))``;
    ``sqrt(-2 ln x) * cos(2 pi x)`` -- the SAME uniform feeds both factors.
    """
    u = torch.rand(shape, generator=generator)
    return dither_transform(u)


def dither_transform(u: torch.Tensor) -> torch.Tensor:
    x = torch.max(torch.tensor(EPS, dtype=u.dtype), u)
    return torch.sqrt(-2 * x.log()) * torch.cos(2 * math.pi * x)


# --------------------------------------------------------------------------- fbank
def fbank(wave: torch.Tensor, sample_rate: float = 16000.0, num_mel_bins: int = 23,
          use_energy: bool = False, dither: float = 1.0, window_type: str = "povey",
          preemph: float = 0.97, remove_dc: bool = True, noise: Optional[torch.Tensor] = None,
          dtype=torch.float32) -> torch.Tensor:
    """Log-mel filterbank of ONE utterance: (n,) -> (m, num_mel_bins + use_energy).

    Restates third_party/kaldi_signal.py:458-552 (fbank) + :163-211 (_get_window)
    + :67-106 (_get_strided, snip_edges=True) with the defaults SPLayer leaves in
    place (blocks/sp_layers.py:40-46): raw_energy, round_to_power_of_two,
    low_freq 20, high_freq Nyquist, energy_floor 0, htk_compat False, use_power,
    use_log_fbank, subtract_mean False.

    ``noise``: optional (m, Nw) tensor used INSTEAD of drawing -- this is the
    reference's ``rand_gauss`` (:177) so tests can feed identical noise to the
    oracle and to the CUDA path.  With ``noise=None`` and ``dither != 0`` the
    noise is drawn from the CPU default generator exactly like the reference.
    ``dtype=torch.float64`` gives the high-precision variant used to bound the
    fp32 self-error.
    """
    assert wave.dim() == 1
    shift, win, nfft = frame_params(sample_rate)
    n = wave.shape[0]
    # :154 -- the reference asserts 2 <= window_size <= len(waveform)
    assert 2 <= win <= n, "choose a window size %d that is [2, %d]" % (win, n)
    assert 0.0 <= preemph <= 1.0
    wave = wave.to(dtype)
    m = num_frames(n, win, shift)
    frames = wave.as_strided((m, win), (shift * wave.stride(0), wave.stride(0)))  # :105-106

    if dither != 0.0:  # :174-178
        if noise is None:
            noise = dither_noise(frames.shape)
        frames = frames + noise.to(dtype) * dither

    if remove_dc:  # :180-183
        frames = frames - torch.mean(frames, dim=1).unsqueeze(1)

    # :185-188, :131-140 (raw_energy=True, energy_floor=0)
    eps = torch.tensor(EPS, dtype=dtype)
    log_energy = torch.max(frames.pow(2).sum(1), eps).log()

    if preemph != 0.0:  # :190-194 -- replicate-pad on the left, then x[j] - c*x[j-1]
        prev = torch.nn.functional.pad(frames.unsqueeze(0), (1, 0), mode="replicate").squeeze(0)
        frames = frames - preemph * prev[:, :-1]

    frames = frames * window_function(window_type, win, dtype=dtype).unsqueeze(0)  # :197-199
    if nfft != win:  # :202-205
        frames = torch.nn.functional.pad(frames.unsqueeze(0), (0, nfft - win), mode="constant",
                                         value=0).squeeze(0)

    # :523-525 -- torch.rfft(x, 1) (removed in torch>=1.8) == view_as_real(fft.rfft)
    spec = torch.view_as_real(torch.fft.rfft(frames, dim=-1))
    power = spec.pow(2).sum(2).unsqueeze(1)  # (m, 1, nfft/2+1)

    # :530-537 -- zero column for the Nyquist bin, broadcast multiply + sum
    banks = mel_banks(num_mel_bins, nfft, sample_rate, dtype=dtype)
    banks = torch.nn.functional.pad(banks, (0, 1), mode="constant", value=0).unsqueeze(0)
    mel = (power * banks).sum(dim=2)
    mel = torch.max(mel, eps).log()  # :538-540

    if use_energy:  # :543-549, htk_compat=False -> energy is the FIRST column
        mel = torch.cat((log_energy.unsqueeze(1), mel), dim=1)
    return mel


# --------------------------------------------------------------------------- CMVN (extension)
def cmvn_stats(feats: torch.Tensor, lengths: Sequence[int]) -> torch.Tensor:
    """fp64 (3, D) = [sum x, sum x^2, count] over valid frames."""
    D = feats.shape[-1]
    out = torch.zeros(3, D, dtype=torch.float64)
    for i, l in enumerate(lengths):
        x = feats[i, :l].double()
        out[0] += x.sum(0)
        out[1] += (x * x).sum(0)
        out[2] += float(l)
    return out


def cmvn_apply(feats: torch.Tensor, lengths: Sequence[int], mode: str, norm_vars: bool = True,
               global_stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """CMVN over valid frames only; padding stays exactly 0.  fp64 math, fp32 result.

    Extension (not in the reference; **parity unpinned**).  ``mode``: 'none' |
    'utterance' | 'global'.  var floor 1e-20 as in Kaldi apply-cmvn.  With
    ``norm_vars=False`` this is the reference's ``subtract_mean`` hook
    (third_party/kaldi_signal.py:214-220).
    """
    if mode == "none":
        return feats
    out = torch.zeros_like(feats)
    if mode == "global":
        assert global_stats is not None
        cnt = global_stats[2, 0]
        g_mean = global_stats[0] / cnt
        g_var = (global_stats[1] / cnt - g_mean * g_mean).clamp_min(1e-20)
    for i, l in enumerate(lengths):
        x = feats[i, :l].double()
        if mode == "utterance":
            mean = x.sum(0) / l
            var = ((x * x).sum(0) / l - mean * mean).clamp_min(1e-20)
        elif mode == "global":
            mean, var = g_mean, g_var
        else:
            raise ValueError(mode)
        y = x - mean
        if norm_vars:
            y = y / var.sqrt()
        out[i, :l] = y.to(feats.dtype)
    return out


# --------------------------------------------------------------------------- SpecAug
def spec_aug(feats: torch.Tensor, lengths: torch.Tensor, conf: dict,
             uniforms: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """In-place SpecAugment.  Restates blocks/sp_layers.py:51-74.

    Means come from the un-masked input once; frequency masks are filled with the
    per-frame mean over D (all T rows, padding included), then time masks with the
    per-utterance mean over valid frames.  RNG: ``torch.rand(size=[B])`` on the
    feature device's generator in the order fs, f0s (per freq mask) then ts, t0s
    (per time mask).  ``uniforms`` (2*(F+T), B) replaces the draws in that order.
    """
    freq_means = torch.mean(feats, dim=-1)
    time_means = torch.sum(feats, dim=1) / lengths[:, None].float()
    B, T, V = feats.shape
    it = iter(uniforms) if uniforms is not None else None

    def draw():
        if it is not None:
            return next(it).to(feats.device)
        return torch.rand(size=[B], device=feats.device)

    for _ in range(conf["freq_mask_num"]):
        fs = (conf["freq_mask_width"] * draw()).long()
        f0s = ((V - fs).float() * draw()).long()
        for b in range(B):
            feats[b, :, f0s[b]:f0s[b] + fs[b]] = freq_means[b][:, None]
    for _ in range(conf["time_mask_num"]):
        ts = (conf["time_mask_width"] * draw()).long()
        t0s = ((lengths - ts).float() * draw()).long()
        for b in range(B):
            feats[b, t0s[b]:t0s[b] + ts[b], :] = time_means[b][None, :]
    return feats, lengths


# --------------------------------------------------------------------------- SPLayer
def splayer_forward(wav_batch: torch.Tensor, lengths: Sequence[int], config: dict,
                    training: bool = False, noises: Optional[List[torch.Tensor]] = None,
                    specaug_uniforms: Optional[torch.Tensor] = None,
                    global_stats: Optional[torch.Tensor] = None,
                    dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """Restates blocks/sp_layers.py:76-101 (fbank and offline branches).

    Two crashes of the mounted fork are repaired with their evident intent
    (SURVEY.md section 0.1): the Python list of frame counts becomes a LongTensor
    (:96) and a missing ``spec_aug`` key means "no SpecAug" (:98).  Extension
    keys (defaults = reference behaviour): ``dither`` (1.0), ``window_type``
    ('povey'), ``cmvn`` ('none'), ``cmvn_norm_vars`` (True).  Order: fbank ->
    zero-pad/stack -> CMVN -> SpecAug.
    """
    ftype = config["feature_type"]
    if ftype == "fbank":
        feats_list = []
        for i in range(wav_batch.shape[0]):  # :81-84
            n_i = int(lengths[i])
            f = fbank(wav_batch[i, :n_i], sample_rate=float(config["sample_rate"]),
                      num_mel_bins=int(config["num_mel_bins"]), use_energy=config["use_energy"],
                      dither=float(config.get("dither", 1.0)),
                      window_type=config.get("window_type", "povey"),
                      noise=None if noises is None else noises[i], dtype=dtype)
            feats_list.append(f)
        flens = [f.shape[0] for f in feats_list]
        padded = torch.zeros(len(flens), max(flens), feats_list[0].shape[-1], dtype=dtype)  # :87-91
        for i, f in enumerate(feats_list):
            padded[i, :flens[i], :] += f
        feat_lengths = torch.tensor(flens).long()
    elif ftype == "offline":  # :92-94
        padded = wav_batch
        feat_lengths = torch.as_tensor(lengths).long()
        flens = [int(v) for v in feat_lengths]
    else:
        raise ValueError("Unknown feature type.")  # :48

    mode = config.get("cmvn", "none")
    if mode != "none":
        padded = cmvn_apply(padded, flens, mode, bool(config.get("cmvn_norm_vars", True)), global_stats)

    if training and config.get("spec_aug") is not None:  # :98-99
        padded, feat_lengths = spec_aug(padded, feat_lengths, config["spec_aug"], specaug_uniforms)
    return padded, feat_lengths


# --------------------------------------------------------------------------- synthetic inputs
def synth_batch(B: int, n_lo: int, n_hi: int, sample_rate: int, seed: int = 1234):
    """SURVEY.md section 8d synthetic input (shared generator; lives in the product package so the
    benchmark's GPU arm does not import the oracle)."""
    from openasr_b200.synth import synth_batch as _sb
    return _sb(B, n_lo, n_hi, sample_rate, seed)
