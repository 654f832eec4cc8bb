"""Python host of the B200 front-end: handle cache, buffer ownership, launches.

PyTorch is used for device memory and streams only; all arithmetic of the hot path runs in
``lib/libspl_b200.so`` (hand-written sm_100a CUDA) behind the C ABI of ``include/spl_capi.h``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _capi, tables

_handles: Dict[tuple, "FbankHandle"] = {}
_handles_lock = threading.Lock()


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError("openasr_b200: %s must be a CUDA tensor -- this front-end has no CPU path "
                           "(the CPU restatement lives in oracle/ and is test infrastructure only)" % what)


def _stream_ptr(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class FbankHandle:
    """One ``spl_handle`` (device-resident window / twiddle / sparse mel tables)."""

    def __init__(self, device_index: int, sample_rate: float, num_mel_bins: int, use_energy: bool,
                 dither: float, window_type: str):
        lib = _capi.load()
        self.shift, self.win, self.padded = tables.frame_geometry(sample_rate)
        self.num_mel_bins = int(num_mel_bins)
        self.use_energy = bool(use_energy)
        self.dither = float(dither)
        self.d_out = self.num_mel_bins + (1 if self.use_energy else 0)
        self.device_index = device_index
        if self.padded not in (256, 512):
            raise ValueError("openasr_b200: sample_rate %s gives a %d-point padded window; the B200 kernels "
                             "support 256 and 512 (8 kHz .. 20 kHz)" % (sample_rate, self.padded))
        window = tables.window_table(window_type, self.win).to(torch.float32).contiguous()
        mel = tables.mel_table(self.num_mel_bins, self.padded, sample_rate).to(torch.float32).contiguous()
        cfg = _capi.SplConfig(_capi.SPL_ABI_VERSION, self.shift, self.win, self.padded, self.num_mel_bins,
                              int(self.use_energy), 1, 0.97, self.dither)
        out = C.c_void_p()
        _capi.check(lib.spl_create(C.byref(cfg), C.c_void_p(window.data_ptr()), C.c_void_p(mel.data_ptr()),
                                   device_index, C.byref(out)), "spl_create")
        self._h = out
        self._lib = lib

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.spl_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def engine_name(self, dtype=torch.float32) -> str:
        """Kernel-A engine calls with this sample dtype run on: 'umma' (tcgen05), 'fft' or 'simple'."""
        fmt = _capi.SAMPLES_F32 if dtype == torch.float32 else _capi.SAMPLES_I16
        return self._lib.spl_engine_name(self._h, fmt).decode()

    def debug_status(self) -> int:
        """SYNCHRONOUS: non-zero if a kernel of this handle gave up on an internal barrier (tests only)."""
        return int(self._lib.spl_debug_status(self._h))

    # ------------------------------------------------------------------ kernel A
    def _fill_args(self, a, wav, wav_len_dev, T, noise, dither_seed, utt_stats, global_stats, out, feat_len):
        _require_cuda(wav, "wav_batch")
        if wav.dim() != 2:
            raise ValueError("wav_batch must be [B, L]")
        if wav.dtype == torch.float32:
            fmt = _capi.SAMPLES_F32
        elif wav.dtype == torch.int16:
            fmt = _capi.SAMPLES_I16
        else:
            raise TypeError("wav_batch must be float32 or int16, got %s" % wav.dtype)
        if wav.stride(1) != 1:
            wav = wav.contiguous()
        B = wav.shape[0]
        a.wav = wav.data_ptr()
        a.wav_pitch = wav.stride(0) if B > 1 else wav.shape[1]
        a.wav_cols = wav.shape[1]
        a.sample_format = fmt
        a.wav_len = wav_len_dev if isinstance(wav_len_dev, int) else wav_len_dev.data_ptr()
        a.B = B
        a.T = T
        a.feats = out.data_ptr()
        a.feat_len = feat_len.data_ptr()
        a.noise = noise.data_ptr() if noise is not None else None
        a.dither_seed = dither_seed & 0xFFFFFFFFFFFFFFFF
        a.utt_stats = utt_stats.data_ptr() if utt_stats is not None else None
        a.global_stats = global_stats.data_ptr() if global_stats is not None else None
        return wav  # keeps a contiguous copy alive

    def fbank_multi(self, items, *, dither_seed: int = 0, global_stats: Optional[torch.Tensor] = None,
                    stream_ptr: Optional[C.c_void_p] = None) -> None:
        """Several batches in ONE launch of kernel A (spl_fbank_forward_multi).  ``items``: dicts with
        wav [B, L], lens (device int64), T, feats [B, T, D_out], flen [B] and optionally stats [B, 2, D_out],
        noise [B, T, Nw]; all on this handle's device, every output pre-allocated by the caller."""
        n = len(items)
        arr = (_capi.SplFbankArgs * n)()
        keep = []
        for a, it in zip(arr, items):
            keep.append(self._fill_args(a, it["wav"], it["lens"], it["T"], it.get("noise"), dither_seed,
                                        it.get("stats"), global_stats, it["feats"], it["flen"]))
        dev = items[0]["wav"].device
        _capi.check(self._lib.spl_fbank_forward_multi(self._h, arr, n, stream_ptr or _stream_ptr(dev)),
                    "spl_fbank_forward_multi")

    def fbank(self, wav: torch.Tensor, wav_len_dev: torch.Tensor, T: int, *,
              noise: Optional[torch.Tensor] = None, dither_seed: int = 0,
              utt_stats: Optional[torch.Tensor] = None, global_stats: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None, feat_len: Optional[torch.Tensor] = None,
              stream_ptr: Optional[C.c_void_p] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """wav [B, L] (fp32 or int16, int16-scaled) + device int64 lengths -> feats [B, T, D_out], feat_len [B]."""
        _require_cuda(wav, "wav_batch")
        if wav.dim() != 2:
            raise ValueError("wav_batch must be [B, L]")
        if wav.dtype == torch.float32:
            fmt = _capi.SAMPLES_F32
        elif wav.dtype == torch.int16:
            fmt = _capi.SAMPLES_I16
        else:
            raise TypeError("wav_batch must be float32 or int16, got %s" % wav.dtype)
        if wav.stride(1) != 1:
            wav = wav.contiguous()
        B = wav.shape[0]
        dev = wav.device
        if out is None:
            out = torch.empty((B, T, self.d_out), dtype=torch.float32, device=dev)
        if feat_len is None:
            feat_len = torch.empty((B,), dtype=torch.int64, device=dev)
        a = _capi.SplFbankArgs()
        a.wav = wav.data_ptr()
        a.wav_pitch = wav.stride(0) if B > 1 else wav.shape[1]
        a.wav_cols = wav.shape[1]
        a.sample_format = fmt
        a.wav_len = wav_len_dev if isinstance(wav_len_dev, int) else wav_len_dev.data_ptr()
        a.B = B
        a.T = T
        a.feats = out.data_ptr()
        a.feat_len = feat_len.data_ptr()
        a.noise = noise.data_ptr() if noise is not None else None
        a.dither_seed = dither_seed & 0xFFFFFFFFFFFFFFFF
        a.utt_stats = utt_stats.data_ptr() if utt_stats is not None else None
        a.global_stats = global_stats.data_ptr() if global_stats is not None else None
        _capi.check(self._lib.spl_fbank_forward(self._h, C.byref(a), stream_ptr or _stream_ptr(dev)),
                    "spl_fbank_forward")
        return out, feat_len


def get_handle(device: torch.device, sample_rate: float, num_mel_bins: int, use_energy: bool,
               dither: float, window_type: str) -> FbankHandle:
    """Handle cache keyed by (device, config); safe under DataParallel's per-GPU threads."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, float(sample_rate), int(num_mel_bins), bool(use_energy), float(dither), str(window_type),
           os.environ.get("SPL_ENGINE", ""), os.environ.get("SPL_CTAS_PER_SM", ""),
           os.environ.get("SPL_UMMA_DEBUG", ""), os.environ.get("SPL_UMMA_CTAS", ""))  # the library reads the switches at spl_create
    with _handles_lock:
        h = _handles.get(key)
        if h is None:
            h = FbankHandle(idx, sample_rate, num_mel_bins, use_energy, dither, window_type)
            _handles[key] = h
        return h


# ---------------------------------------------------------------------- kernel B
def post_inplace(feats: torch.Tensor, feat_len: torch.Tensor, *, cmvn_mode: str = "none",
                 norm_vars: bool = True, utt_stats: Optional[torch.Tensor] = None,
                 global_mean: Optional[torch.Tensor] = None, global_istd: Optional[torch.Tensor] = None,
                 mask_params: Optional[torch.Tensor] = None, n_freq: int = 0, n_time: int = 0,
                 mask_uniforms: Optional[torch.Tensor] = None, freq_width: float = 0.0, time_width: float = 0.0,
                 handle: Optional["FbankHandle"] = None, stream_ptr: Optional[C.c_void_p] = None) -> None:
    """CMVN + SpecAug in place on a contiguous CUDA [B, T, D] fp32 tensor.  With ``handle`` the library
    selects the handle's device itself (no Python device context needed).  Masks: either resolved rectangles
    (``mask_params`` int32 [B, F+T, 2]) or the ``2 (F+T) x B`` uniforms in the reference's draw order
    (``mask_uniforms`` + widths), which kernel B resolves against the device ``feat_len`` itself."""
    _require_cuda(feats, "features")
    if feats.dtype != torch.float32 or not feats.is_contiguous():
        raise ValueError("features must be contiguous float32")
    lib = _capi.load()
    B, T, Dm = feats.shape
    a = _capi.SplPostArgs()
    a.feats = feats.data_ptr()
    a.feat_len = feat_len.data_ptr()
    a.B, a.T, a.Dm = B, T, Dm
    a.cmvn_mode = _capi.CMVN_MODES[cmvn_mode]
    a.norm_vars = int(bool(norm_vars))
    a.utt_stats = utt_stats.data_ptr() if utt_stats is not None else None
    a.global_mean = global_mean.data_ptr() if global_mean is not None else None
    a.global_istd = global_istd.data_ptr() if global_istd is not None else None
    a.n_freq_masks, a.n_time_masks = int(n_freq), int(n_time)
    a.mask_params = (mask_params if isinstance(mask_params, int) else mask_params.data_ptr()) \
        if mask_params is not None else None
    if mask_uniforms is not None:
        a.mask_uniforms = mask_uniforms if isinstance(mask_uniforms, int) else mask_uniforms.data_ptr()
        a.freq_mask_width, a.time_mask_width = float(freq_width), float(time_width)
    if handle is not None:
        _capi.check(lib.spl_post_inplace(handle._h, C.byref(a), stream_ptr or _stream_ptr(feats.device)),
                    "spl_post_inplace")
        return
    with torch.cuda.device(feats.device):
        _capi.check(lib.spl_post_inplace(None, C.byref(a), stream_ptr or _stream_ptr(feats.device)),
                    "spl_post_inplace")


def column_stats(feats: torch.Tensor, feat_len: torch.Tensor) -> torch.Tensor:
    """fp64 [B, 2, D] per-utterance sum x / sum x^2 over valid frames (offline-feature path)."""
    _require_cuda(feats, "features")
    lib = _capi.load()
    B, T, Dm = feats.shape
    st = torch.empty((B, 2, Dm), dtype=torch.float64, device=feats.device)
    with torch.cuda.device(feats.device):
        _capi.check(lib.spl_column_stats(None, C.c_void_p(feats.data_ptr()), C.c_void_p(feat_len.data_ptr()),
                                         B, T, Dm, C.c_void_p(st.data_ptr()), _stream_ptr(feats.device)),
                    "spl_column_stats")
    return st


# ---------------------------------------------------------------------- host-side SpecAug draws
def specaug_uniforms(B: int, n_freq: int, n_time: int, device=None) -> torch.Tensor:
    """2*(F+T) x B uniforms in the reference's draw order (sp_layers.py:58-71).

    On the CPU generator one (2(F+T), B) draw equals the reference's sequence of
    ``torch.rand(size=[B])`` calls bit for bit (mt19937 stream; checked in tests).  On a CUDA
    generator the calls are issued one by one, as the reference does, because Philox offsets
    advance per call.
    """
    n = 2 * (n_freq + n_time)
    if device is None or torch.device(device).type == "cpu":
        return torch.rand(n, B)
    return torch.stack([torch.rand(size=[B], device=device) for _ in range(n)])


def specaug_rectangles(uniforms: torch.Tensor, feat_len: torch.Tensor, T: int, V: int, conf: dict) -> torch.Tensor:
    """Turn the uniforms into int32 [B, F+T, 2] half-open (start, end) ranges.

    Same float32 arithmetic as sp_layers.py:59-62 / :68-71 (``(W * u).long()``,
    ``((V - fs).float() * u).long()``), then Python slice semantics of ``x[b, s:s+w]``
    (:64, :73) resolved explicitly so negative starts / spills behave like the reference.
    Works on CPU or CUDA tensors (no host sync when everything is on the device).
    """
    F_, T_ = conf["freq_mask_num"], conf["time_mask_num"]
    flen = feat_len.to(uniforms.device).long()
    rects = []
    r = 0

    def resolve(start, width, size):
        end = start + width
        s = torch.where(start < 0, start + size, start).clamp(0, size)
        e = torch.where(end < 0, end + size, end).clamp(0, size)
        return torch.stack([s, torch.max(s, e)], dim=-1)

    for _ in range(F_):
        fs = (conf["freq_mask_width"] * uniforms[r]).long()
        f0s = ((V - fs).float() * uniforms[r + 1]).long()
        r += 2
        rects.append(resolve(f0s, fs, V))
    for _ in range(T_):
        ts = (conf["time_mask_width"] * uniforms[r]).long()
        t0s = ((flen - ts).float() * uniforms[r + 1]).long()
        r += 2
        rects.append(resolve(t0s, ts, T))
    return torch.stack(rects, dim=1).to(torch.int32).contiguous()


# ---------------------------------------------------------------------- fast host path
def specaug_rectangles_np(uniforms: np.ndarray, frames: np.ndarray, T: int, V: int, conf: dict) -> np.ndarray:
    """numpy twin of :func:`specaug_rectangles` for the per-call host path (same float32 arithmetic,
    same truncation, same slice resolution), vectorised over the masks: ~15 numpy calls per forward."""
    F_, T_ = int(conf["freq_mask_num"]), int(conf["time_mask_num"])
    B = uniforms.shape[1]
    M = F_ + T_
    u_w = uniforms[0::2]                      # [M, B] width draws  (fs, ts)
    u_s = uniforms[1::2]                      # [M, B] start draws  (f0s, t0s)
    wmax = np.empty((M, 1), dtype=np.float32)
    wmax[:F_] = np.float32(conf["freq_mask_width"])
    wmax[F_:] = np.float32(conf["time_mask_width"])
    width = (wmax * u_w).astype(np.int64)                        # (W * rand).long()
    limit = np.empty((M, B), dtype=np.int64)
    limit[:F_] = V
    limit[F_:] = frames.astype(np.int64)[None, :]
    start = ((limit - width).astype(np.float32) * u_s).astype(np.int64)   # ((V - fs).float() * rand).long()
    size = np.empty((M, 1), dtype=np.int64)
    size[:F_] = V
    size[F_:] = T
    end = start + width
    s_ = np.clip(np.where(start < 0, start + size, start), 0, size)      # Python slice semantics
    e_ = np.clip(np.where(end < 0, end + size, end), 0, size)
    out = np.empty((B, M, 2), dtype=np.int32)
    out[:, :, 0] = s_.T
    out[:, :, 1] = np.maximum(s_, e_).T
    return out


def specaug_rectangles_c(uniforms: np.ndarray, frames: np.ndarray, T: int, V: int, conf: dict) -> np.ndarray:
    """Same as :func:`specaug_rectangles_np`, computed by the library's host helper (one ctypes call)."""
    lib = _capi.load()
    F_, T_ = int(conf["freq_mask_num"]), int(conf["time_mask_num"])
    B = uniforms.shape[1]
    u = np.ascontiguousarray(uniforms, dtype=np.float32)
    fr = np.ascontiguousarray(frames, dtype=np.int64)
    out = np.empty((B, F_ + T_, 2), dtype=np.int32)
    _capi.check(lib.spl_specaug_rects(u.ctypes.data, fr.ctypes.data, B, int(T), int(V), F_,
                                      float(conf["freq_mask_width"]), T_, float(conf["time_mask_width"]),
                                      out.ctypes.data), "spl_specaug_rects")
    return out


class HostStager:
    """Rings of pinned host slots for the few hundred bytes a call uploads (lengths + mask rectangles / uniforms):
    one asynchronous H2D per forward instead of several pageable copies.  One ring PER DEVICE (DataParallel
    replicas share the module's attributes, and a CUDA event belongs to the device it was first recorded on);
    slot selection, growth and reuse all happen under the lock.  A slot is reused only after the event recorded
    behind its copy has completed."""

    def __init__(self, slots: int = 16, nbytes: int = 1 << 14):
        self._nslots, self._nbytes = slots, nbytes
        self._rings = {}  # device index -> {"slots": [...], "np": [...], "events": [...], "next": int}
        self._lock = threading.Lock()

    def _ring(self, index: int) -> dict:
        ring = self._rings.get(index)
        if ring is None:  # pinned memory needs a driver: allocate on first use
            slots = [torch.empty(self._nbytes, dtype=torch.uint8).pin_memory() for _ in range(self._nslots)]
            ring = {"slots": slots, "np": [t.numpy() for t in slots], "events": [None] * self._nslots, "next": 0}
            self._rings[index] = ring
        return ring

    def upload(self, arrays, device: torch.device, stream=None) -> Tuple[torch.Tensor, list]:
        """arrays: list of contiguous numpy arrays (8-byte aligned sizes handled here).
        Returns (device uint8 tensor keeping the memory alive, list of device pointers)."""
        sizes = [(a.nbytes + 7) & ~7 for a in arrays]
        total = sum(sizes)
        index = device.index if device.index is not None else torch.cuda.current_device()
        with self._lock:
            ring = self._ring(index)
            i = ring["next"]
            ring["next"] = (i + 1) % self._nslots
            if total > ring["slots"][i].numel():
                ring["slots"][i] = torch.empty(total * 2, dtype=torch.uint8).pin_memory()
                ring["np"][i] = ring["slots"][i].numpy()
            ev = ring["events"][i]
            if ev is None:
                with torch.cuda.device(index):
                    ev = ring["events"][i] = torch.cuda.Event()
                first = True
            else:
                first = False
            slot, host = ring["slots"][i], ring["np"][i]
        if not first:
            ev.synchronize()  # the copy that last used this slot has left the host buffer
        offs, o = [], 0
        for a, sz in zip(arrays, sizes):
            host[o:o + a.nbytes] = a.reshape(-1).view(np.uint8)
            offs.append(o)
            o += sz
        dev = torch.empty(total, dtype=torch.uint8, device=device)
        dev.copy_(slot[:total], non_blocking=True)
        ev.record(stream if stream is not None else torch.cuda.current_stream(device))
        base = dev.data_ptr()
        return dev, [base + x for x in offs]
