"""Drop-in replacement of the reference's ``blocks.sp_layers`` (``src/blocks/sp_layers.py``).

``SPLayer(config)`` keeps the reference's module API -- ``forward(wav_batch, lengths) ->
(padded_features, feature_lengths)``, ``spec_aug(padded_features, feature_lengths)``, the
attributes ``config`` / ``feature_type`` / ``spec_aug_conf`` and an empty ``state_dict()`` -- so
``src/frameworks/Speech_Models.py`` (constructor at :66-69, call at :123, package/restore at
:219-255) uses it unchanged.  The arithmetic runs in hand-written sm_100a CUDA behind
``include/spl_capi.h``; there is no CPU fallback.

Config keys (read exactly like ``sp_layers.py:27-46``):
    feature_type : "fbank" | "offline"                     (anything else: ValueError, :48)
    sample_rate, use_energy, num_mel_bins                  (fbank)
    spec_aug : {freq_mask_num, freq_mask_width, time_mask_num, time_mask_width}   (optional)
Optional extension keys; the defaults reproduce the reference's behaviour:
    dither         : 1.0      (the reference never overrides kaldi_signal.fbank's default)
    window_type    : "povey"  ("hamming" | "hanning" | "rectangular" | "blackman")
    cmvn           : "none" | "utterance" | "global"       (extension)
    cmvn_norm_vars : True
    dither_rng     : "device" (Philox stream on the GPU, same distribution) |
                     "host"   (the reference's exact CPU-generator stream, uploaded; parity mode)
    sync_free      : False    (True: CUDA ``lengths`` are never read back -- T comes from the padded width, which
                               the reference's collate makes equal to the longest utterance; the too-short
                               utterance assertion then only covers the padded width)
    specaug_rng    : "host"   (uniforms from the CPU default generator in the reference's draw
                               order: bit-exact against the reference run on CPU) |
                     "device" (torch.rand on the feature device, as the reference does on a GPU)
``self.config`` is kept un-mutated so ``Speech_Models.restore`` (:230-255) still matches old
checkpoints key by key.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from .. import frontend, tables


def _dither_transform(u: torch.Tensor) -> torch.Tensor:
    # kaldi_signal.py:176-177: x = max(eps, rand); sqrt(-2 ln x) * cos(2 pi x) with ONE uniform
    x = torch.max(torch.tensor(torch.finfo(torch.float32).eps), u)
    return torch.sqrt(-2 * x.log()) * torch.cos(2 * 3.141592653589793 * x)


class SPLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.feature_type = config["feature_type"]
        self.spec_aug_conf = None  # the reference leaves it unset without the key (crashes at :98)
        if "spec_aug" in config:
            sa = config["spec_aug"]
            self.spec_aug_conf = {k: sa[k] for k in
                                  ("freq_mask_num", "freq_mask_width", "time_mask_num", "time_mask_width")}
            # limits of kernel B, checked here rather than at the first call (spl_post_inplace: <= 32 masks)
            if int(sa["freq_mask_num"]) + int(sa["time_mask_num"]) > 32 or min(int(sa["freq_mask_num"]), int(sa["time_mask_num"])) < 0:
                raise ValueError("spec_aug: freq_mask_num + time_mask_num must be in [0, 32]")
        self.num_ceps = None
        if self.feature_type == "offline":
            self.func = None
        elif self.feature_type == "fbank":
            self._sample_rate = float(config["sample_rate"])
            self._use_energy = bool(config["use_energy"])
            self._num_mel_bins = int(config["num_mel_bins"])
            self._shift, self._win, self._padded = tables.frame_geometry(self._sample_rate)
            if self._padded not in (256, 512):
                raise ValueError("sample_rate %s gives a %d-point padded window; the B200 kernels support 256 and 512 "
                                 "(about 5.2 .. 20.4 kHz)" % (config["sample_rate"], self._padded))
            if not 4 <= self._num_mel_bins <= 128:
                raise ValueError("num_mel_bins must be in [4, 128]")
            self.func = self._fbank_single
        else:
            raise ValueError("Unknown feature type.")
        get = config.get if hasattr(config, "get") else (lambda k, d=None: config[k] if k in config else d)
        self._dither = float(get("dither", 1.0))
        self._window_type = str(get("window_type", "povey"))
        self._cmvn = str(get("cmvn", "none"))
        self._cmvn_norm_vars = bool(get("cmvn_norm_vars", True))
        self._dither_rng = str(get("dither_rng", "device"))
        self._specaug_rng = str(get("specaug_rng", "host"))
        self._sync_free = bool(get("sync_free", False))
        if self._cmvn not in ("none", "utterance", "global"):
            raise ValueError("cmvn must be one of none|utterance|global")
        if self._window_type not in tables.WINDOW_TYPES:
            raise ValueError("Invalid window type " + self._window_type)
        if self._dither_rng not in ("host", "device") or self._specaug_rng not in ("host", "device"):
            raise ValueError("dither_rng / specaug_rng must be 'host' or 'device'")
        # global CMVN (mean, 1/std): non-persistent so state_dict() stays empty (Speech_Models.py:219-228)
        self.register_buffer("_gmean", None, persistent=False)
        self.register_buffer("_gistd", None, persistent=False)
        self._stager = frontend.HostStager()

    # ------------------------------------------------------------------ helpers
    def _handle(self, device: torch.device) -> frontend.FbankHandle:
        return frontend.get_handle(device, self._sample_rate, self._num_mel_bins, self._use_energy,
                                   self._dither, self._window_type)

    def _fbank_single(self, waveform: torch.Tensor) -> torch.Tensor:
        """The reference's ``self.func`` closure (:40-46): (1, n) waveform -> (m, D) features."""
        n = waveform.shape[-1]
        f, _ = self._fbank_batch(waveform.reshape(1, -1), [n], need_stats=False)
        return f[0]

    def _host_lengths(self, lengths) -> List[int]:
        if isinstance(lengths, torch.Tensor):
            return [int(v) for v in lengths.detach().cpu().tolist()]  # one D2H sync if on the GPU
        return [int(v) for v in lengths]

    def _fbank_batch(self, wav_batch: torch.Tensor, lengths, need_stats: bool,
                     global_stats: Optional[torch.Tensor] = None):
        dev = wav_batch.device
        h = self._handle(dev)
        lens = self._host_lengths(lengths)
        B = wav_batch.shape[0]
        if len(lens) != B:
            raise ValueError("lengths must have one entry per utterance")
        for n in lens:
            # kaldi_signal.py:154 -- the reference asserts, never returns a silent empty row
            assert 2 <= h.win <= n, "choose a window size %d that is [2, %d]" % (h.win, n)
            if n > wav_batch.shape[1]:
                raise ValueError("length %d exceeds the padded batch width %d" % (n, wav_batch.shape[1]))
        frames = [tables.frame_count(n, h.win, h.shift) for n in lens]
        T = max(frames)
        if isinstance(lengths, torch.Tensor) and lengths.is_cuda and lengths.dtype == torch.int64:
            lens_dev = lengths.contiguous()
        else:
            lens_dev = torch.tensor(lens, dtype=torch.int64).to(dev, non_blocking=True)
        noise = None
        seed = 0
        if self._dither != 0.0:
            if self._dither_rng == "host":
                # the reference's consumption order: one (m_i, Nw) draw per utterance, batch order,
                # on the CPU default generator (kaldi_signal.py:176)
                noise_host = torch.zeros(B, T, h.win)
                for i, m in enumerate(frames):
                    noise_host[i, :m] = _dither_transform(torch.rand((m, h.win)))
                noise = noise_host.to(dev)
            else:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())  # CPU generator: torch.manual_seed applies
        utt_stats = torch.empty((B, 2, h.d_out), dtype=torch.float64, device=dev) if need_stats else None
        feats, feat_len = h.fbank(wav_batch, lens_dev, T, noise=noise, dither_seed=seed,
                                  utt_stats=utt_stats, global_stats=global_stats)
        return feats, (feat_len, frames, utt_stats)

    def _draw_rectangles(self, B: int, T: int, V: int, feat_len_dev: torch.Tensor,
                         frames: Optional[Sequence[int]]) -> Tuple[torch.Tensor, int, int]:
        conf = self.spec_aug_conf
        nf, nt = int(conf["freq_mask_num"]), int(conf["time_mask_num"])
        dev = feat_len_dev.device
        if self._specaug_rng == "host":
            u = frontend.specaug_uniforms(B, nf, nt, None)
            flen = torch.tensor(list(frames), dtype=torch.int64) if frames is not None else feat_len_dev.cpu()
            rect = frontend.specaug_rectangles(u, flen, T, V, conf).to(dev, non_blocking=True)
        else:
            u = frontend.specaug_uniforms(B, nf, nt, dev)
            rect = frontend.specaug_rectangles(u, feat_len_dev, T, V, conf)
        return rect, nf, nt

    # ------------------------------------------------------------------ reference API
    def spec_aug(self, padded_features, feature_lengths):
        """In-place SpecAugment of ``[B, T, V]`` features (sp_layers.py:51-74)."""
        if self.spec_aug_conf is None:
            raise AttributeError("SPLayer has no spec_aug configuration")
        frontend._require_cuda(padded_features, "padded_features")
        x = padded_features if (padded_features.is_contiguous() and padded_features.dtype == torch.float32) \
            else padded_features.float().contiguous()
        B, T, V = x.shape
        if V > 160:
            raise ValueError("spec_aug: feature rows wider than 160 are not supported by kernel B (got %d)" % V)
        flen_dev = torch.as_tensor(feature_lengths).long().to(x.device)
        frames = None if (isinstance(feature_lengths, torch.Tensor) and feature_lengths.is_cuda) \
            else [int(v) for v in torch.as_tensor(feature_lengths).tolist()]
        rect, nf, nt = self._draw_rectangles(B, T, V, flen_dev, frames)
        stats = frontend.column_stats(x, flen_dev) if nt > 0 else None
        frontend.post_inplace(x, flen_dev, cmvn_mode="none", utt_stats=stats, mask_params=rect, n_freq=nf, n_time=nt)
        if x is not padded_features:
            padded_features.copy_(x)
        return padded_features, feature_lengths

    def forward(self, wav_batch, lengths):
        aug = self.training and self.spec_aug_conf is not None
        if self.func is None:  # "offline": pre-computed features pass through (:92-94)
            padded_features = wav_batch
            feature_lengths = torch.as_tensor(lengths).long().to(padded_features.device)
            if self._cmvn != "none":
                raise ValueError("cmvn is only applied in the online fbank mode")
            if aug:
                padded_features, feature_lengths = self.spec_aug(padded_features, feature_lengths)
            return padded_features, feature_lengths

        frontend._require_cuda(wav_batch, "wav_batch")
        need_stats = self._cmvn == "utterance" or (aug and int(self.spec_aug_conf["time_mask_num"]) > 0)
        if self._dither_rng == "host" and self._dither != 0.0 or self._specaug_rng != "host":
            return self._forward_general(wav_batch, lengths, aug, need_stats)
        return self.forward_multi([(wav_batch, lengths)])[0]

    def forward_multi(self, batches):
        """Several ``(wav_batch, lengths)`` pairs in ONE call of the library (``spl_forward_multi``): one persistent
        launch of kernel A over all batches + one launch of kernel B, one pinned upload of the per-call integers.
        Returns ``[(padded_features, feature_lengths), ...]`` like consecutive ``forward`` calls (host RNG: ONE dither
        seed per call, then the SpecAug uniforms batch by batch in the reference's draw order).

        ``lengths`` on the host (list / CPU tensor): frame counts, ``T = max m_i`` and the short-utterance assertion
        (kaldi_signal.py:154) are evaluated here.  ``lengths`` as a CUDA int64 tensor with ``sync_free: True`` in the
        config: nothing is read back -- ``T`` comes from the padded width (the collate pads to the longest utterance,
        data_utils.py:126-138) and the SpecAug rectangles are resolved in kernel B from the uploaded uniforms."""
        if self.func is None:
            return [self.forward(w, l) for w, l in batches]
        aug = self.training and self.spec_aug_conf is not None
        need_stats = self._cmvn == "utterance" or (aug and int(self.spec_aug_conf["time_mask_num"]) > 0)
        if self._dither_rng == "host" and self._dither != 0.0 or self._specaug_rng != "host":
            return [self._forward_general(w, l, aug, need_stats) for w, l in batches]
        if self._cmvn == "global" and self._gmean is None:
            raise RuntimeError("cmvn='global' needs set_global_cmvn() (see openasr_b200.cmvn)")
        dev = batches[0][0].device
        h = self._handle(dev)
        n = len(batches)
        nf = nt = 0
        if aug:
            nf, nt = int(self.spec_aug_conf["freq_mask_num"]), int(self.spec_aug_conf["time_mask_num"])
        arrays, meta = [], []
        seed = 0
        for wav_batch, lengths in batches:
            frontend._require_cuda(wav_batch, "wav_batch")
            B = wav_batch.shape[0]
            lens_dev = None
            if isinstance(lengths, torch.Tensor) and lengths.is_cuda and self._sync_free:
                if lengths.dtype != torch.int64 or lengths.shape != (B,):
                    raise ValueError("device lengths must be an int64 tensor with one entry per utterance")
                lens_dev = lengths.contiguous()
                T = tables.frame_count(int(wav_batch.shape[1]), h.win, h.shift)
                assert wav_batch.shape[1] >= h.win, "choose a window size %d that is [2, %d]" % (h.win, wav_batch.shape[1])
            else:
                if isinstance(lengths, torch.Tensor):
                    lens_np = lengths.detach().cpu().numpy().astype(np.int64, copy=False)  # one D2H sync if on the GPU
                else:
                    lens_np = np.asarray(lengths, dtype=np.int64)
                if lens_np.shape != (B,):
                    raise ValueError("lengths must have one entry per utterance")
                n_min = int(lens_np.min())
                assert 2 <= h.win <= n_min, "choose a window size %d that is [2, %d]" % (h.win, n_min)  # kaldi_signal.py:154
                if int(lens_np.max()) > wav_batch.shape[1]:
                    raise ValueError("length %d exceeds the padded batch width %d" % (int(lens_np.max()), wav_batch.shape[1]))
                T = int((1 + (lens_np - h.win) // h.shift).max())
                arrays.append(lens_np)
            if self._dither != 0.0 and seed == 0:
                seed = int(torch.randint(1, 2 ** 62, (1,)).item())  # CPU generator: torch.manual_seed applies
            uni_idx = None
            if aug:
                uni_idx = len(arrays)
                arrays.append(frontend.specaug_uniforms(B, nf, nt, None).numpy())  # reference draw order, CPU generator
            meta.append((wav_batch, lens_dev, B, T, uni_idx))
        stream = torch.cuda.current_stream(dev)  # looked up once: every launch of this call goes there
        keep, ptrs = self._stager.upload(arrays, dev, stream) if arrays else (None, [])
        fa = (frontend._capi.SplFbankArgs * n)()
        pa = (frontend._capi.SplPostArgs * n)()
        outs, alive = [], [keep]
        ai = 0
        do_post = self._cmvn != "none" or aug
        # one contiguous fp64 buffer for the per-utterance sums of all batches: the library zeroes it with one memset
        stats_all = torch.empty((sum(m[2] for m in meta), 2, h.d_out), dtype=torch.float64, device=dev) if need_stats else None
        sb0 = 0
        for k, (wav_batch, lens_dev, B, T, uni_idx) in enumerate(meta):
            if lens_dev is None:
                lens_ptr = ptrs[ai]
                ai += 1
            else:
                lens_ptr = lens_dev.data_ptr()
            if uni_idx is not None:
                ai += 1
            feats = torch.empty((B, T, h.d_out), dtype=torch.float32, device=dev)
            feat_len = torch.empty((B,), dtype=torch.int64, device=dev)
            utt_stats = stats_all[sb0:sb0 + B] if need_stats else None
            sb0 += B
            alive.append(h._fill_args(fa[k], wav_batch, lens_ptr, T, None, seed, utt_stats, None, feats, feat_len))
            alive.append(stats_all)
            if do_post:
                a = pa[k]
                a.Dm = h.d_out
                a.cmvn_mode = frontend._capi.CMVN_MODES[self._cmvn]
                a.norm_vars = int(self._cmvn_norm_vars)
                a.global_mean = self._gmean.data_ptr() if self._gmean is not None else None
                a.global_istd = self._gistd.data_ptr() if self._gistd is not None else None
                if aug:
                    a.n_freq_masks, a.n_time_masks = nf, nt
                    a.mask_uniforms = ptrs[uni_idx]
                    a.freq_mask_width = float(self.spec_aug_conf["freq_mask_width"])
                    a.time_mask_width = float(self.spec_aug_conf["time_mask_width"])
            outs.append((feats, feat_len))
        frontend._capi.check(h._lib.spl_forward_multi(h._h, fa, pa if do_post else None, n,
                                                       frontend.C.c_void_p(stream.cuda_stream)), "spl_forward_multi")
        del alive  # stream-ordered allocator: the blocks are only reused behind the launches above
        return outs

    def _forward_general(self, wav_batch, lengths, aug, need_stats):
        """Parity-mode / device-RNG variants (host dither stream upload, torch.rand on the GPU)."""
        feats, (feat_len, frames, utt_stats) = self._fbank_batch(wav_batch, lengths, need_stats)
        if self._cmvn != "none" or aug:
            rect, nf, nt = (None, 0, 0)
            if aug:
                rect, nf, nt = self._draw_rectangles(feats.shape[0], feats.shape[1], feats.shape[2], feat_len, frames)
            if self._cmvn == "global" and self._gmean is None:
                raise RuntimeError("cmvn='global' needs set_global_cmvn() (see openasr_b200.cmvn)")
            frontend.post_inplace(feats, feat_len, cmvn_mode=self._cmvn, norm_vars=self._cmvn_norm_vars,
                                  utt_stats=utt_stats, global_mean=self._gmean, global_istd=self._gistd,
                                  mask_params=rect, n_freq=nf, n_time=nt)
        return feats, feat_len

    # ------------------------------------------------------------------ global CMVN (extension)
    def accumulate_cmvn_stats(self, wav_batch, lengths, stats: torch.Tensor) -> None:
        """Add this batch's (sum x, sum x^2, frame count) into ``stats`` (fp64 [2*D_out+1], on device).

        The local reduction is fused into kernel A's epilogue; see ``openasr_b200.cmvn`` for the
        cross-GPU all-reduce."""
        self._fbank_batch(wav_batch, lengths, need_stats=False, global_stats=stats)

    def set_global_cmvn(self, mean: torch.Tensor, istd: torch.Tensor) -> None:
        self._gmean = mean.detach().to(torch.float32).contiguous()
        self._gistd = istd.detach().to(torch.float32).contiguous()

    @property
    def feature_dim(self) -> Optional[int]:
        if self.func is None:
            return None
        return self._num_mel_bins + (1 if self._use_energy else 0)


class WavConv(nn.Module):
    """Raw-waveform conv encoder of the CPC / GRU-CTC recipes (``sp_layers.py:104-133``).

    Not on the accelerated path (learned convolutions; cuDNN) -- kept so that
    ``from blocks.sp_layers import WavConv`` (``CPC_Models.py:88``, ``Speech_Models.py:854``)
    keeps working when this module replaces the reference file, with identical parameter names
    (``encoder.{0,3,6,9,12}.weight`` + BatchNorm at ``{1,4,7,10,13}``) for checkpoint loading.
    Total stride 5*4*2*2*2 = 160.
    """

    _LAYERS = ((10, 5, 3), (8, 4, 2), (4, 2, 1), (4, 2, 1), (4, 2, 1))  # (kernel, stride, padding)

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.d_model = config["d_model"]
        mods: List[nn.Module] = []
        c_in = 1
        for k, s, pad in self._LAYERS:
            mods += [nn.Conv1d(c_in, self.d_model, kernel_size=k, stride=s, padding=pad, bias=False),
                     nn.BatchNorm1d(self.d_model), nn.ReLU(inplace=True)]
            c_in = self.d_model
        self.encoder = nn.Sequential(*mods)

    def forward(self, feats, feat_lengths):
        len_x = feat_lengths // 160
        x = self.encoder(feats.unsqueeze(1)).transpose(1, 2)
        return x[:, :len_x.max(), :], len_x
