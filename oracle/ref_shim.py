"""Import the UNMODIFIED reference front-end from /root/reference (build container only).

TEST INFRASTRUCTURE.  ``/root/reference`` does not exist on the GPU box, so this
module is used solely (a) by ``oracle/make_golden.py`` to generate the committed
vectors under ``tests/golden/`` and (b) by ``tests/test_oracle_vs_reference.py``
(skipped when the tree is absent) to pin ``oracle/frontend_oracle.py`` against
the real thing.  Nothing is copied: the reference files are imported where they
lie, with one external monkey-patch because ``torch.rfft`` (used at
``third_party/kaldi_signal.py:523``) was removed in torch 1.8.
"""
from __future__ import annotations

import os
import sys

import torch

REF_SRC = "/root/reference/src"
REF_TESTDATA = "/root/reference/test/testdata"


def available() -> bool:
    return os.path.isfile(os.path.join(REF_SRC, "third_party", "kaldi_signal.py"))


def _install_rfft_shim():
    if not hasattr(torch, "rfft"):
        def rfft(x, signal_ndim, normalized=False, onesided=True):
            assert signal_ndim == 1 and onesided and not normalized
            return torch.view_as_real(torch.fft.rfft(x, dim=-1))
        torch.rfft = rfft  # external shim; the reference file itself stays untouched


def load():
    """Returns (kaldi_signal module, SPLayer class) of the reference."""
    if not available():
        raise RuntimeError("reference tree not present")
    _install_rfft_shim()
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    from third_party import kaldi_signal as ksp  # noqa: E402
    from blocks.sp_layers import SPLayer  # noqa: E402
    return ksp, SPLayer


def read_wav(name: str):
    """int16 PCM fixture -> (sample_rate, float32 numpy, int16-scaled) like utils.load_wave
    (src/utils.py:77-104: ``data.astype(np.float32)``, no scaling)."""
    import numpy as np
    import scipy.io.wavfile as wavfile
    sr, data = wavfile.read(os.path.join(REF_TESTDATA, name))
    return sr, data.astype(np.float32)
