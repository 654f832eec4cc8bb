"""openasr_b200 -- B200-native (sm_100a) implementation of OpenASR's online speech front-end.

Scope: the batched waveform-to-feature path of the reference's ``SPLayer``
(``src/blocks/sp_layers.py`` + ``src/third_party/kaldi_signal.py``): framing, dither, DC removal,
pre-emphasis, povey/hamming window, power spectrum, mel filterbank, log, CMVN, SpecAug.
The hot path is hand-written CUDA behind the C ABI of ``include/spl_capi.h``; see DESIGN.md.
"""
from .blocks.sp_layers import SPLayer, WavConv  # noqa: F401

__all__ = ["SPLayer", "WavConv"]
__version__ = "0.1.0"
