"""ctypes binding of ``include/spl_capi.h`` (the C-ABI shared library ``lib/libspl_b200.so``).

This is the stub a maintainer of the reference would add next to
``src/blocks/sp_layers.py``; there is no CPU fallback: if the library is missing the
import of the product path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libspl_b200.so")

SPL_ABI_VERSION = 2
SPL_OK = 0
CMVN_MODES = {"none": 0, "utterance": 1, "global": 2}
SAMPLES_F32, SAMPLES_I16 = 0, 1


class SplConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("window_shift", C.c_int32),
        ("window_size", C.c_int32),
        ("padded_size", C.c_int32),
        ("num_mel_bins", C.c_int32),
        ("use_energy", C.c_int32),
        ("remove_dc", C.c_int32),
        ("preemph", C.c_float),
        ("dither", C.c_float),
    ]


class SplFbankArgs(C.Structure):
    _fields_ = [
        ("wav", C.c_void_p),
        ("wav_pitch", C.c_int64),
        ("wav_cols", C.c_int64),
        ("sample_format", C.c_int32),
        ("wav_len", C.c_void_p),
        ("B", C.c_int32),
        ("T", C.c_int32),
        ("feats", C.c_void_p),
        ("feat_len", C.c_void_p),
        ("noise", C.c_void_p),
        ("dither_seed", C.c_uint64),
        ("utt_stats", C.c_void_p),
        ("global_stats", C.c_void_p),
    ]


class SplPostArgs(C.Structure):
    _fields_ = [
        ("feats", C.c_void_p),
        ("feat_len", C.c_void_p),
        ("B", C.c_int32),
        ("T", C.c_int32),
        ("Dm", C.c_int32),
        ("cmvn_mode", C.c_int32),
        ("norm_vars", C.c_int32),
        ("utt_stats", C.c_void_p),
        ("global_mean", C.c_void_p),
        ("global_istd", C.c_void_p),
        ("n_freq_masks", C.c_int32),
        ("n_time_masks", C.c_int32),
        ("mask_params", C.c_void_p),
        ("mask_uniforms", C.c_void_p),
        ("freq_mask_width", C.c_float),
        ("time_mask_width", C.c_float),
    ]


EXPORTS = ("spl_create", "spl_destroy", "spl_fbank_forward", "spl_post_inplace", "spl_column_stats",
           "spl_feature_dim", "spl_abi_version", "spl_last_error", "spl_launch_count", "spl_tc_selftest", "spl_specaug_rects",
           "spl_conv0_relu", "spl_fbank_forward_multi", "spl_engine_name", "spl_debug_status", "spl_debug_umma_tables", "spl_debug_umma_acc", "spl_post_inplace_multi", "spl_forward_multi", "spl_debug_dither_noise")

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            "openasr_b200: CUDA extension %s is missing -- build it with "
            "`make -C openasr_b200/csrc` or `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback by design." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.spl_create.argtypes = [C.POINTER(SplConfig), C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.spl_create.restype = C.c_int
    lib.spl_destroy.argtypes = [C.c_void_p]
    lib.spl_destroy.restype = None
    lib.spl_fbank_forward.argtypes = [C.c_void_p, C.POINTER(SplFbankArgs), C.c_void_p]
    lib.spl_fbank_forward.restype = C.c_int
    lib.spl_fbank_forward_multi.argtypes = [C.c_void_p, C.POINTER(SplFbankArgs), C.c_int32, C.c_void_p]
    lib.spl_fbank_forward_multi.restype = C.c_int
    lib.spl_engine_name.argtypes = [C.c_void_p, C.c_int32]
    lib.spl_engine_name.restype = C.c_char_p
    lib.spl_debug_status.argtypes = [C.c_void_p]
    lib.spl_debug_status.restype = C.c_int
    lib.spl_debug_dither_noise.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, C.c_void_p]
    lib.spl_debug_dither_noise.restype = C.c_int
    lib.spl_debug_umma_acc.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.spl_debug_umma_acc.restype = C.c_int
    lib.spl_debug_umma_tables.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                          C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.spl_debug_umma_tables.restype = C.c_int
    lib.spl_post_inplace.argtypes = [C.c_void_p, C.POINTER(SplPostArgs), C.c_void_p]
    lib.spl_post_inplace.restype = C.c_int
    lib.spl_post_inplace_multi.argtypes = [C.c_void_p, C.POINTER(SplPostArgs), C.c_int32, C.c_void_p]
    lib.spl_post_inplace_multi.restype = C.c_int
    lib.spl_forward_multi.argtypes = [C.c_void_p, C.POINTER(SplFbankArgs), C.POINTER(SplPostArgs), C.c_int32, C.c_void_p]
    lib.spl_forward_multi.restype = C.c_int
    lib.spl_column_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_void_p]
    lib.spl_column_stats.restype = C.c_int
    lib.spl_tc_selftest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.spl_tc_selftest.restype = C.c_int
    lib.spl_specaug_rects.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                      C.c_int32, C.c_float, C.c_void_p]
    lib.spl_specaug_rects.restype = C.c_int
    lib.spl_conv0_relu.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                   C.c_int32, C.c_void_p, C.c_void_p]
    lib.spl_conv0_relu.restype = C.c_int
    lib.spl_feature_dim.argtypes = [C.c_void_p]
    lib.spl_feature_dim.restype = C.c_int
    lib.spl_abi_version.argtypes = []
    lib.spl_abi_version.restype = C.c_int
    lib.spl_last_error.argtypes = []
    lib.spl_last_error.restype = C.c_char_p
    lib.spl_launch_count.argtypes = []
    lib.spl_launch_count.restype = C.c_uint64
    if lib.spl_abi_version() != SPL_ABI_VERSION:
        raise RuntimeError("openasr_b200: libspl_b200.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != SPL_OK:
        msg = load().spl_last_error().decode("utf-8", "replace")
        raise RuntimeError("openasr_b200 %s failed (%d): %s" % (what, status, msg))


def launch_count() -> int:
    return int(load().spl_launch_count())
