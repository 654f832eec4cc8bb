"""Kaldi ark / scp matrix reader for the offline-feature path.

Replaces, for this path only, ``third_party/kaldi_io.py:362-448`` (``read_mat`` and its binary / ascii /
compressed decoders) and ``dataload/data_utils.py:141-154`` (``load_feat_batch``): every shipped recipe feeds
``SPLayer(feature_type="offline")`` with pre-computed Kaldi features read this way (``sp_layers.py:92-99``).

    rxfilename  :=  "path"  |  "path:byte_offset"          (the second column of a feats.scp)
    read_mat(rx)            -> float32 / float64 ndarray [rows, cols]
    read_scp(path)          -> [(key, rxfilename), ...]
    load_feat_batch(paths)  -> (padded [B, T, D] float32 tensor (pinned when CUDA is present), lengths int64 [B])

Formats: binary ``FM`` / ``DM`` (float / double), binary ``CM`` / ``CM2`` / ``CM3`` (Kaldi compressed matrices;
the reference decodes ``CM`` only), and text matrices.  Pipes and gzip are data-loader plumbing and out of scope.
"""
from __future__ import annotations

import io
import os
import struct
from typing import BinaryIO, List, Sequence, Tuple, Union

import numpy as np
import torch


class ArkFormatError(ValueError):
    pass


def _open(rx: Union[str, BinaryIO]) -> Tuple[BinaryIO, bool]:
    if not isinstance(rx, str):
        return rx, False
    path, offset = rx, None
    if ":" in rx:
        head, tail = rx.rsplit(":", 1)
        if tail.isdigit():
            path, offset = head, int(tail)
    fd = open(path, "rb")
    if offset is not None:
        fd.seek(offset)
    return fd, True


def _need(fd: BinaryIO, n: int) -> bytes:
    buf = fd.read(n)
    if len(buf) != n:
        raise ArkFormatError("unexpected end of file (wanted %d bytes, got %d)" % (n, len(buf)))
    return buf


def _dims(fd: BinaryIO) -> Tuple[int, int]:
    # '\4' int32 rows '\4' int32 cols
    s1, rows, s2, cols = struct.unpack("<bibi", _need(fd, 10))
    if s1 != 4 or s2 != 4 or rows < 0 or cols < 0:
        raise ArkFormatError("bad matrix dimension header")
    return rows, cols


def _read_compressed(fd: BinaryIO, token: str) -> np.ndarray:
    """Kaldi CompressedMatrix (compressed-matrix.h): global header (min, range, rows, cols), then per format
    CM : per-column uint16 percentiles (0, 25, 75, 100) + uint8 data, column-major, piecewise-linear decode
    CM2: uint16 data, row-major;  CM3: uint8 data, row-major (both linear in [min, min + range])."""
    gmin, grange, rows, cols = struct.unpack("<ffii", _need(fd, 16))
    if token == "CM ":
        hdr = np.frombuffer(_need(fd, cols * 8), dtype="<u2").reshape(cols, 4).astype(np.float32)
        # same float32 operation order as the reference decoder (kaldi_io.py:418): (u16 * range) * 2^-16-ish + min
        pct = (hdr * np.float32(grange) * np.float32(1.52590218966964e-05) + np.float32(gmin)).astype(np.float32)
        data = np.frombuffer(_need(fd, rows * cols), dtype=np.uint8).reshape(cols, rows).astype(np.float32)
        p0, p25, p75, p100 = (pct[:, i:i + 1] for i in range(4))
        lo = p0 + (p25 - p0) / np.float32(64.0) * data
        mid = p25 + (p75 - p25) / np.float32(128.0) * (data - np.float32(64.0))
        hi = p75 + (p100 - p75) / np.float32(63.0) * (data - np.float32(192.0))
        out = np.where(data <= 64, lo, np.where(data <= 192, mid, hi)).astype(np.float32)
        return np.ascontiguousarray(out.T)
    if token == "CM2":
        data = np.frombuffer(_need(fd, rows * cols * 2), dtype="<u2").reshape(rows, cols).astype(np.float32)
        return (np.float32(gmin) + np.float32(grange) * np.float32(1.0 / 65535.0) * data).astype(np.float32)
    if token == "CM3":
        data = np.frombuffer(_need(fd, rows * cols), dtype=np.uint8).reshape(rows, cols).astype(np.float32)
        return (np.float32(gmin) + np.float32(grange) * np.float32(1.0 / 255.0) * data).astype(np.float32)
    raise ArkFormatError("unknown compressed-matrix token %r" % token)


def _read_text(fd: BinaryIO) -> np.ndarray:
    rows: List[np.ndarray] = []
    while True:
        line = fd.readline()
        if not line:
            raise ArkFormatError("text matrix without closing bracket")
        tok = line.decode().split()
        if not tok:
            continue
        last = tok[-1] == "]"
        if last:
            tok = tok[:-1]
        if tok:
            rows.append(np.asarray(tok, dtype=np.float32))
        if last:
            return np.vstack(rows) if rows else np.zeros((0, 0), np.float32)


def read_mat(rx: Union[str, BinaryIO]) -> np.ndarray:
    """One Kaldi matrix from an rxfilename (``path`` or ``path:offset``) or an open binary file positioned at it."""
    fd, own = _open(rx)
    try:
        flag = _need(fd, 2)
        if flag == b"\0B":
            token = _need(fd, 3).decode("ascii", "replace")
            if token.startswith("CM"):
                return _read_compressed(fd, token)
            if token == "FM ":
                dtype, size = "<f4", 4
            elif token == "DM ":
                dtype, size = "<f8", 8
            else:
                raise ArkFormatError("unknown matrix header %r" % token)
            rows, cols = _dims(fd)
            return np.frombuffer(_need(fd, rows * cols * size), dtype=dtype).reshape(rows, cols)
        if flag == b" [":
            return _read_text(fd)
        raise ArkFormatError("not a Kaldi matrix (starts with %r)" % flag)
    finally:
        if own:
            fd.close()


def read_scp(path: str) -> List[Tuple[str, str]]:
    """``key rxfilename`` lines of a Kaldi script file."""
    out = []
    with open(path, "r") as f:
        for line in f:
            line = line.strip()
            if line:
                key, rx = line.split(None, 1)
                out.append((key, rx))
    return out


def read_ark(path: str):
    """Generator over ``(key, matrix)`` of a binary / text ark file (sequential read)."""
    with open(path, "rb") as fd:
        while True:
            key = bytearray()
            while True:
                ch = fd.read(1)
                if not ch:
                    if key:
                        raise ArkFormatError("truncated key")
                    return
                if ch == b" ":
                    break
                key += ch
            yield key.decode(), read_mat(fd)


def write_mat(fd: BinaryIO, key: str, mat: np.ndarray) -> int:
    """Append ``key`` + a binary float matrix to an open ark; returns the byte offset for the scp line."""
    mat = np.ascontiguousarray(mat)
    token = {np.dtype(np.float32): b"FM ", np.dtype(np.float64): b"DM "}[mat.dtype]
    fd.write(key.encode() + b" ")
    off = fd.tell()
    fd.write(b"\0B" + token + struct.pack("<bibi", 4, mat.shape[0], 4, mat.shape[1]) + mat.tobytes())
    return off


def load_feat_batch(paths: Sequence[str], pin_memory: bool = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """``load_feat_batch`` of ``data_utils.py:141-154``: read every matrix, zero-pad to ``[B, T_max, D]`` float32 and
    return the int64 lengths.  The batch is assembled directly in a pinned buffer (when CUDA is available) so that
    ``padded.to(device, non_blocking=True)`` -> ``SPLayer(feature_type="offline")`` is one asynchronous copy."""
    mats = [read_mat(p) for p in paths]
    lengths = [int(m.shape[0]) for m in mats]
    dim = int(mats[-1].shape[1])
    for m in mats:
        if m.shape[1] != dim:
            raise ArkFormatError("feature matrices of one batch must have the same width")
    if pin_memory is None:
        pin_memory = torch.cuda.is_available()
    padded = torch.zeros((len(mats), max(lengths), dim), dtype=torch.float32)
    if pin_memory:
        padded = padded.pin_memory()
    dst = padded.numpy()
    for i, m in enumerate(mats):
        dst[i, :lengths[i]] = m
    return padded, torch.tensor(lengths, dtype=torch.int64)
