"""CPU: the Kaldi ark/scp reader of the offline-feature path (row f3) against what the UNMODIFIED reference reader
(third_party/kaldi_io.py:362-448) decoded from the same archive (oracle/make_golden_ark.py wrote both)."""
import io
import os

import numpy as np
import pytest
import torch

from openasr_b200 import ark_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _scp():
    return [(k, os.path.join(ROOT, rx)) for k, rx in ark_io.read_scp(os.path.join(ROOT, "tests", "golden", "feats.scp"))]


def test_read_mat_matches_reference_reader(golden_dir):
    ref = np.load(os.path.join(golden_dir, "ark_ref.npz"))
    seen = set()
    for key, rx in _scp():
        got = ark_io.read_mat(rx)
        assert got.dtype == ref[key].dtype and got.shape == ref[key].shape, key
        assert np.array_equal(got, ref[key]), key  # bit-identical, the compressed format included
        seen.add(key)
    assert seen == {"utt_f32", "utt_f64", "utt_cm", "utt_txt"}
    # sequential ark read yields the same matrices in file order
    keys = [k for k, _ in ark_io.read_ark(os.path.join(golden_dir, "feats.ark"))]
    assert keys == ["utt_f32", "utt_f64", "utt_cm", "utt_txt"]


def test_load_feat_batch_matches_reference_collate(golden_dir):
    """data_utils.py:141-154: zero-padded [B, T, D] float32 + int64 lengths."""
    ref = np.load(os.path.join(golden_dir, "ark_ref.npz"))
    paths = [rx for _, rx in _scp()]
    padded, lengths = ark_io.load_feat_batch(paths, pin_memory=False)
    assert padded.dtype == torch.float32 and lengths.dtype == torch.int64
    assert lengths.tolist() == [37, 11, 53, 5] and tuple(padded.shape) == (4, 53, 40)
    for i, (key, _) in enumerate(_scp()):
        want = torch.from_numpy(np.asarray(ref[key], dtype=np.float32))
        assert torch.allclose(padded[i, :lengths[i]], want, atol=2e-6, rtol=0)
        assert (padded[i, lengths[i]:] == 0).all()


def test_roundtrip_and_errors(tmp_path):
    m = np.random.RandomState(0).randn(7, 5).astype(np.float32)
    p = tmp_path / "a.ark"
    with open(p, "wb") as fd:
        off = ark_io.write_mat(fd, "k", m)
    assert np.array_equal(ark_io.read_mat("%s:%d" % (p, off)), m)
    with pytest.raises(ark_io.ArkFormatError):
        ark_io.read_mat(io.BytesIO(b"\0BXX \4\0\0\0\0\4\0\0\0\0"))
    with pytest.raises(ark_io.ArkFormatError):
        ark_io.read_mat(io.BytesIO(b"\0BFM \4\7\0\0\0\4\5\0\0\0abc"))  # truncated payload


def test_cm2_cm3_decode():
    import struct
    rows, cols = 3, 4
    raw16 = np.arange(rows * cols, dtype="<u2").reshape(rows, cols) * 5000
    buf = b"\0BCM2" + struct.pack("<ffii", -1.0, 2.0, rows, cols) + raw16.tobytes()
    got = ark_io.read_mat(io.BytesIO(buf))
    np.testing.assert_allclose(got, -1.0 + 2.0 * raw16.astype(np.float64) / 65535.0, atol=1e-6)
    raw8 = (np.arange(rows * cols, dtype=np.uint8).reshape(rows, cols) * 20)
    buf = b"\0BCM3" + struct.pack("<ffii", 0.5, 4.0, rows, cols) + raw8.tobytes()
    got = ark_io.read_mat(io.BytesIO(buf))
    np.testing.assert_allclose(got, 0.5 + 4.0 * raw8.astype(np.float64) / 255.0, atol=1e-6)
