"""Writes tests/golden/feats.ark / feats.scp / ark_ref.npz: a small synthetic Kaldi archive (binary float, binary
double, compressed CM, text) and what the UNMODIFIED reference reader (src/third_party/kaldi_io.py:362-448,
imported from /root/reference) decodes from it.  Run in the build container (the reference is not on the GPU box):
    python oracle/make_golden_ark.py
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
G = os.path.join(ROOT, "tests", "golden")


def compress_cm(mat):
    """Kaldi CompressedMatrix 'CM ' encoder (compressed-matrix.cc), enough for a fixture."""
    rows, cols = mat.shape
    gmin, gmax = float(mat.min()), float(mat.max())
    grange = max(gmax - gmin, 1e-5)
    def to_u16(v):
        return int(np.clip(round((v - gmin) / grange * 65535.0), 0, 65535))
    out = b"\0B" + b"CM " + struct.pack("<ffii", gmin, grange, rows, cols)
    hdrs, datas = b"", b""
    for c in range(cols):
        col = np.sort(mat[:, c])
        q = [col[0], col[rows // 4], col[3 * rows // 4], col[-1]]
        p = [to_u16(v) for v in q]
        p[1] = max(p[1], p[0] + 1); p[2] = max(p[2], p[1] + 1); p[3] = max(p[3], p[2] + 1)
        hdrs += struct.pack("<HHHH", *p)
        pf = [gmin + grange * x / 65535.0 for x in p]
        v = mat[:, c]
        d = np.where(v < pf[1], np.clip(np.floor((v - pf[0]) / (pf[1] - pf[0]) * 64 + 0.5), 0, 64),
                     np.where(v < pf[2], np.clip(np.floor(64 + (v - pf[1]) / (pf[2] - pf[1]) * 128 + 0.5), 64, 192),
                              np.clip(np.floor(192 + (v - pf[2]) / (pf[3] - pf[2]) * 63 + 0.5), 192, 255)))
        datas += d.astype(np.uint8).tobytes()
    return out + hdrs + datas


def main():
    from openasr_b200 import ark_io
    rng = np.random.RandomState(7)
    mats = {"utt_f32": (4.0 * rng.randn(37, 40) + 8.0).astype(np.float32),
            "utt_f64": (4.0 * rng.randn(11, 40) + 8.0).astype(np.float64),
            "utt_cm": (4.0 * rng.randn(53, 40) + 8.0).astype(np.float32),
            "utt_txt": np.round(4.0 * rng.randn(5, 40) + 8.0, 3).astype(np.float32)}
    ark = os.path.join(G, "feats.ark")
    lines = []
    with open(ark, "wb") as fd:
        for key in ("utt_f32", "utt_f64"):
            lines.append("%s tests/golden/feats.ark:%d" % (key, ark_io.write_mat(fd, key, mats[key])))
        fd.write(b"utt_cm ")
        lines.append("utt_cm tests/golden/feats.ark:%d" % fd.tell())
        fd.write(compress_cm(mats["utt_cm"]))
        fd.write(b"utt_txt ")
        lines.append("utt_txt tests/golden/feats.ark:%d" % fd.tell())
        fd.write(b" [\n" + b"\n".join((" " + " ".join("%g" % v for v in row)).encode() for row in mats["utt_txt"]) + b" ]\n")
    with open(os.path.join(G, "feats.scp"), "w") as f:
        f.write("\n".join(lines) + "\n")
    # decode with the unmodified reference reader
    sys.path.insert(0, "/root/reference/src")
    from third_party import kaldi_io as kio
    ref = {}
    for line in lines:
        key, rx = line.split()
        ref[key] = np.array(kio.read_mat(os.path.join(ROOT, rx)))
    np.savez(os.path.join(G, "ark_ref.npz"), **ref)
    for k, v in ref.items():
        print(k, v.shape, v.dtype, float(np.abs(v - mats[k]).max()))


if __name__ == "__main__":
    main()
