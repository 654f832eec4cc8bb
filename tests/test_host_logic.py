"""CPU: host-side logic of the drop-in (tables, frame arithmetic, SpecAug rectangles, module API,
C-ABI exports).  No compute calls -- there is no GPU here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from oracle import frontend_oracle as fo


def test_library_loads_and_exports_every_declared_symbol():
    from openasr_b200 import _capi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "spl_capi.h")).read()
    declared = set(re.findall(r"SPL_API\s+[\w\s\*]+?\b(spl_\w+)\s*\(", hdr))
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    lib = _capi.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.spl_abi_version() == _capi.SPL_ABI_VERSION
    assert ctypes.sizeof(_capi.SplConfig) == 36
    # argument validation paths that need no device
    assert lib.spl_fbank_forward(None, None, None) != 0
    assert b"null" in lib.spl_last_error()


def test_header_is_plain_c_and_struct_layouts_match_ctypes(tmp_path):
    """include/spl_capi.h compiles as C99 and as C++17 (no torch / CUDA types in the signatures), and the ctypes
    mirrors in openasr_b200/_capi.py have the size and field offsets the C compiler gives the structs."""
    import shutil
    import subprocess
    from openasr_b200 import _capi
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = os.path.join(root, "include", "spl_capi.h")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr], check=True)
    structs = {"spl_config": _capi.SplConfig, "spl_fbank_args": _capi.SplFbankArgs, "spl_post_args": _capi.SplPostArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "spl_capi.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(root, "include"), "-o", str(exe), str(src)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(out[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_tables_bit_identical_to_oracle():
    from openasr_b200 import tables
    for sr in (16000.0, 8000.0, 11025.0):
        S, Nw, Nfft = tables.frame_geometry(sr)
        assert (S, Nw, Nfft) == fo.frame_params(sr)
        for wt in tables.WINDOW_TYPES:
            assert torch.equal(tables.window_table(wt, Nw), fo.window_function(wt, Nw))
        for D in (23, 40, 80):
            assert torch.equal(tables.mel_table(D, Nfft, sr), fo.mel_banks(D, Nfft, sr))
    assert tables.frame_geometry(16000.0) == (160, 400, 512)
    assert tables.frame_geometry(8000.0) == (80, 200, 256)
    m = tables.mel_table(80, 512, 16000.0)
    assert int((m != 0).sum()) == 501 and float(m[:, 0].abs().max()) == 0.0
    assert int((tables.mel_table(40, 256, 8000.0) != 0).sum()) == 247
    assert abs(float(tables.window_table("povey", 400).sum()) - 212.14699) < 1e-3


def test_frame_count_matches_oracle():
    from openasr_b200 import tables
    for n in (399, 400, 559, 560, 561, 32640, 67263):
        assert tables.frame_count(n, 400, 160) == fo.num_frames(n, 400, 160)
    assert tables.frame_count(399, 400, 160) == 0 and tables.frame_count(560, 400, 160) == 2


def test_specaug_rectangles_equal_reference_slicing():
    """Closed form (rectangles + time/freq means) == the sequential in-place reference code,
    including overlapping masks, negative starts and spills into padding."""
    from openasr_b200 import frontend
    g = torch.Generator().manual_seed(0)
    B, T, V = 6, 90, 24
    lens = torch.tensor([90, 3, 45, 60, 10, 77])
    feats = torch.randn(B, T, V, generator=g)
    for i, l in enumerate(lens.tolist()):
        feats[i, l:] = 0
    for trial, conf in enumerate(({"freq_mask_num": 2, "freq_mask_width": 10, "time_mask_num": 2, "time_mask_width": 40},
                                  {"freq_mask_num": 1, "freq_mask_width": 30, "time_mask_num": 3, "time_mask_width": 120},
                                  {"freq_mask_num": 0, "freq_mask_width": 5, "time_mask_num": 1, "time_mask_width": 8})):
        for seed in range(20):
            torch.manual_seed(100 * trial + seed)
            uni = frontend.specaug_uniforms(B, conf["freq_mask_num"], conf["time_mask_num"])
            ref, _ = fo.spec_aug(feats.clone(), lens, conf, uniforms=uni)
            rect = frontend.specaug_rectangles(uni, lens, T, V, conf)
            assert rect.shape == (B, conf["freq_mask_num"] + conf["time_mask_num"], 2) and rect.dtype == torch.int32
            fm = feats.mean(-1)
            tm = feats.sum(1) / lens[:, None].float()
            out = feats.clone()
            nf = conf["freq_mask_num"]
            for b in range(B):
                for j in range(nf):
                    s, e = rect[b, j].tolist()
                    out[b, :, s:e] = fm[b][:, None]
                for j in range(nf, rect.shape[1]):
                    s, e = rect[b, j].tolist()
                    out[b, s:e, :] = tm[b][None, :]
            assert torch.equal(out, ref)


def test_numpy_rectangles_equal_torch_rectangles():
    from openasr_b200 import frontend
    lens = torch.tensor([90, 3, 45, 60, 10, 77, 1, 500])
    for conf in ({"freq_mask_num": 2, "freq_mask_width": 27, "time_mask_num": 2, "time_mask_width": 40},
                 {"freq_mask_num": 1, "freq_mask_width": 100.0, "time_mask_num": 3, "time_mask_width": 700},
                 {"freq_mask_num": 0, "freq_mask_width": 5, "time_mask_num": 1, "time_mask_width": 8}):
        for seed in range(50):
            torch.manual_seed(seed)
            u = frontend.specaug_uniforms(8, conf["freq_mask_num"], conf["time_mask_num"])
            a = frontend.specaug_rectangles(u, lens, 500, 80, conf)
            b = frontend.specaug_rectangles_np(u.numpy(), lens.numpy(), 500, 80, conf)
            assert np.array_equal(a.numpy(), b)
            assert np.array_equal(b, frontend.specaug_rectangles_c(u.numpy(), lens.numpy(), 500, 80, conf))


def test_specaug_uniform_stream_equals_sequential_draws():
    from openasr_b200 import frontend
    torch.manual_seed(4)
    a = frontend.specaug_uniforms(7, 2, 2)
    torch.manual_seed(4)
    b = torch.stack([torch.rand(size=[7]) for _ in range(8)])
    assert torch.equal(a, b)


def test_splayer_module_contract():
    from openasr_b200 import SPLayer, WavConv
    from openasr_b200.blocks import sp_layers
    conf = {"feature_type": "fbank", "sample_rate": 16000, "num_mel_bins": 80, "use_energy": False,
            "spec_aug": {"freq_mask_num": 2, "freq_mask_width": 27, "time_mask_num": 2, "time_mask_width": 40}}
    keys = dict(conf)
    layer = SPLayer(conf)
    assert layer.config is conf and conf == keys  # un-mutated (Speech_Models.restore compares key by key)
    assert layer.feature_type == "fbank" and layer.spec_aug_conf == conf["spec_aug"]
    assert len(layer.state_dict()) == 0 and len(list(layer.parameters())) == 0
    layer.load_state_dict({}, strict=True)
    assert SPLayer({"feature_type": "offline"}).spec_aug_conf is None
    with pytest.raises(ValueError, match="Unknown feature type"):
        SPLayer({"feature_type": "spectrogram"})
    with pytest.raises(ValueError):
        SPLayer({"feature_type": "fbank", "sample_rate": 16000, "num_mel_bins": 80, "use_energy": False, "cmvn": "x"})
    # no CPU fallback: a CPU waveform is a hard error, never a silent slow path
    with pytest.raises(RuntimeError, match="no CPU path"):
        layer(torch.zeros(1, 16000), [16000])
    # offline pass-through in eval mode needs no device
    off = SPLayer({"feature_type": "offline"}).eval()
    x = torch.randn(2, 5, 4)
    y, l = off(x, torch.tensor([5, 3]))
    assert y is x and l.dtype == torch.int64
    assert hasattr(sp_layers, "WavConv")
    names = [k for k in WavConv({"d_model": 4}).state_dict().keys() if k.endswith("weight")]
    assert names[0] == "encoder.0.weight" and "encoder.12.weight" in names


def test_wavconv_matches_reference_module():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("/root/reference not mounted")
    _, _ = ref_shim.load()
    from blocks.sp_layers import WavConv as RefWavConv  # the reference's
    from openasr_b200 import WavConv
    torch.manual_seed(0)
    ref = RefWavConv({"d_model": 16}).eval()
    ours = WavConv({"d_model": 16}).eval()
    ours.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(2, 4800)
    l = torch.tensor([4800, 3200])
    a, la = ref(x, l)
    b, lb = ours(x, l)
    assert torch.equal(a, b) and torch.equal(la, lb)


# --------------------------------------------------------------------------------------------- batching (row f4)
def _reference_time_batches(lengths, duration, ngpu):
    """The rule of src/dataload/samplers.py:15-32 restated on a plain list (test-side oracle)."""
    out, batch, dur = [], [], 0.0
    for idx, n in enumerate(lengths):
        batch.append(idx)
        dur += n
        if dur >= duration and len(batch) % ngpu == 0:
            out.append(batch)
            batch, dur = [], 0.0
    if batch:
        out.append(batch if len(batch) % ngpu == 0 else batch[len(batch) // ngpu * ngpu:])
    return out


def test_time_based_sampler_matches_reference_rule():
    from openasr_b200.batching import BucketedTimeSampler, TimeBasedSampler
    rng = np.random.default_rng(0)
    lengths = rng.uniform(1.0, 15.0, size=203).tolist()
    data = [{"feat_length": n} for n in lengths]
    for ngpu in (1, 2, 4):
        s = TimeBasedSampler(data, duration=60, ngpu=ngpu)
        assert list(s) == _reference_time_batches(lengths, 60, ngpu)
        assert len(s) == len(_reference_time_batches(lengths, 60, ngpu))
    ref_path = "/root/reference/src/dataload/samplers.py"
    if os.path.exists(ref_path):  # the unmodified reference class, when the mount is present
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_samplers", ref_path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        for ngpu in (1, 2, 4):
            assert list(mod.TimeBasedSampler(data, duration=60, ngpu=ngpu)) == list(TimeBasedSampler(data, 60, ngpu))
    b = BucketedTimeSampler(data, duration=60, ngpu=2, num_buckets=6)
    seen = [i for batch in b for i in batch]
    assert len(seen) == len(set(seen)) and set(seen) <= set(range(len(data)))
    plain = TimeBasedSampler(data, duration=60, ngpu=2)
    pad_plain = 1.0 - sum(lengths[i] for bt in plain.batchs for i in bt) / sum(max(lengths[i] for i in bt) * len(bt) for bt in plain.batchs)
    assert b.padding_fraction() < 0.5 * pad_plain


def test_pad_wave_batch_matches_reference_collate():
    from openasr_b200.batching import pad_wave_batch
    rng = np.random.default_rng(1)
    ws16 = [rng.integers(-2000, 2000, size=n).astype(np.int16) for n in (400, 1234, 777)]
    x, lens = pad_wave_batch(ws16)
    assert x.dtype == torch.int16 and lens.dtype == torch.int64 and lens.tolist() == [400, 1234, 777]
    # reference semantics (data_utils.py:134-138): zeros + per-row add, float32
    ref = torch.zeros(3, 1234)
    for i, w in enumerate(ws16):
        ref[i, :len(w)] += torch.from_numpy(w.astype(np.float32))
    assert torch.equal(x.float(), ref)
    xf, _ = pad_wave_batch([w.astype(np.float32) for w in ws16])
    assert xf.dtype == torch.float32 and torch.equal(xf, ref)
    out = torch.full((4, 2000), 7, dtype=torch.int16)
    xo, _ = pad_wave_batch(ws16, out=out)
    assert torch.equal(xo, x) and xo.data_ptr() == out.data_ptr()
    with pytest.raises(ValueError):
        pad_wave_batch(ws16, out=torch.zeros(2, 2000, dtype=torch.int16))
