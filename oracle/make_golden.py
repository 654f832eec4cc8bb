"""Generate tests/golden/* by running the UNMODIFIED reference in the build container.

TEST INFRASTRUCTURE.  Usage (here, where /root/reference is mounted):

    python -m oracle.make_golden

Outputs (committed; the GPU box has no /root/reference):
  tests/golden/wav0.npy, wav1.npy      int16 PCM of the reference's two WAV fixtures
                                       (test/testdata/100-121669-0000.wav, BAC009S0764W0121.wav)
  tests/golden/fbank_ref.npz           reference ``kaldi_signal.fbank`` outputs per config
  tests/golden/specaug_ref.npz         reference ``SPLayer.spec_aug`` output + the uniforms drawn
  tests/golden/known_answers.json      scalar known answers

Every array in fbank_ref.npz / specaug_ref.npz is produced by code imported from
/root/reference (oracle/ref_shim.py), never by the oracle restatement.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from oracle import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (wav index, sample_rate, decimate, D, use_energy, window, dither, seed)
FBANK_CASES = {
    "w0_d80": (0, 16000.0, 1, 80, False, "povey", 0.0, None),
    "w1_d80": (1, 16000.0, 1, 80, False, "povey", 0.0, None),
    "w0_d40": (0, 16000.0, 1, 40, False, "povey", 0.0, None),
    "w1_d40": (1, 16000.0, 1, 40, False, "povey", 0.0, None),
    "w0_d80_energy": (0, 16000.0, 1, 80, True, "povey", 0.0, None),
    "w1_d80_hamming": (1, 16000.0, 1, 80, False, "hamming", 0.0, None),
    "w1_8k_d40": (1, 8000.0, 2, 40, False, "povey", 0.0, None),
    "w0_d80_dither_seed7": (0, 16000.0, 1, 80, False, "povey", 1.0, 7),
    "w1_8k_d40_energy_dither_seed11": (1, 8000.0, 2, 40, True, "povey", 1.0, 11),
}


def main():
    torch.set_num_threads(1)
    ksp, RefSPLayer = ref_shim.load()
    os.makedirs(OUT, exist_ok=True)
    wavs = []
    for i, name in enumerate(["100-121669-0000.wav", "BAC009S0764W0121.wav"]):
        sr, w = ref_shim.read_wav(name)
        assert sr == 16000
        pcm = w.astype(np.int16)
        assert np.array_equal(pcm.astype(np.float32), w)
        np.save(os.path.join(OUT, "wav%d.npy" % i), pcm)
        wavs.append(w)

    fb = {}
    known = {}
    for key, (wi, sr, dec, D, en, wt, dither, seed) in FBANK_CASES.items():
        x = torch.from_numpy(wavs[wi][::dec].copy())
        if seed is not None:
            torch.manual_seed(seed)
        f = ksp.fbank(x.view(1, -1), sample_frequency=sr, use_energy=en, num_mel_bins=D,
                      dither=dither, window_type=wt)
        fb[key] = f.numpy()
        known[key] = {
            "shape": list(f.shape), "first": float(f[0, 0]), "mid": float(f[100, D // 2]),
            "last": float(f[-1, -1]), "sum": float(f.sum(dtype=torch.float64)),
        }
    # all-zero input -> log(eps) everywhere
    for n in (400, 559, 560):
        f = ksp.fbank(torch.zeros(1, n), num_mel_bins=80, dither=0.0)
        known["zeros_%d" % n] = {"shape": list(f.shape), "first": float(f[0, 0]),
                                 "all_equal": bool((f == f[0, 0]).all())}
    np.savez_compressed(os.path.join(OUT, "fbank_ref.npz"), **fb)

    # SpecAug: reference SPLayer.spec_aug on the padded batch [wav0, wav1], D=80, seed 0
    conf = {"feature_type": "fbank", "sample_rate": 16000, "num_mel_bins": 80, "use_energy": False,
            "spec_aug": {"freq_mask_num": 2, "freq_mask_width": 27, "time_mask_num": 2,
                         "time_mask_width": 40}}
    layer = RefSPLayer(conf)
    f0, f1 = torch.from_numpy(fb["w0_d80"]), torch.from_numpy(fb["w1_d80"])
    T = max(f0.shape[0], f1.shape[0])
    pad = torch.zeros(2, T, 80)
    pad[0, :f0.shape[0]] += f0
    pad[1, :f1.shape[0]] += f1
    lens = torch.tensor([f0.shape[0], f1.shape[0]]).long()
    torch.manual_seed(0)
    uni = torch.rand(8, 2)  # == 8 sequential torch.rand(size=[2]) draws (CPU mt19937 stream)
    torch.manual_seed(0)
    aug, _ = layer.spec_aug(pad.clone(), lens)
    # second case: wide time masks (T=100) incl. the len<width quirk on a short utterance
    conf2 = dict(conf, spec_aug={"freq_mask_num": 1, "freq_mask_width": 15, "time_mask_num": 3,
                                 "time_mask_width": 300})
    layer2 = RefSPLayer(conf2)
    torch.manual_seed(3)
    uni2 = torch.rand(8, 2)
    torch.manual_seed(3)
    aug2, _ = layer2.spec_aug(pad.clone(), lens)
    np.savez_compressed(os.path.join(OUT, "specaug_ref.npz"), aug_seed0=aug.numpy(), uniforms_seed0=uni.numpy(),
                        aug2_seed3=aug2.numpy(), uniforms2_seed3=uni2.numpy(), lengths=lens.numpy())
    known["specaug_seed0_sum"] = float(aug.sum(dtype=torch.float64))
    known["specaug2_seed3_sum"] = float(aug2.sum(dtype=torch.float64))
    known["torch_version"] = torch.__version__
    with open(os.path.join(OUT, "known_answers.json"), "w") as fh:
        json.dump(known, fh, indent=1, sort_keys=True)
    for k, v in known.items():
        print(k, v)


if __name__ == "__main__":
    main()
