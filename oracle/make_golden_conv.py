"""Generate tests/golden/conv_ref.npz by running the UNMODIFIED reference ``Conv2dSubsampleV2``
(/root/reference/src/blocks/conv_layers.py:122-150) in the build container.

TEST INFRASTRUCTURE.  Usage (where /root/reference is mounted):  python -m oracle.make_golden_conv

Stored: the module's parameters (seeded init), a feature batch taken from the committed fbank vectors
(wav0 / wav1, zero-padded like SPLayer pads), its lengths, the activations after the first
Conv2d + ReLU (``module.conv[0:2]``) and the module output / output lengths.  Every output array comes
from the reference class, never from oracle/conv_oracle.py.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

REF_SRC = "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    torch.set_num_threads(1)
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    from blocks.conv_layers import Conv2dSubsampleV2  # the reference class, imported where it lies
    fb = np.load(os.path.join(OUT, "fbank_ref.npz"))
    f0, f1 = torch.from_numpy(fb["w0_d80"])[:61], torch.from_numpy(fb["w1_d80"])[:44]
    feats = torch.zeros(2, 61, 80)
    feats[0, :61] += f0
    feats[1, :44] += f1
    lens = torch.tensor([61, 44]).long()
    torch.manual_seed(5)
    m = Conv2dSubsampleV2(80, 24, layer_num=2).eval()
    with torch.no_grad():
        act0 = m.conv[1](m.conv[0](feats.unsqueeze(1)))
        out, olen = m(feats, lens)
    arrays = {"feats": feats.numpy(), "lengths": lens.numpy(), "act0": act0.numpy(), "out": out.numpy(),
              "out_lengths": olen.numpy()}
    for k, v in m.state_dict().items():
        arrays["param:" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "conv_ref.npz"), **arrays)
    print({k: v.shape for k, v in arrays.items()})


if __name__ == "__main__":
    main()
