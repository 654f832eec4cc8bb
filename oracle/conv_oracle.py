"""CPU restatement of the first conv-subsampling layer.

TEST INFRASTRUCTURE ONLY -- imported by tests/ and by bench.py's cpu side; the product path never
routes through it.

Follows ``src/blocks/conv_layers.py:122-150`` (``Conv2dSubsampleV2``): ``feats.unsqueeze(1)`` (:139),
``Conv2d(1, 32, 3, (2, 1))`` + ``ReLU`` (:125-126), the remaining ``Conv2d(32, 32, 3, (2, 1))`` + ``ReLU``
layers (:128-131), ``permute(0, 2, 1, 3).contiguous().view(B, T, C*D)`` (:142-143), the affine layer
(:145) and the length rule ``((len - 1) / 2).long()`` per layer (:147-148).

Pinned: ``oracle/make_golden.py`` runs the UNMODIFIED reference class and stores its parameters, input and
outputs in ``tests/golden/conv_ref.npz``; ``tests/test_oracle_golden.py`` checks this restatement against them.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def conv0_relu(feats: torch.Tensor, weight: torch.Tensor, bias=None) -> torch.Tensor:
    """conv_layers.py:125-126 + :139 on ``[B, T, D]`` features -> ``[B, C, (T-3)//2+1, D-2]``."""
    return F.relu(F.conv2d(feats.unsqueeze(1), weight, bias, stride=(2, 1)))


def subsample_lengths(lengths: torch.Tensor, layer_num: int) -> torch.Tensor:
    """conv_layers.py:146-148 (true division, then truncation)."""
    out = lengths
    for _ in range(layer_num):
        out = ((out - 1) / 2).long()
    return out


def conv2d_subsample_v2(state: dict, feats: torch.Tensor, lengths: torch.Tensor, layer_num: int):
    """Whole-module forward from a reference ``state_dict`` (conv_layers.py:138-150)."""
    x = feats.unsqueeze(1)
    for i in range(layer_num):
        x = F.relu(F.conv2d(x, state["conv.subsample/conv%d.weight" % i], state["conv.subsample/conv%d.bias" % i],
                            stride=(2, 1)))
    B, C, T, D = x.shape
    x = x.permute(0, 2, 1, 3).contiguous().view(B, T, C * D)
    x = F.linear(x, state["affine.weight"], state["affine.bias"])
    return x, subsample_lengths(lengths, layer_num)
