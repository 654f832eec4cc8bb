"""GPU parity of row f2: Conv2d(1 -> 32, 3x3, stride (2, 1)) + ReLU on the front-end's features
(csrc/conv0_kernel.cu through the C ABI) against the reference's own outputs and the CPU oracle.
fp32 tolerance: |d| <= 1e-4 + 1e-5 |ref| (nine-term dot products of O(10) features)."""
import os

import numpy as np
import pytest
import torch

from oracle import conv_oracle as co
from oracle import frontend_oracle as fo

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_library_layers():
    """The layers AFTER conv0 run on cuDNN / cuBLAS through torch; keep them in true fp32 for the comparisons."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def close(a, b, atol=1e-4, rtol=1e-5):
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    assert a.shape == b.shape, (a.shape, b.shape)
    d = (a - b).abs()
    assert (d <= atol + rtol * b.abs()).all(), d.max().item()


def load_ref(golden_dir):
    g = np.load(os.path.join(golden_dir, "conv_ref.npz"))
    state = {k[len("param:"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("param:")}
    return g, state


def test_conv0_matches_reference_activations(golden_dir):
    from openasr_b200.blocks.conv_layers import Conv2dSubsampleV2, conv0_relu_forward
    g, state = load_ref(golden_dir)
    feats = torch.from_numpy(g["feats"]).cuda()
    act0 = conv0_relu_forward(feats, state["conv.subsample/conv0.weight"].cuda(), state["conv.subsample/conv0.bias"].cuda())
    close(act0, torch.from_numpy(g["act0"]))
    assert (act0 >= 0).all()
    # whole module from the reference's state_dict (same keys): output and lengths
    m = Conv2dSubsampleV2(80, 24, layer_num=2).cuda().eval()
    m.load_state_dict(state, strict=True)
    with torch.no_grad():
        out, olen = m(feats, torch.from_numpy(g["lengths"]).cuda())
    assert torch.equal(olen.cpu(), torch.from_numpy(g["out_lengths"]))
    close(out, torch.from_numpy(g["out"]), atol=5e-4, rtol=1e-4)  # 2 432-term affine on top (TF32-free cuBLAS fp32)


@pytest.mark.parametrize("B,T,D,C", [(32, 649, 80, 32), (64, 398, 40, 32), (3, 3, 3, 1), (2, 130, 81, 64), (5, 1000, 24, 7)])
def test_conv0_shapes_vs_oracle(B, T, D, C):
    """BASELINE shapes (AISHELL 32 x 649 x 80, HKUST 64 x 398 x 40) and edge shapes: minimum 3 x 3 input,
    odd D, C = 1 / 64 / 7, tile tails; without bias too."""
    from openasr_b200.blocks.conv_layers import conv0_relu_forward
    gen = torch.Generator().manual_seed(B * 1000 + T)
    x = 4.0 * torch.randn(B, T, D, generator=gen) + 8.0
    x[:, T - T // 5:] = 0.0  # zero padding rows like a ragged batch
    w = 0.3 * torch.randn(C, 1, 3, 3, generator=gen)
    b = 0.5 * torch.randn(C, generator=gen)
    close(conv0_relu_forward(x.cuda(), w.cuda(), b.cuda()), co.conv0_relu(x, w, b))
    close(conv0_relu_forward(x.cuda(), w.cuda(), None), co.conv0_relu(x, w, None))


def test_conv0_gradients_match_torch():
    from openasr_b200.blocks.conv_layers import conv0_relu
    gen = torch.Generator().manual_seed(9)
    x = (torch.randn(3, 21, 12, generator=gen)).cuda().requires_grad_(True)
    w = (0.3 * torch.randn(32, 1, 3, 3, generator=gen)).cuda().requires_grad_(True)
    b = (0.1 * torch.randn(32, generator=gen)).cuda().requires_grad_(True)
    go = torch.randn(3, 32, 10, 10, generator=gen).cuda()
    conv0_relu(x, w, b).backward(go)
    x2, w2, b2 = (t.detach().clone().requires_grad_(True) for t in (x, w, b))
    torch.nn.functional.relu(torch.nn.functional.conv2d(x2.unsqueeze(1), w2, b2, stride=(2, 1))).backward(go)
    close(x.grad, x2.grad, atol=1e-4, rtol=1e-4)
    close(w.grad, w2.grad, atol=1e-3, rtol=1e-4)
    close(b.grad, b2.grad, atol=1e-3, rtol=1e-4)


def test_frontend_to_conv0_pipeline(wavs):
    """SPLayer features feed conv0 directly: end of the path of SURVEY 8a into row f2."""
    from openasr_b200 import SPLayer
    from openasr_b200.blocks.conv_layers import Conv2dSubsampleV2
    conf = {"feature_type": "fbank", "sample_rate": 16000, "num_mel_bins": 80, "use_energy": False, "dither": 0.0}
    layer = SPLayer(conf).cuda().eval()
    lens = [wavs[0].shape[0], wavs[1].shape[0]]
    x = torch.zeros(2, max(lens))
    x[0, :lens[0]] += wavs[0]
    x[1, :lens[1]] += wavs[1]
    feats, flen = layer(x.cuda(), lens)
    torch.manual_seed(1)
    m = Conv2dSubsampleV2(80, 16).cuda().eval()
    with torch.no_grad():
        out, olen = m(feats, flen)
    ref_f, ref_l = fo.splayer_forward(x, lens, conf)
    state = {k: v.cpu() for k, v in m.state_dict().items()}
    ref_out, ref_olen = co.conv2d_subsample_v2(state, ref_f, ref_l, 2)
    assert torch.equal(olen.cpu(), ref_olen)
    close(out, ref_out, atol=5e-3, rtol=1e-3)  # fbank tolerance (1e-3) propagated through two convs + affine

    with pytest.raises(RuntimeError):
        m.cpu()(ref_f, ref_l)  # no CPU path
