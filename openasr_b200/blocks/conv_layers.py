"""Conv-subsampling input block of the encoder with its first layer on the B200 path.

``Conv2dSubsampleV2`` keeps the reference module's constructor, parameter names and forward contract
(``src/blocks/conv_layers.py:122-150``): ``forward(feats[B, T, D], feat_lengths) -> (outputs[B, T', d_model],
output_lengths)``, ``state_dict()`` keys ``conv.subsample/conv{i}.weight|bias`` and ``affine.weight|bias`` -- a
reference checkpoint loads unchanged.  What changes is the execution of ``subsample/conv0`` +
``subsample/relu0``: one hand-written kernel (``csrc/conv0_kernel.cu`` through ``spl_conv0_relu``) that reads the
``[B, T, D]`` features the front-end just produced (still in L2) and writes the ``[B, 32, T1, D1]`` activations,
without the ``unsqueeze(1)`` view / NCHW staging.  The remaining layers (32 -> 32 convolutions, the affine
projection) are library GEMM-shaped work and stay on cuDNN / cuBLAS through torch.

Training: ``conv0_relu`` is a ``torch.autograd.Function``; the backward pass (weight / bias / input gradients)
is expressed with torch's own convolution-gradient operators on the saved input and the ReLU mask.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import torch

from .. import _capi, frontend


def conv0_relu_forward(feats: torch.Tensor, weight: torch.Tensor, bias) -> torch.Tensor:
    """[B, T, D] fp32 CUDA features -> relu(conv2d(feats[:, None], weight, bias, stride=(2, 1))) [B, C, T1, D1]."""
    frontend._require_cuda(feats, "feats")
    if feats.dim() != 3:
        raise ValueError("feats must be [B, T, D]")
    if tuple(weight.shape[1:]) != (1, 3, 3):
        raise ValueError("weight must be [C, 1, 3, 3]")
    x = feats if (feats.dtype == torch.float32 and feats.is_contiguous()) else feats.float().contiguous()
    w = weight.detach().float().contiguous()
    b = bias.detach().float().contiguous() if bias is not None else None
    B, T, D = x.shape
    Cout = w.shape[0]
    if T < 3 or D < 3:
        raise ValueError("conv0 needs T >= 3 and D >= 3, got T=%d D=%d" % (T, D))
    out = torch.empty((B, Cout, (T - 3) // 2 + 1, D - 2), dtype=torch.float32, device=x.device)
    lib = _capi.load()
    with torch.cuda.device(x.device):
        _capi.check(lib.spl_conv0_relu(None, C.c_void_p(x.data_ptr()), B, T, D, C.c_void_p(w.data_ptr()),
                                       C.c_void_p(b.data_ptr()) if b is not None else None, Cout,
                                       C.c_void_p(out.data_ptr()), frontend._stream_ptr(x.device)), "spl_conv0_relu")
    return out


class _Conv0ReLU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, weight, bias):
        out = conv0_relu_forward(feats, weight, bias)
        ctx.save_for_backward(feats, weight, out)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        feats, weight, out = ctx.saved_tensors
        g = grad_out * (out > 0).to(grad_out.dtype)  # ReLU
        x4 = feats.unsqueeze(1).float()
        grad_x = grad_w = grad_b = None
        if ctx.needs_input_grad[0]:
            grad_x = torch.nn.grad.conv2d_input(x4.shape, weight, g, stride=(2, 1)).squeeze(1).to(feats.dtype)
        if ctx.needs_input_grad[1]:
            grad_w = torch.nn.grad.conv2d_weight(x4, weight.shape, g, stride=(2, 1))
        if ctx.has_bias and ctx.needs_input_grad[2]:
            grad_b = g.sum(dim=(0, 2, 3))
        return grad_x, grad_w, grad_b


def conv0_relu(feats: torch.Tensor, weight: torch.Tensor, bias=None) -> torch.Tensor:
    """Differentiable ``relu(conv2d(feats.unsqueeze(1), weight, bias, stride=(2, 1)))`` on the B200 kernel."""
    return _Conv0ReLU.apply(feats, weight, bias)


class Conv2dSubsampleV2(torch.nn.Module):
    """conv_layers.py:122-150 with ``subsample/conv0`` + ``subsample/relu0`` fused into one kernel."""

    def __init__(self, d_input, d_model, layer_num=2):
        super().__init__()
        assert layer_num >= 1
        self.layer_num = layer_num
        layers = [("subsample/conv0", torch.nn.Conv2d(1, 32, 3, (2, 1))),
                  ("subsample/relu0", torch.nn.ReLU())]
        for i in range(layer_num - 1):
            layers += [("subsample/conv{}".format(i + 1), torch.nn.Conv2d(32, 32, 3, (2, 1))),
                       ("subsample/relu{}".format(i + 1), torch.nn.ReLU())]
        self.conv = torch.nn.Sequential(OrderedDict(layers))
        self.affine = torch.nn.Linear(32 * (d_input - 2 * layer_num), d_model)
        self.d_model = d_model

    def forward(self, feats, feat_lengths):
        conv0 = self.conv[0]
        outputs = conv0_relu(feats, conv0.weight, conv0.bias)  # [B, 32, T1, D1]; layers 0 and 1 of self.conv
        for layer in list(self.conv)[2:]:
            outputs = layer(outputs)
        B, Cc, T, D = outputs.size()
        outputs = outputs.permute(0, 2, 1, 3).contiguous().view(B, T, Cc * D)
        outputs = self.affine(outputs)
        output_lengths = feat_lengths
        for _ in range(self.layer_num):
            output_lengths = ((output_lengths - 1) / 2).long()  # conv_layers.py:147-148
        return outputs, output_lengths
