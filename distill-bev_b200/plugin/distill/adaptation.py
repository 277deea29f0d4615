"""Student adaptation layers on the B200 kernels.

``Conv1x1Adaptation`` mirrors the '1x1conv' adaptation of BEVDetDistill
(mmdet3d/models/detectors/bevdet_distill.py:216-351: ``nn.Conv2d(student_channel,
teacher_channel, kernel_size=1)``; same parameter names ``weight`` [Ct, Cs, 1, 1] / ``bias``).
Forward runs the tcgen05 TF32 GEMM of csrc/adapt_gemm.cu; the backward (input and weight
gradients) currently goes through cuDNN via torch - library code, to be replaced.
"""
import torch
from torch import nn

from ... import _lib
from ..ops.bev_pool import transpose_batched


def conv1x1_forward(x, weight, bias):
    """y = conv1x1(x) [B, Cout, H, W] on the tcgen05 GEMM (no autograd); also returns the x that
    a backward should save."""
    lib = _lib.load()
    _lib.require_cuda(x, "x", torch.float32)
    _lib.require_cuda(weight, "weight", torch.float32)
    x, weight = x.detach(), weight.detach()
    B, Cin, H, W = x.shape
    Cout = weight.shape[0]
    if x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous():
        x_cl = x                                   # [B, HW, Cin] in memory already
    else:
        x = x.contiguous()
        x_cl = transpose_batched(x, B, Cin, H * W)  # NCHW -> channels-last rows
    w2 = weight.reshape(Cout, Cin).contiguous()
    y = torch.empty((B, Cout, H, W), dtype=torch.float32, device=x.device)
    b = bias.detach().contiguous() if bias is not None else None
    with torch.cuda.device(x.device):
        rc = lib.dbev_adapt_conv1x1_forward(_lib.ptr(x_cl), _lib.ptr(w2), _lib.ptr(b), B, Cin, Cout,
                                            H * W, _lib.ptr(y), _lib.stream_ptr(x.device))
    _lib.check(rc, "dbev_adapt_conv1x1_forward")
    return y, x


class _Conv1x1(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, weight, bias):
        y, x = conv1x1_forward(x, weight, bias)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gy = gy.contiguous()
        gx = torch.nn.grad.conv2d_input(x.shape, weight, gy) if ctx.needs_input_grad[0] else None
        gw = torch.nn.grad.conv2d_weight(x, weight.shape, gy) if ctx.needs_input_grad[1] else None
        gb = gy.sum(dim=(0, 2, 3)) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return gx, gw, gb


def conv1x1(x, weight, bias=None):
    return _Conv1x1.apply(x, weight, bias)


class Conv1x1Adaptation(nn.Conv2d):
    """Drop-in for nn.Conv2d(Cs, Ct, kernel_size=1): same parameters / state_dict."""

    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__(in_channels, out_channels, kernel_size=1, bias=bias)

    def forward(self, x):
        return conv1x1(x, self.weight, self.bias)
