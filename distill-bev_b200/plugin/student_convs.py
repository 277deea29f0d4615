"""Trainable 2D convolutions of the student's BEV encoder with the forward on the tcgen05 conv kernels.

Reference rows: ResNetForBEVDet.forward mmdet3d/models/backbones/resnet.py:51-62 (+ BasicBlock
bricks/res_block.py:70-99) and FPN_LSS.forward mmdet3d/models/necks/lss_fpn.py:62-72 - plain nn.Conv2d + training
BatchNorm + ReLU through cuDNN. Here the convolution itself (the FLOPs) runs on csrc/conv2d_tc.cu in the forward:
NHWC fp32 activations, TF32 multiply / fp32 accumulate (what cuDNN does under torch's default cudnn.allow_tf32),
C_out split into 256 / 128 / 64-column launches that write their channel slice of one NHWC output; BatchNorm
(batch statistics) and ReLU stay torch ops on the channels_last result. The backward (input and weight gradients)
is aten::convolution_backward - cuDNN - on the same channels_last tensors: forward only is replaced, stated in
DESIGN.md. No CPU path: a CPU tensor goes through nn.Conv2d unchanged.
"""
import torch
import torch.nn as nn

from .dense_teacher import conv_nhwc


def _splits(c_out):
    """C_out as a sum of the kernel's N sizes (256, 128, 64), e.g. 512 -> 256 + 256, 384 -> 256 + 128."""
    parts, left = [], c_out
    for n in (256, 128, 64):
        while left >= n:
            parts.append(n)
            left -= n
    return parts if left == 0 else None


def conv2d_tc_supported(weight, stride, padding, dilation=(1, 1), groups=1):
    co, ci, kh, kw = weight.shape
    return (groups == 1 and tuple(dilation) == (1, 1) and ci % 32 == 0 and _splits(co) is not None and kh == kw
            and kh in (1, 2, 3) and stride[0] == stride[1] and stride[0] in (1, 2) and padding[0] == padding[1]
            and padding[0] in (0, 1))


class _Conv2dTcFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, stride, padding):
        co, ci, kh, kw = weight.shape
        xh = x.permute(0, 2, 3, 1).contiguous()                 # free for channels_last activations
        n, h, w, _ = xh.shape
        ho, wo = (h + 2 * padding - kh) // stride + 1, (w + 2 * padding - kw) // stride + 1
        wp = weight.detach().permute(0, 2, 3, 1).reshape(co, kh * kw * ci).contiguous()
        out = torch.empty((n, ho, wo, co), dtype=torch.float32, device=x.device)
        c0 = 0
        for part in _splits(co):
            conv_nhwc(xh, wp[c0:c0 + part], part, kh, kw, stride, padding, out=out, c_off=c0)
            c0 += part
        ctx.save_for_backward(x, weight)
        ctx.conf = (stride, padding)
        return out.permute(0, 3, 1, 2)                           # NCHW shape, channels_last memory

    @staticmethod
    def backward(ctx, grad):
        x, weight = ctx.saved_tensors
        stride, padding = ctx.conf
        gx, gw, _ = torch.ops.aten.convolution_backward(
            grad.contiguous(memory_format=torch.channels_last), x.contiguous(memory_format=torch.channels_last), weight,
            None, [stride, stride], [padding, padding], [1, 1], False, [0, 0], 1,
            [ctx.needs_input_grad[0], ctx.needs_input_grad[1], False])
        return gx, gw, None, None


def conv2d_tc(x, weight, bias=None, stride=1, padding=0):
    """F.conv2d(x, weight, bias, stride, padding) with the forward on tcgen05 (CUDA fp32 only)."""
    out = _Conv2dTcFn.apply(x, weight, int(stride), int(padding))
    if bias is not None:
        out = out + bias.view(1, -1, 1, 1)
    return out


class Conv2dTC(nn.Conv2d):
    """nn.Conv2d whose CUDA fp32 forward runs on the tcgen05 conv kernels when the shape is supported."""

    def forward(self, x):
        if (x.is_cuda and x.dtype == torch.float32 and self.padding_mode == "zeros" and not isinstance(self.padding, str)
                and conv2d_tc_supported(self.weight, self.stride, self.padding, self.dilation, self.groups)
                and x.shape[2] * x.shape[3] >= 128):
            return conv2d_tc(x, self.weight, self.bias, self.stride[0], self.padding[0])
        return super().forward(x)


def convert_convs(module):
    """Re-class every nn.Conv2d of a module tree to Conv2dTC in place (parameters, hooks, state_dict keys unchanged)."""
    for m in module.modules():
        if type(m) is nn.Conv2d:
            m.__class__ = Conv2dTC
    return module
