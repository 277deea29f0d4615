"""Frozen LiDAR teacher's dense BEV backbone and neck on the tcgen05 conv kernel (csrc/conv2d_tc.cu),
same registry names, constructor arguments and ``state_dict`` keys as the reference:

  SECOND      mmdet3d/models/backbones/second.py:12-93     (blocks.{i}.{3j}.weight, blocks.{i}.{3j+1}.*)
  SECONDFPN   mmdet3d/models/necks/second_fpn.py:12-93     (deblocks.{i}.0.weight, deblocks.{i}.1.*)

In eval mode (the teacher is frozen and forced to eval(), bevdet_distill.py:1591-1597) every
conv + BatchNorm2d + ReLU triple runs as ONE implicit-GEMM kernel on NHWC activations (BN folded
into the epilogue); SECONDFPN's three branches write straight into their channel slice of the
concatenated output. In training mode the plain torch modules run (batch statistics are needed).
"""
import numpy as np
import torch
from torch import nn

from .. import _lib


def _fold_bn2d(bn):
    ts = [bn.weight, bn.bias, bn.running_mean, bn.running_var]
    key = tuple((t.data_ptr(), t._version) for t in ts)
    cache = getattr(bn, "_dbev_folded", None)
    if cache is not None and cache[0] == key:
        return cache[1]
    with torch.no_grad():
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        shift = bn.bias - bn.running_mean * scale
        folded = (scale.float().contiguous(), shift.float().contiguous())
    bn._dbev_folded = (key, folded)
    return folded


def _packed(mod, fn):
    key = (mod.weight.data_ptr(), mod.weight._version)
    cache = getattr(mod, "_dbev_packed", None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            cache = (key, fn(mod.weight.detach().float()))
        mod._dbev_packed = cache
    return cache[1]


def conv_nhwc(x, w_packed, c_out, kh, kw, stride, pad, scale=None, shift=None, relu=False, out=None,
              c_off=0, out_mul=1, out_add=(0, 0), out_nchw=False, out_groups=1):
    """x [N, H, W, C_in] contiguous fp32 -> out [N, Ho*out_mul, Wo*out_mul, ld] (channel slice c_off),
    or, with out_nchw, out [N, ld, Ho*out_mul, Wo*out_mul]. out_groups > 1: the c_out columns are out_groups
    blocks of c_out / out_groups channels, block g lands at lattice x + g (x taps of a transposed conv)."""
    lib = _lib.load()
    _lib.require_cuda(x, "x", torch.float32)
    if not x.is_contiguous():
        raise RuntimeError("x must be a contiguous NHWC tensor")
    n, h, w, c_in = x.shape
    ho = (h + 2 * pad - kh) // stride + 1
    wo = (w + 2 * pad - kw) // stride + 1
    if out is None:
        out = torch.empty((n, ho * out_mul, wo * out_mul, c_out), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        oh, ow, ld = (out.shape[2], out.shape[3], out.shape[1]) if out_nchw else \
            (out.shape[1], out.shape[2], out.shape[3])
        rc = lib.dbev_conv2d_tc_forward_grouped(_lib.ptr(x), n, h, w, c_in, _lib.ptr(w_packed), c_out, kh, kw,
                                                stride, pad, _lib.ptr(scale), _lib.ptr(shift), 1 if relu else 0,
                                                _lib.ptr(out), oh, ow, ld, c_off, out_mul, out_add[0], out_add[1],
                                                1 if out_nchw else 0, out_groups, _lib.stream_ptr(x.device))
    _lib.check(rc, "dbev_conv2d_tc_forward")
    return out


def _to_nhwc(x):
    """NCHW-shaped tensor (any memory format) -> contiguous [N, H, W, C] (free for channels_last)."""
    return x.permute(0, 2, 3, 1).contiguous()


class SECOND(nn.Module):
    """second.py:12-93."""

    def __init__(self, in_channels=128, out_channels=[128, 128, 256], layer_nums=[3, 5, 5],
                 layer_strides=[2, 2, 2], norm_cfg=dict(type="BN", eps=1e-3, momentum=0.01),
                 conv_cfg=dict(type="Conv2d", bias=False), init_cfg=None, pretrained=None,
                 act_cfg=dict(type="ReLU", inplace=True)):
        super(SECOND, self).__init__()
        assert len(layer_strides) == len(layer_nums) == len(out_channels)
        eps, mom = norm_cfg.get("eps", 1e-5), norm_cfg.get("momentum", 0.1)
        in_filters = [in_channels, *out_channels[:-1]]
        blocks = []
        for i, layer_num in enumerate(layer_nums):
            block = [nn.Conv2d(in_filters[i], out_channels[i], 3, stride=layer_strides[i], padding=1, bias=False),
                     nn.BatchNorm2d(out_channels[i], eps=eps, momentum=mom), nn.ReLU(inplace=True)]
            for _ in range(layer_num):
                block += [nn.Conv2d(out_channels[i], out_channels[i], 3, padding=1, bias=False),
                          nn.BatchNorm2d(out_channels[i], eps=eps, momentum=mom), nn.ReLU(inplace=True)]
            blocks.append(nn.Sequential(*block))
        self.blocks = nn.ModuleList(blocks)

    def _fast_ok(self, x):
        return (not self.training) and x.is_cuda and x.dtype == torch.float32 and not (
            torch.is_grad_enabled() and x.requires_grad)

    def forward(self, x):
        if not self._fast_ok(x):
            outs = []
            for b in self.blocks:
                x = b(x)
                outs.append(x)
            return tuple(outs)
        h = _to_nhwc(x)
        outs = []
        for block in self.blocks:
            mods = list(block)
            for j in range(0, len(mods), 3):
                conv, bn = mods[j], mods[j + 1]
                wp = _packed(conv, lambda w: w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous())
                scale, shift = _fold_bn2d(bn)
                h = conv_nhwc(h, wp, conv.out_channels, 3, 3, conv.stride[0], 1, scale, shift, relu=True)
            outs.append(h.permute(0, 3, 1, 2))      # NCHW view of NHWC memory (channels_last)
        return tuple(outs)


class SECONDFPN(nn.Module):
    """second_fpn.py:12-93."""

    def __init__(self, in_channels=[128, 128, 256], out_channels=[256, 256, 256], upsample_strides=[1, 2, 4],
                 norm_cfg=dict(type="BN", eps=1e-3, momentum=0.01), upsample_cfg=dict(type="deconv", bias=False),
                 conv_cfg=dict(type="Conv2d", bias=False), use_conv_for_no_stride=False, init_cfg=None,
                 act_cfg=dict(type="ReLU", inplace=True)):
        super(SECONDFPN, self).__init__()
        assert len(out_channels) == len(upsample_strides) == len(in_channels)
        self.in_channels, self.out_channels = in_channels, out_channels
        eps, mom = norm_cfg.get("eps", 1e-5), norm_cfg.get("momentum", 0.1)
        deblocks = []
        for i, out_channel in enumerate(out_channels):
            stride = upsample_strides[i]
            if stride > 1 or (stride == 1 and not use_conv_for_no_stride):
                up = nn.ConvTranspose2d(in_channels[i], out_channel, int(stride), stride=int(stride), bias=False)
            else:
                k = int(np.round(1 / stride))
                up = nn.Conv2d(in_channels[i], out_channel, k, stride=k, bias=False)
            deblocks.append(nn.Sequential(up, nn.BatchNorm2d(out_channel, eps=eps, momentum=mom),
                                          nn.ReLU(inplace=True)))
        self.deblocks = nn.ModuleList(deblocks)

    def forward(self, x):
        assert len(x) == len(self.in_channels)
        fast = (not self.training) and all(t.is_cuda and t.dtype == torch.float32 for t in x) and not (
            torch.is_grad_enabled() and any(t.requires_grad for t in x)) and all(
            isinstance(d[0], nn.Conv2d) or d[0].kernel_size[0] in (1, 2) for d in self.deblocks)
        if not fast:
            ups = [deblock(x[i]) for i, deblock in enumerate(self.deblocks)]
            return [torch.cat(ups, dim=1) if len(ups) > 1 else ups[0]]
        ld = sum(self.out_channels)
        out, c_off = None, 0
        for i, deblock in enumerate(self.deblocks):
            up, bn = deblock[0], deblock[1]
            h = _to_nhwc(x[i])
            n, hh, ww, _ = h.shape
            scale, shift = _fold_bn2d(bn)
            co = self.out_channels[i]
            if isinstance(up, nn.Conv2d):
                k = up.kernel_size[0]
                ho, wo, mul = hh // k, ww // k, 1
            else:
                k = up.kernel_size[0]
                ho, wo, mul = hh, ww, k
            if out is None:   # NCHW-contiguous concat buffer: what the distillation-loss kernels read
                out = torch.empty((n, ld, ho * mul, wo * mul), dtype=torch.float32, device=h.device)
            if isinstance(up, nn.Conv2d):
                wp = _packed(up, lambda w: w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous())
                conv_nhwc(h, wp, co, k, k, k, 0, scale, shift, True, out=out, c_off=c_off, out_nchw=True)
            else:
                # ConvTranspose2d(k, stride k): out[2y+dy, 2x+dx] = W[:, :, dy, dx]^T . in[y, x] -> k*k 1x1 convs;
                # the k x taps of a row share the input tile: one launch with k column groups when it fits N <= 256
                if k * co <= 256 and co % 32 == 0:
                    wps = _packed(up, lambda w: [torch.cat([w[:, :, dy, dx].t() for dx in range(k)], 0).contiguous()
                                                 for dy in range(k)])
                    for dy in range(k):
                        conv_nhwc(h, wps[dy], k * co, 1, 1, 1, 0, scale, shift, True, out=out, c_off=c_off,
                                  out_mul=k, out_add=(dy, 0), out_nchw=True, out_groups=k)
                else:
                    wps = _packed(up, lambda w: [w[:, :, dy, dx].t().contiguous() for dy in range(k) for dx in range(k)])
                    for dy in range(k):
                        for dx in range(k):
                            conv_nhwc(h, wps[dy * k + dx], co, 1, 1, 1, 0, scale, shift, True, out=out, c_off=c_off,
                                      out_mul=k, out_add=(dy, dx), out_nchw=True)
            c_off += co
        return [out]
