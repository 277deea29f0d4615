"""The student's BEV encoder, TRAINED on the tcgen05 conv kernels. Same registry names,
constructor arguments, sub-module layout and ``state_dict`` keys as the reference:

  BasicBlock        mmdet3d/models/bricks/res_block.py:10-99      conv1 / bn1 / conv2 / bn2 / downsample
  ResNetForBEVDet   mmdet3d/models/backbones/resnet.py:12-62      layers.{i}.{j}.*
  FPN_LSS           mmdet3d/models/necks/lss_fpn.py:10-72         conv.{0,1,3,4}.*, up2.{1,2,4}.*, lateral_conv.*

The reference runs these as nn.Conv2d + nn.BatchNorm2d (training mode: batch statistics) + nn.ReLU + nn.Upsample
through cuDNN / ATen. Here every conv + BN + (residual) + ReLU group is one autograd node on NHWC fp32 memory:

  forward   pack W -> conv (tcgen05 TF32, TMA halo tiles) -> per-channel batch statistics -> a*y + b (+ identity) -> ReLU
  backward  g = dz * (z > 0); BN backward (two passes over y) -> dy; bias / gamma / beta gradients from the same sums;
            dx = conv of dy with the flipped filter (same tcgen05 kernel; stride 2 = four parity-class convs);
            dW = tcgen05 MN-major GEMM over the pixels (conv_wgrad_tc.cu)

TF32 multiply / fp32 accumulate - what the reference's cuDNN path does under torch's default
``torch.backends.cudnn.allow_tf32 = True``. Tensors between the nodes are NCHW-shaped views of NHWC memory
(``torch.channels_last``), so the modules compose with ordinary torch code. CUDA fp32 only: anything else raises
(no CPU fallback).
"""
import torch
from torch import nn

from .ops import conv_train as ct


def _check_input(x):
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32):
        raise RuntimeError("distill_bev_b200 BEV encoder: CUDA fp32 tensors only (got %s %s); there is no CPU path"
                           % (getattr(x, "device", None), getattr(x, "dtype", None)))


class _ConvBNActFn(torch.autograd.Function):
    """z = relu?(BN_train(conv(x, W) + bias) + residual). BN optional (bn=None), residual optional."""

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, residual, bn, owner, stride, pad, relu):
        xh = ct.as_nhwc(x)
        co, ci, kh, kw = weight.shape
        pre = getattr(owner, "_dbev_prepacked", None)
        if pre is not None and pre[0] == (weight.data_ptr(), weight._version):
            w_fwd, w_bwd, ready = pre[1]                           # packed ahead of time on the side stream (prepack())
            if ready is not None:
                torch.cuda.current_stream(x.device).wait_event(ready)
        else:
            w_fwd, w_bwd = ct.pack_weights_train(weight, stride)   # the backward reads w_bwd (same step, same weights)
        y = ct.conv_forward(xh, w_fwd, co, kh, kw, stride, pad, bias=bias.detach() if bias is not None else None)
        rh = ct.as_nhwc(residual) if residual is not None else None
        fwd = None
        if bn is not None:
            if bn.training or bn.running_mean is None:
                track = bn.training and bn.track_running_stats and bn.running_mean is not None
                if track and bn.momentum is None:
                    raise NotImplementedError("BatchNorm2d(momentum=None) (cumulative average) is not implemented")
                fwd = ct.bn_batch_stats(y, gamma, beta, bn.eps, bn.momentum if bn.momentum is not None else 0.1,
                                        bn.running_mean if track else None, bn.running_var if track else None,
                                        ws=_stats_ws(bn, y))
            else:   # eval: running statistics (same kernels; mean / invstd rows are what the backward would need)
                invstd = torch.rsqrt(bn.running_var + bn.eps)
                g = gamma if gamma is not None else torch.ones_like(invstd)
                b = beta if beta is not None else torch.zeros_like(invstd)
                a = g * invstd
                fwd = torch.stack([a, b - bn.running_mean * a, bn.running_mean, invstd]).contiguous()
            z, mask = ct.bn_act(y, fwd, rh, relu, want_mask=True)
        elif rh is not None or relu:
            z, mask = ct.bn_act(y, None, rh, relu, want_mask=True)
        else:
            z, mask = y, None
        ctx.conf = (stride, pad, relu, bn is not None, bn is not None and (bn.training or bn.running_mean is None), bn,
                    tuple(xh.shape[1:3]), residual is not None, bias is not None, owner)
        ctx.save_for_backward(xh, weight, y, mask, fwd, gamma, w_bwd)    # mask: the ReLU decision, 1 byte per 4 channels
        return ct.as_nchw(z)

    @staticmethod
    def backward(ctx, dz):
        xh, weight, y, mask, fwd, gamma, w_bwd = ctx.saved_tensors
        stride, pad, relu, has_bn, batch_stats, bn, in_hw, has_res, has_bias, owner = ctx.conf
        co, ci, kh, kw = weight.shape
        dzh = ct.as_nhwc(dz)
        need = ctx.needs_input_grad
        d_gamma = d_beta = d_bias = d_res = None
        if has_bn and batch_stats:
            dy, bwd, g = ct.bn_backward(dzh, None, y, fwd, want_g=has_res and need[5], ws=_stats_ws(bn, y), mask=mask)
            d_gamma, d_beta = (bwd[0] if need[3] else None), (bwd[1] if need[4] else None)
            d_res = g
        else:
            g = ct.relu_backward(dzh, mask=mask) if relu else dzh
            d_res = g if has_res else None
            if has_bn:     # eval-mode BN: a fixed per-channel affine
                dy = ct.bn_act(g, torch.stack([fwd[0], torch.zeros_like(fwd[0])]).contiguous(), None, False)
                if need[3]:
                    d_gamma = ((g * ((y - fwd[2]) * fwd[3])).sum((0, 1, 2)))
                if need[4]:
                    d_beta = g.sum((0, 1, 2))
            else:
                dy = g
        if has_bias and need[2]:
            d_bias = ct.channel_sums(dy, ws=_stats_ws(owner, dy))
        dx = None
        if need[0]:
            dx = ct.as_nchw(ct.conv_input_grad(dy, w_bwd, ci, kh, kw, stride, pad, in_hw))
        dw = None
        if need[1]:
            side = ct.side_stream(dy.device)
            if side is None:
                dw = ct.conv_weight_grad(xh, dy, kh, kw, stride, pad)
            else:
                # nothing in the backward chain reads dW: run it beside the input-gradient chain (joined by
                # ct.join_side_stream() before the optimizer / all-reduce)
                cur = torch.cuda.current_stream(dy.device)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    if ct._side["test_delay_cycles"]:          # tests: hold the side stream back to expose any consumer racing ahead
                        torch.cuda._sleep(ct._side["test_delay_cycles"])
                    dw = ct.conv_weight_grad(xh, dy, kh, kw, stride, pad)
                for t in (xh, dy, dw):
                    t.record_stream(side)
                dw.record_stream(cur)
        if d_res is not None:
            d_res = ct.as_nchw(d_res) if need[5] else None
        return dx, dw, d_bias, d_gamma, d_beta, d_res, None, None, None, None, None


def _stats_ws(owner, y):
    """Per-module reduction workspace (per-block partial sums)."""
    rows, c = y.shape[0] * y.shape[1] * y.shape[2], y.shape[3]
    key = (rows, c, y.device)
    cache = getattr(owner, "_dbev_stats_ws", None)
    if cache is None or cache[0] != key:
        cache = (key, ct.stats_workspace(rows, c, y.device))
        owner._dbev_stats_ws = cache
    return cache[1]


def _pack_plan(module, convs, dev):
    """Persistent packed-weight buffers of every conv of ``module`` + the device job table of the one-launch pack.
    Built on first use (a host -> device copy: outside CUDA graph capture), rebuilt when a weight tensor was replaced."""
    key = tuple((m.weight.data_ptr(), tuple(m.weight.shape), m.stride[0]) for m in convs)
    plan = getattr(module, "_dbev_pack_plan", None)
    if plan is not None and plan["key"] == key:
        return plan
    if torch.cuda.is_current_stream_capturing():
        return None
    total = sum(m.weight.numel() for m in convs)
    buf = torch.empty((2, total), dtype=torch.float32, device=dev)
    rows, views, off, tile = [], [], 0, 0
    for m in convs:
        co, ci, kh, kw = m.weight.shape
        n = m.weight.numel()
        w_fwd, w_bwd = buf[0, off:off + n], buf[1, off:off + n]
        rows.append([m.weight.data_ptr(), w_fwd.data_ptr(), w_bwd.data_ptr(), co, ci, kh * 256 + kw,
                     1 if m.stride[0] == 1 else 2, tile])
        views.append((w_fwd, w_bwd))
        off += n
        tile += ((co + 31) // 32) * ((ci + 31) // 32)
    plan = {"key": key, "buf": buf, "views": views, "tiles": tile,
            "jobs": torch.tensor(rows, dtype=torch.int64).to(dev)}
    module._dbev_pack_plan = plan
    return plan


def prepack(module):
    """Pack the weights of every conv of ``module`` that conv_bn_act will run, now, on the side stream (or the current
    stream when overlap is off), off the critical path of the step that follows: ONE launch over all layers into
    persistent buffers (per-layer launches when the job table cannot be built, i.e. first use inside a graph capture).
    Call it once per step after the optimizer has updated the weights; a conv whose weight changed since falls back to
    packing in its forward."""
    convs = [m for m in module.modules() if isinstance(m, nn.Conv2d) and m.weight.is_cuda and m.groups == 1
             and m.kernel_size[0] == m.kernel_size[1] and m.kernel_size[0] in (1, 3)]
    if not convs:
        return
    dev = convs[0].weight.device
    batched = all(m.weight.is_contiguous() and m.weight.dtype == torch.float32 and m.weight.device == dev
                  and m.stride[0] == m.stride[1] and (m.stride[0] == 1 or (m.stride[0] == 2 and m.kernel_size[0] == 3))
                  for m in convs)
    plan = _pack_plan(module, convs, dev) if batched else None
    side = ct.side_stream(dev)
    cur = torch.cuda.current_stream(dev)
    if side is not None:
        side.wait_stream(cur)          # also orders the overwrite of the persistent buffers after their last readers
    with torch.cuda.stream(side if side is not None else cur):
        if plan is not None:
            ct.pack_weights_batch(plan["jobs"], plan["tiles"])
            ready = None
            if side is not None:
                ready = torch.cuda.Event()
                ready.record(side)
            for m, (w_fwd, w_bwd) in zip(convs, plan["views"]):
                m._dbev_prepacked = ((m.weight.data_ptr(), m.weight._version), (w_fwd, w_bwd, ready))
            return
        for m in convs:
            w_fwd, w_bwd = ct.pack_weights_train(m.weight, m.stride[0])
            ready = None
            if side is not None:
                ready = torch.cuda.Event()
                ready.record(side)
                w_fwd.record_stream(cur), w_bwd.record_stream(cur)
            m._dbev_prepacked = ((m.weight.data_ptr(), m.weight._version), (w_fwd, w_bwd, ready))


def conv_bn_act(x, conv, bn=None, residual=None, relu=True):
    """conv (nn.Conv2d) -> bn (nn.BatchNorm2d or None) -> (+ residual) -> ReLU? as one autograd node."""
    _check_input(x)
    if conv.groups != 1 or conv.dilation != (1, 1) or conv.padding_mode != "zeros" or conv.stride[0] != conv.stride[1] \
            or conv.padding[0] != conv.padding[1] or conv.kernel_size[0] != conv.kernel_size[1]:
        raise NotImplementedError("conv_bn_act: square filters, groups = dilation = 1, zero padding only")
    gamma = bn.weight if bn is not None else None
    beta = bn.bias if bn is not None else None
    return _ConvBNActFn.apply(x, conv.weight, conv.bias, gamma, beta, residual, bn, conv, conv.stride[0], conv.padding[0], relu)


class _UpsampleFn(torch.autograd.Function):
    """out = cat([skip, upsample_bilinear(x, scale, align_corners=True)], channel) (skip optional), NHWC memory."""

    @staticmethod
    def forward(ctx, x, skip, scale):
        xh = ct.as_nhwc(x)
        n, h, w, c = xh.shape
        big_h, big_w = int(h * scale), int(w * scale)
        cs = skip.shape[1] if skip is not None else 0
        out = torch.empty((n, big_h, big_w, cs + c), dtype=torch.float32, device=x.device)
        if skip is not None:
            ct.bn_act(ct.as_nhwc(skip), None, None, False, out=out[..., :cs])       # copy into the concat's channel slice
        ct.upsample_bilinear(xh, scale, out=out[..., cs:])
        ctx.conf = (h, w, cs)
        return ct.as_nchw(out)

    @staticmethod
    def backward(ctx, dout):
        h, w, cs = ctx.conf
        dh = ct.as_nhwc(dout)
        dx = ct.as_nchw(ct.upsample_bilinear_backward(dh[..., cs:], (h, w))) if ctx.needs_input_grad[0] else None
        dskip = ct.as_nchw(dh[..., :cs]) if (cs and ctx.needs_input_grad[1]) else None
        return dx, dskip, None


def upsample_cat(x, skip, scale):
    _check_input(x)
    return _UpsampleFn.apply(x, skip, scale)


def _bn(norm_cfg, channels):
    cfg = dict(norm_cfg or dict(type="BN"))
    kind = cfg.pop("type", "BN")
    if kind not in ("BN", "BN2d"):
        raise NotImplementedError("norm_cfg type %r: only plain BatchNorm2d (the shipped configs) is implemented" % kind)
    cfg.pop("requires_grad", None)
    return nn.BatchNorm2d(channels, **cfg)


def _check_act(act_cfg):
    if dict(act_cfg or {}).get("type", "ReLU") != "ReLU":
        raise NotImplementedError("act_cfg %r: only ReLU (the shipped configs) is implemented" % (act_cfg,))


class BasicBlock(nn.Module):
    """bricks/res_block.py:10-99 (sub-module names conv1 / bn1 / conv2 / bn2 / downsample as in mmdet's BasicBlock)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, style="pytorch", with_cp=False,
                 conv_cfg=None, norm_cfg=dict(type="BN"), dcn=None, plugins=None, init_cfg=None,
                 act_cfg=dict(type="ReLU", inplace=True)):
        super(BasicBlock, self).__init__()
        assert dcn is None and plugins is None, "Not implemented yet."
        if dilation != 1 or conv_cfg is not None:
            raise NotImplementedError("BasicBlock: dilation 1 and plain Conv2d only")
        _check_act(act_cfg)
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn1 = _bn(norm_cfg, planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = _bn(norm_cfg, planes)
        self.downsample = downsample
        self.stride, self.dilation, self.with_cp = stride, dilation, with_cp

    @property
    def norm1(self):
        return self.bn1

    @property
    def norm2(self):
        return self.bn2

    def forward(self, x):
        out = conv_bn_act(x, self.conv1, self.bn1, relu=True)
        identity = x
        if self.downsample is not None:
            if not isinstance(self.downsample, nn.Conv2d):
                raise NotImplementedError("BasicBlock: downsample must be the nn.Conv2d ResNetForBEVDet builds")
            identity = conv_bn_act(x, self.downsample, None, relu=False)
        return conv_bn_act(out, self.conv2, self.bn2, residual=identity, relu=True)


class Bottleneck(nn.Module):
    """bricks/res_block.py:102-311 (mmdet's Bottleneck with act_cfg; sub-module names conv1 / bn1 / conv2 / bn2 / conv3 /
    bn3 / downsample): 1x1 -> 3x3 (the strided one, style 'pytorch') -> 1x1 (x4 channels) + identity, ReLU. Every
    conv + BatchNorm (+ residual) + ReLU group is one autograd node on the tcgen05 training kernels; `planes` must be a
    multiple of 128 (the weight-gradient kernel's tile), e.g. num_channels = [512, 1024, 2048]."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, style="pytorch", with_cp=False,
                 conv_cfg=None, norm_cfg=dict(type="BN"), dcn=None, plugins=None, init_cfg=None,
                 act_cfg=dict(type="ReLU", inplace=True)):
        super(Bottleneck, self).__init__()
        assert style in ("pytorch", "caffe")
        assert dcn is None and plugins is None, "Not implemented yet."
        if dilation != 1 or conv_cfg is not None:
            raise NotImplementedError("Bottleneck: dilation 1 and plain Conv2d only")
        if style == "caffe" and stride != 1:
            raise NotImplementedError("Bottleneck style 'caffe' with stride > 1 (a strided 1x1 conv) is not implemented")
        _check_act(act_cfg)
        self.inplanes, self.planes, self.stride, self.dilation, self.style, self.with_cp = inplanes, planes, stride, dilation, style, with_cp
        self.conv1_stride, self.conv2_stride = (1, stride) if style == "pytorch" else (stride, 1)
        self.conv1 = nn.Conv2d(inplanes, planes, 1, stride=self.conv1_stride, bias=False)
        self.bn1 = _bn(norm_cfg, planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=self.conv2_stride, padding=1, bias=False)
        self.bn2 = _bn(norm_cfg, planes)
        self.conv3 = nn.Conv2d(planes, planes * self.expansion, 1, bias=False)
        self.bn3 = _bn(norm_cfg, planes * self.expansion)
        self.downsample = downsample

    @property
    def norm1(self):
        return self.bn1

    @property
    def norm2(self):
        return self.bn2

    @property
    def norm3(self):
        return self.bn3

    def forward(self, x):
        out = conv_bn_act(x, self.conv1, self.bn1, relu=True)
        out = conv_bn_act(out, self.conv2, self.bn2, relu=True)
        identity = x
        if self.downsample is not None:
            if not isinstance(self.downsample, nn.Conv2d):
                raise NotImplementedError("Bottleneck: downsample must be the nn.Conv2d ResNetForBEVDet builds")
            identity = conv_bn_act(x, self.downsample, None, relu=False)
        return conv_bn_act(out, self.conv3, self.bn3, residual=identity, relu=True)


class ResNetForBEVDet(nn.Module):
    """backbones/resnet.py:12-62 (block_type 'Basic' - the shipped configs - or 'BottleNeck')."""

    def __init__(self, numC_input, num_layer=[2, 2, 2], num_channels=None, stride=[2, 2, 2], backbone_output_ids=None,
                 norm_cfg=dict(type="BN"), act_cfg=dict(type="ReLU", inplace=True), with_cp=False, block_type="Basic"):
        super(ResNetForBEVDet, self).__init__()
        assert len(num_layer) == len(stride)
        assert block_type in ("Basic", "BottleNeck")
        num_channels = [numC_input * 2 ** (i + 1) for i in range(len(num_layer))] if num_channels is None else num_channels
        self.backbone_output_ids = range(len(num_layer)) if backbone_output_ids is None else backbone_output_ids
        layers, curr = [], numC_input
        for i in range(len(num_layer)):
            down = nn.Conv2d(curr, num_channels[i], 3, stride[i], 1)
            if block_type == "BottleNeck":            # resnet.py:26-35
                layer = [Bottleneck(curr, num_channels[i] // 4, stride=stride[i], downsample=down, norm_cfg=norm_cfg, act_cfg=act_cfg)]
                curr = num_channels[i]
                layer.extend([Bottleneck(curr, curr // 4, norm_cfg=norm_cfg, act_cfg=act_cfg) for _ in range(num_layer[i] - 1)])
            else:                                     # resnet.py:36-44
                layer = [BasicBlock(curr, num_channels[i], stride=stride[i], downsample=down, norm_cfg=norm_cfg, act_cfg=act_cfg)]
                curr = num_channels[i]
                layer.extend([BasicBlock(curr, curr, norm_cfg=norm_cfg, act_cfg=act_cfg) for _ in range(num_layer[i] - 1)])
            layers.append(nn.Sequential(*layer))
        self.layers = nn.Sequential(*layers)
        self.with_cp = with_cp

    def forward(self, x):
        _check_input(x)
        feats, x_tmp = [], x
        for lid, layer in enumerate(self.layers):
            x_tmp = layer(x_tmp)
            if lid in self.backbone_output_ids:
                feats.append(x_tmp)
        return feats


class FPN_LSS(nn.Module):
    """necks/lss_fpn.py:10-72."""

    def __init__(self, in_channels, out_channels, scale_factor=4, input_feature_index=(0, 2), norm_cfg=dict(type="BN"),
                 extra_upsample=2, lateral=None, extra_norm_act=False, act_cfg=dict(type="ReLU", inplace=True)):
        super().__init__()
        _check_act(act_cfg)
        self.input_feature_index = input_feature_index
        self.extra_upsample = extra_upsample is not None
        self.up = nn.Upsample(scale_factor=scale_factor, mode="bilinear", align_corners=True)
        f = 2 if self.extra_upsample else 1
        self.conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels * f, kernel_size=3, padding=1, bias=False), _bn(norm_cfg, out_channels * f),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels * f, out_channels * f, kernel_size=3, padding=1, bias=False), _bn(norm_cfg, out_channels * f),
            nn.ReLU(inplace=True))
        if self.extra_upsample:
            up2 = [nn.Upsample(scale_factor=extra_upsample, mode="bilinear", align_corners=True),
                   nn.Conv2d(out_channels * f, out_channels, kernel_size=3, padding=1, bias=False), _bn(norm_cfg, out_channels),
                   nn.ReLU(inplace=True), nn.Conv2d(out_channels, out_channels, kernel_size=1, padding=0)]
            if extra_norm_act:
                up2 += [_bn(norm_cfg, out_channels), nn.ReLU(inplace=True)]
            self.up2 = nn.Sequential(*up2)
        self.extra_norm_act = extra_norm_act
        self.lateral = lateral is not None
        if self.lateral:
            self.lateral_conv = nn.Sequential(nn.Conv2d(lateral, lateral, kernel_size=1, padding=0, bias=False),
                                              _bn(norm_cfg, lateral), nn.ReLU(inplace=True))

    def forward(self, feats):
        x2, x1 = feats[self.input_feature_index[0]], feats[self.input_feature_index[1]]
        _check_input(x1)
        if self.lateral:
            x2 = conv_bn_act(x2, self.lateral_conv[0], self.lateral_conv[1], relu=True)
        x = upsample_cat(x1, x2, self.up.scale_factor)              # cat([x2, up(x1)], dim=1)
        x = conv_bn_act(x, self.conv[0], self.conv[1], relu=True)
        x = conv_bn_act(x, self.conv[3], self.conv[4], relu=True)
        if self.extra_upsample:
            x = upsample_cat(x, None, self.up2[0].scale_factor)
            x = conv_bn_act(x, self.up2[1], self.up2[2], relu=True)
            if self.extra_norm_act:
                x = conv_bn_act(x, self.up2[4], self.up2[5], relu=True)
            else:
                x = conv_bn_act(x, self.up2[4], None, relu=False)
        return x
