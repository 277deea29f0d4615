"""Student adaptation layers on the B200 kernels.

``Conv1x1Adaptation`` mirrors the '1x1conv' adaptation of BEVDetDistill
(mmdet3d/models/detectors/bevdet_distill.py:216-351: ``nn.Conv2d(student_channel,
teacher_channel, kernel_size=1)``; same parameter names ``weight`` [Ct, Cs, 1, 1] / ``bias``).
Forward runs the tcgen05 TF32 GEMM of csrc/adapt_gemm.cu; the backward (input and weight gradients) runs on the
tcgen05 training conv kernels (csrc/conv2d_tc.cu, conv_wgrad_tc.cu) when both channel counts are multiples of 128
(the shipped 256 -> 384 head position), through aten convolution_backward otherwise.
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


def conv1x1_backward(x, weight, gy, need_x=True, need_w=True):
    """(d x, d weight) of y = conv1x1(x, weight) from gy [B, Cout, H, W] (NCHW memory, what the loss backward kernel
    writes). Cin, Cout multiples of 128: gy is transposed once to channels-last rows and both gradients run on the
    tcgen05 training kernels (input gradient: the per-tap conv kernel with the transposed filter; weight gradient:
    conv_wgrad_tc's MN-major GEMM over the pixels); d x comes back in channels_last memory, the layout the
    student BEV encoder's backward reads. Other shapes: aten convolution_backward (cuDNN)."""
    B, Cin, H, W = x.shape
    Cout = weight.shape[0]
    if Cin % 128 == 0 and Cout % 128 == 0 and gy.is_cuda and gy.dtype == torch.float32:
        from ..ops import conv_train as ct
        g_cl = transpose_batched(gy.contiguous(), B, Cout, H * W).view(B, H, W, Cout)
        gx = gw = None
        if need_x:
            gx = ct.as_nchw(ct.conv_input_grad(g_cl, ct.pack_weights(weight, 1), Cin, 1, 1, 1, 0, (H, W)))
        if need_w:
            gw = ct.conv_weight_grad(ct.as_nhwc(x), g_cl, 1, 1, 1, 0)
        return gx, gw
    gy = gy.contiguous()
    gx = torch.nn.grad.conv2d_input(x.shape, weight, gy) if need_x else None
    gw = torch.nn.grad.conv2d_weight(x, weight.shape, gy) if need_w else None
    return gx, gw


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
        gx, gw = conv1x1_backward(x, weight, gy, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
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


# ------------------------------------------------------------------------------------------------------
# The other adaptation layers BEVDetDistill.__init__ can build (bevdet_distill.py:216-345): same classes,
# constructor arguments, sub-module names and state_dict keys; the forward of every conv + BatchNorm2d
# (training: batch statistics) + ReLU group is one autograd node on the tcgen05 conv kernels (forward,
# input gradient, weight gradient) through plugin/bev_encoder.conv_bn_act, the bilinear upsampling the NHWC
# kernels of csrc/bev_encoder_ops.cu. The shipped recipe (scripts/teacher_to_bevdepth4d/centerpoint2bevdepth.sh:32-37)
# uses 'upsample_3layer' at the two backbone positions and '1x1conv' at the head position.
# ------------------------------------------------------------------------------------------------------

def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def _bn_layer(norm_layer, channels):
    if isinstance(norm_layer, dict):
        cfg = dict(norm_layer)
        if cfg.pop("type") not in ("BN", "BN2d"):
            raise NotImplementedError("adaptation norm_layer %r: only BatchNorm2d is implemented" % (norm_layer,))
        cfg.pop("requires_grad", None)
        return nn.BatchNorm2d(channels, **cfg)
    layer = norm_layer(channels)
    assert isinstance(layer, nn.BatchNorm2d)
    return layer


def _relu_only(act_layer):
    if act_layer is not None and not isinstance(act_layer(), nn.ReLU):
        raise NotImplementedError("adaptation act_layer: only nn.ReLU is implemented")


def _conv_group(x, conv, bn, relu):
    from ..bev_encoder import conv_bn_act, _ConvBNActFn
    k, s = conv.kernel_size[0], conv.stride[0]
    pad = conv.padding[0]
    if (k == 1 and s == 1 and pad == 0) or (k == 3 and s in (1, 2) and pad == 1):
        return conv_bn_act(x, conv, bn, relu=relu)
    if (k == s and k > 1 and pad == 0 and conv.kernel_size[1] == k and conv.stride[1] == k and conv.groups == 1
            and conv.dilation == (1, 1) and x.shape[2] % k == 0 and x.shape[3] % k == 0):
        # 'downsample_2layer' (bevdet_distill.py:252-257): a k x k / stride k / pad 0 conv reads disjoint patches, i.e. it is
        # a 1x1 conv over the space-to-depth rearrangement of x with the filter flattened (ky, kx, ci) -> K. Both
        # rearrangements are plain tensor views / one copy that autograd differentiates; the GEMM, its input / weight
        # gradients and the BatchNorm run on the same tcgen05 training kernels as every other adaptation conv.
        n, c, h, w = x.shape
        x2 = x.reshape(n, c, h // k, k, w // k, k).permute(0, 3, 5, 1, 2, 4).reshape(n, k * k * c, h // k, w // k)
        w2 = conv.weight.permute(0, 2, 3, 1).reshape(conv.out_channels, k * k * c, 1, 1)
        gamma = bn.weight if bn is not None else None
        beta = bn.bias if bn is not None else None
        return _ConvBNActFn.apply(x2, w2, conv.bias, gamma, beta, None, bn, conv, 1, 0, relu)
    raise NotImplementedError("adaptation conv %dx%d / stride %d / pad %d: the tcgen05 training kernels cover 1x1 / stride 1, "
                              "3x3 / pad 1 / stride 1-2 and k x k / stride k / pad 0 (patch) convolutions" % (k, k, s, pad))


class Mlp(nn.Module):
    """bevdet_distill.py:48-68 ('mlp'): fc1 (1x1 conv) - ReLU - fc2 (1x1 conv); dropout p = 0 in every use."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=None, drop=0.):
        super().__init__()
        _relu_only(act_layer)
        if drop != 0.:
            raise NotImplementedError("Mlp adaptation: drop must be 0 (the reference never sets it)")
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.stride = _pair(1)
        self.fc1 = nn.Conv2d(in_features, hidden_features, kernel_size=1, stride=1, padding=0)
        self.act = nn.ReLU(inplace=True)
        self.fc2 = nn.Conv2d(hidden_features, out_features, kernel_size=1, stride=1, padding=0)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return _conv_group(_conv_group(x, self.fc1, None, True), self.fc2, None, False)


class TwoLayer(nn.Module):
    """bevdet_distill.py:71-96: conv1 - norm1 - ReLU - conv2 (1x1) - norm2 - ReLU."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=None, norm_layer=nn.BatchNorm2d,
                 kernel_size=4, stride=4, padding=0):
        super().__init__()
        _relu_only(act_layer)
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.stride = _pair(stride)
        self.conv1 = nn.Conv2d(in_features, hidden_features, kernel_size=kernel_size, stride=stride, padding=padding)
        self.norm1 = _bn_layer(norm_layer, hidden_features)
        self.act1 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(hidden_features, out_features, kernel_size=1, stride=1, padding=0)
        self.norm2 = _bn_layer(norm_layer, out_features)
        self.act2 = nn.ReLU(inplace=True)

    def forward(self, x):
        return _conv_group(_conv_group(x, self.conv1, self.norm1, True), self.conv2, self.norm2, True)


class ThreeLayer(nn.Module):
    """bevdet_distill.py:99-130: TwoLayer with one more 1x1 conv - norm - ReLU in the middle."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=None, norm_layer=nn.BatchNorm2d,
                 kernel_size=4, stride=4, padding=0):
        super().__init__()
        _relu_only(act_layer)
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.stride = _pair(stride)
        self.conv1 = nn.Conv2d(in_features, hidden_features, kernel_size=kernel_size, stride=stride, padding=padding)
        self.norm1 = _bn_layer(norm_layer, hidden_features)
        self.act1 = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(hidden_features, hidden_features, kernel_size=1, stride=1, padding=0)
        self.norm2 = _bn_layer(norm_layer, hidden_features)
        self.act2 = nn.ReLU(inplace=True)
        self.conv3 = nn.Conv2d(hidden_features, out_features, kernel_size=1, stride=1, padding=0)
        self.norm3 = _bn_layer(norm_layer, out_features)
        self.act3 = nn.ReLU(inplace=True)

    def forward(self, x):
        x = _conv_group(x, self.conv1, self.norm1, True)
        x = _conv_group(x, self.conv2, self.norm2, True)
        return _conv_group(x, self.conv3, self.norm3, True)


class BilinearUpsample(nn.Upsample):
    """nn.Upsample(scale_factor, mode='bilinear', align_corners=True) on the NHWC kernel (no parameters)."""

    def __init__(self, scale_factor):
        super().__init__(scale_factor=scale_factor, mode="bilinear", align_corners=True)

    def forward(self, x):
        from ..bev_encoder import upsample_cat
        return upsample_cat(x, None, self.scale_factor)


class Conv3x3Adaptation(nn.Conv2d):
    """'3x3conv' (:233-235): nn.Conv2d(Cs, Ct, 3, 1, 1) with bias."""

    def __init__(self, in_channels, out_channels):
        super().__init__(in_channels, out_channels, kernel_size=3, stride=1, padding=1)

    def forward(self, x):
        return _conv_group(x, self, None, False)


def build_adaptation_layers(distill_params, student_channels=None, teacher_channels=None):
    """The 'fgd' branch of BEVDetDistill.__init__ (bevdet_distill.py:216-351): returns
    (channel_wise_adaptations, teacher_adaptations, spatial_wise_adaptations or None) as nn.ModuleLists with the
    reference's indices, ``.stride`` attributes and state_dict keys."""
    p = distill_params
    student_channels = student_channels if student_channels is not None else p["student_channels"]
    teacher_channels = teacher_channels if teacher_channels is not None else p["teacher_channels"]
    assert len(student_channels) == len(teacher_channels)
    n = len(student_channels)
    atypes = p["adaptation_type"]
    atypes = [atypes] * n if isinstance(atypes, str) else list(atypes)
    ttypes = p.get("teacher_adaptation_type", "identity")
    ttypes = [ttypes] * n if isinstance(ttypes, str) else list(ttypes)
    sp = p.get("student_adaptation_params", {})
    tp = p.get("teacher_adaptation_params", {})
    student, teacher = [], []
    for index, (atype, ttype, cs, ct) in enumerate(zip(atypes, ttypes, student_channels, teacher_channels)):
        if atype == "1x1conv":
            m = Conv1x1Adaptation(cs, ct)
        elif atype == "3x3conv":
            m = Conv3x3Adaptation(cs, ct)
        elif atype == "mlp":
            m = Mlp(in_features=cs, out_features=ct)
        elif atype in ("2layer", "3layer"):
            assert sp["kernel_size"] == 1 and sp["stride"] == 1      # :240-241, :246-247
            m = (TwoLayer if atype == "2layer" else ThreeLayer)(in_features=cs, out_features=ct,
                                                                kernel_size=sp["kernel_size"], stride=sp["stride"])
        elif atype == "downsample_2layer":
            m = TwoLayer(in_features=cs, out_features=ct, kernel_size=sp["downsample_kernel_size"], stride=sp["downsample_stride"])
        elif atype == "identity":
            m = nn.Identity()
            m.stride = _pair(1)
        elif atype in ("upsample_2layer", "upsample_3layer", "upsample_1x1conv"):
            assert sp["upsample_factor"] % sp["stride"] == 0 and sp["stride"] == 1      # :267-268
            inner = (Conv1x1Adaptation(cs, ct) if atype == "upsample_1x1conv" else
                     (TwoLayer if atype == "upsample_2layer" else ThreeLayer)(in_features=cs, out_features=ct,
                                                                             kernel_size=sp["kernel_size"], stride=sp["stride"]))
            m = nn.Sequential(BilinearUpsample(sp["upsample_factor"]), inner)
            m.stride = _pair(sp["stride"] / sp["upsample_factor"])
        elif atype == "avgpool_1x1conv":
            m = nn.Sequential(nn.AvgPool2d(kernel_size=sp["downsample_kernel_size"]), Conv1x1Adaptation(cs, ct))
            m.stride = _pair(sp["downsample_kernel_size"])
        else:
            raise NotImplementedError("adaptation_type=%r" % (atype,))
        student.append(m)
        # teacher side (:328-345): parameter-free pooling of a frozen feature map (torch's pooling kernels, no autograd)
        if ttype == "avgpool":
            t = nn.AvgPool2d(**tp)
            t.stride = _pair(t.stride)
        elif ttype == "maxpool":
            t = nn.MaxPool2d(**tp)
            t.stride = _pair(t.stride)
        elif ttype == "identity":
            t = nn.Identity()
            t.stride = _pair(1)
        elif ttype == "downsample_3layer":
            t = ThreeLayer(in_features=ct, out_features=cs, kernel_size=tp["kernel_size"], stride=tp["stride"])
        else:
            raise NotImplementedError("teacher_adaptation_type=%r" % (ttype,))
        teacher.append(t)
    spatial = None
    if p.get("spatial_mask"):
        spatial = nn.ModuleList([nn.Conv2d(1, 1, kernel_size=3, stride=1, padding=1) for _ in student_channels])
    return nn.ModuleList(student), nn.ModuleList(teacher), spatial
