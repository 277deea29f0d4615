"""Training primitives of the student's BEV encoder on NHWC fp32 tensors (csrc/conv2d_tc.cu, conv_wgrad_tc.cu,
bev_encoder_ops.cu through the C-ABI): convolution forward / input gradient / weight gradient on tcgen05,
BatchNorm with batch statistics (+ ReLU, + residual) forward / backward, bilinear upsampling forward / backward.

Reference: the reference has no native code here - ResNetForBEVDet
(mmdet3d/models/backbones/resnet.py:51-62), BasicBlock (bricks/res_block.py:70-99) and FPN_LSS
(necks/lss_fpn.py:62-72) are nn.Conv2d / nn.BatchNorm2d / nn.ReLU / nn.Upsample modules, i.e.
aten::cudnn_convolution, convolution_backward, native_batch_norm(_backward), upsample_bilinear2d(_backward).

An "NHWC tensor" below is a torch tensor of shape [N, H, W, C] whose last stride is 1 and whose other strides are
(H*W*ld, W*ld, ld) for some ld >= C: a contiguous tensor or a channel slice of one (so torch.cat never needs a
separate pass). No CPU path: CPU tensors raise.
"""
import os

import torch

from ... import _lib

_TILE_N = (256, 128, 64)

# ---------------------------------------------------------------------------------------------------------------
# Side stream for work that is off the critical path of a training step: weight gradients (nothing in the backward
# chain depends on dW; only the optimizer / gradient all-reduce does) and the per-step weight packing (depends only on
# the weights). The small-map layers of the BEV encoder (16 x 16 ... 64 x 64) do not fill 148 SMs, so these launches
# overlap the input-gradient chain instead of queueing behind it. Off by default (plain single-stream semantics);
# bench.py / a training loop turns it on and must call ``join_side_stream()`` before reading the weight gradients
# (GradientAllReduce does it by itself).
_side = {"enabled": False, "streams": {}, "test_delay_cycles": 0}


def set_side_stream(enabled):
    _side["enabled"] = bool(enabled)


def side_stream(device):
    """The side stream of ``device`` if overlap is enabled, else None."""
    if not _side["enabled"]:
        return None
    device = torch.device(device)
    key = device.index if device.index is not None else torch.cuda.current_device()
    st = _side["streams"].get(key)
    if st is None:
        st = _side["streams"][key] = torch.cuda.Stream(torch.device("cuda", key))
    return st


def join_side_stream(device, stream=None):
    """Make ``stream`` (default: the current stream) wait for everything queued on the side stream."""
    st = side_stream(device)
    if st is not None:
        (stream if stream is not None else torch.cuda.current_stream(st.device)).wait_stream(st)


def nhwc_ld(t, name="tensor"):
    """Row stride (floats) of an NHWC tensor / channel slice; raises if the layout is anything else."""
    _lib.require_cuda(t, name, torch.float32)
    if t.dim() != 4:
        raise RuntimeError("%s must be [N, H, W, C]" % name)
    n, h, w, c = t.shape
    ld = t.stride(2) if w > 1 else (t.stride(1) if h > 1 else max(c, 1))
    ok = t.stride(3) == 1 or c == 1
    ok = ok and ld >= c and (w == 1 or t.stride(2) == ld) and (h == 1 or t.stride(1) == w * ld) and (
        n == 1 or t.stride(0) == h * w * ld)
    if not ok or ld % 4 or t.data_ptr() % 16:
        raise RuntimeError("%s must be an NHWC tensor or a channel slice of one (shape %s, strides %s)"
                           % (name, tuple(t.shape), tuple(t.stride())))
    return ld


def as_nhwc(x):
    """NCHW-shaped tensor -> NHWC view (no copy for channels_last memory or channel slices of it)."""
    v = x.permute(0, 2, 3, 1)
    try:
        nhwc_ld(v)
        return v
    except RuntimeError:
        return v.contiguous()


def as_nchw(x_nhwc):
    return x_nhwc.permute(0, 3, 1, 2)


def out_size(h, k, stride, pad):
    return (h + 2 * pad - k) // stride + 1


def _splits(c, n_pixels, k_split=False):
    """C output columns as launches of (tile width, column blocks). Layers with enough 256-pixel tiles to fill the
    SMs take the widest tiles (most reuse of the pixel operand); small BEV maps take the widest width that still
    gives ~100 (tile, block) work items, down to 64 columns."""
    tiles = max(1, n_pixels // 256)
    if tiles >= 120:
        widths = _TILE_N
    else:
        # small maps run 128-pixel work items (the kernels halve their tiles when that fills more SMs): the widest
        # column tile that still gives ~100 items wins (128-column MMAs run at twice the rate of 64-column ones)
        if os.environ.get("DBEV_SPLIT_HALF_TILES", "1") != "0":
            tiles = max(1, n_pixels // 128)
        fit = [wdt for wdt in _TILE_N if c % wdt == 0 and tiles * (c // wdt) >= 100]
        if k_split and (not fit or fit[0] == 64) and os.environ.get("DBEV_HALO_KSPLIT", "1") != "0":
            # 3x3 / stride 1 layers with a long K: the halo kernel splits K in two when its grid would be half empty
            # (csrc/conv2d_tc.cu HaloShape::ksplit), so 128-column tiles still give ~100 work items
            fit2 = [wdt for wdt in _TILE_N if wdt > 64 and c % wdt == 0 and 2 * tiles * (c // wdt) >= 100]
            fit = fit2[:1] or fit
        widths = (fit[0],) if fit else ((64,) if c % 64 == 0 else _TILE_N)
    parts, left = [], c
    for wdt in widths:
        if left >= wdt:
            parts.append((wdt, left // wdt))
            left -= wdt * (left // wdt)
    if left:
        raise RuntimeError("channel count %d is not a sum of 256 / 128 / 64 column blocks" % c)
    return parts


def pack_weights(weight, mode):
    """torch [C_out, C_in, KH, KW] -> flat fp32 buffer of K-major matrices (mode 0 fwd, 1 dgrad s1, 2 dgrad s2)."""
    lib = _lib.load()
    _lib.require_cuda(weight, "weight", torch.float32)
    w = weight.detach().contiguous()
    co, ci, kh, kw = w.shape
    out = torch.empty(w.numel(), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        rc = lib.dbev_pack_conv_weights(_lib.ptr(w), co, ci, kh, kw, mode, _lib.ptr(out), _lib.stream_ptr(w.device))
    _lib.check(rc, "dbev_pack_conv_weights")
    return out


def pack_weights_train(weight, stride):
    """(forward matrix, input-gradient matrix) of one layer from one pass over the weights."""
    lib = _lib.load()
    _lib.require_cuda(weight, "weight", torch.float32)
    w = weight.detach().contiguous()
    co, ci, kh, kw = w.shape
    out = torch.empty((2, w.numel()), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        rc = lib.dbev_pack_conv_weights_train(_lib.ptr(w), co, ci, kh, kw, 1 if stride == 1 else 2, _lib.ptr(out[0]),
                                              _lib.ptr(out[1]), _lib.stream_ptr(w.device))
    _lib.check(rc, "dbev_pack_conv_weights_train")
    return out[0], out[1]


PACK_JOB_WORDS = 8


def pack_weights_batch(jobs, total_tiles):
    """All layers in one launch. ``jobs``: int64 device tensor [n_jobs, 8] of {w, out_fwd, out_dgrad, C_out, C_in,
    KH * 256 + KW, dgrad_mode, first tile} (include/distill_bev_b200.h: dbev_pack_conv_weights_batch)."""
    lib = _lib.load()
    _lib.require_cuda(jobs, "jobs", torch.int64)
    if jobs.dim() != 2 or jobs.shape[1] != PACK_JOB_WORDS or not jobs.is_contiguous():
        raise ValueError("pack_weights_batch: jobs must be a contiguous [n_jobs, %d] int64 tensor" % PACK_JOB_WORDS)
    with torch.cuda.device(jobs.device):
        rc = lib.dbev_pack_conv_weights_batch(_lib.ptr(jobs), jobs.shape[0], int(total_tiles), _lib.stream_ptr(jobs.device))
    _lib.check(rc, "dbev_pack_conv_weights_batch")


def _conv_launch(x, wmat, n_cols, kh, kw, stride, pad, out, shift=None, relu=False, out_mul=1, out_add=(0, 0),
                 force=(0, 0), accumulate=False):
    """One family of launches: x NHWC (C_in = x.shape[3]) * wmat [n_cols, kh*kw*C_in] -> out NHWC view [..., n_cols]."""
    lib = _lib.load()
    n, h, w, c_in = x.shape
    x_ld, o_ld = nhwc_ld(x, "x"), nhwc_ld(out, "out")
    oh, ow = out.shape[1], out.shape[2]
    c0 = 0
    with torch.cuda.device(x.device):
        st = _lib.stream_ptr(x.device)
        n_pix = n * (force[0] or out_size(h, kh, stride, pad)) * (force[1] or out_size(w, kw, stride, pad))
        k_split = (kh == 3 and kw == 3 and stride == 1 and pad == 1 and c_in >= 256 and c_in % 64 == 0 and not relu
                   and out_mul == 1 and force == (0, 0) and h <= 16)
        for width, blocks in _splits(n_cols, n_pix, k_split):
            part = width * blocks
            o = out[..., c0:c0 + part]
            sh = shift[c0:c0 + part] if shift is not None else None
            rc = lib.dbev_conv2d_tc_forward_ex(_lib.ptr(x), n, h, w, c_in, x_ld, _lib.ptr(wmat[c0:c0 + part]), width, blocks,
                                               kh, kw, stride, pad, None, _lib.ptr(sh), 1 if relu else 0, _lib.ptr(o), oh, ow,
                                               o_ld, 0, out_mul, out_add[0], out_add[1], 0, 1, force[0], force[1],
                                               1 if accumulate else 0, st)
            _lib.check(rc, "dbev_conv2d_tc_forward_ex")
            c0 += part
    return out


def conv_forward(x, w_fwd, c_out, kh, kw, stride, pad, bias=None, out=None):
    """y = conv(x, W) (+ bias): x NHWC, w_fwd = pack_weights(W, 0). Returns NHWC [N, Ho, Wo, C_out]."""
    n, h, w, c_in = x.shape
    ho, wo = out_size(h, kh, stride, pad), out_size(w, kw, stride, pad)
    if out is None:
        out = torch.empty((n, ho, wo, c_out), dtype=torch.float32, device=x.device)
    return _conv_launch(x, w_fwd.view(c_out, kh * kw * c_in), c_out, kh, kw, stride, pad, out, shift=bias)


def conv_input_grad(dy, w_bwd, c_in, kh, kw, stride, pad, in_hw, out=None, accumulate=False):
    """dx = conv_transpose(dy, W): dy NHWC [N, Ho, Wo, C_out]; w_bwd = pack_weights(W, 1 if stride == 1 else 2).
    Stride 1: one convolution of dy with the flipped, transposed filter. Stride 2 (3x3, pad 1): the four parity
    classes of the input pixels are stride-1 convolutions of dy with 1 / 2 / 2 / 4 taps, written on the 2x lattice."""
    n, ho, wo, c_out = dy.shape
    h, w = in_hw
    if out is None:
        out = torch.empty((n, h, w, c_in), dtype=torch.float32, device=dy.device)
    if stride == 1:
        return _conv_launch(dy, w_bwd.view(c_in, kh * kw * c_out), c_in, kh, kw, 1, kh - 1 - pad, out, accumulate=accumulate)
    if not (stride == 2 and kh == 3 and kw == 3 and pad == 1 and h == 2 * ho and w == 2 * wo):
        raise RuntimeError("conv_input_grad: stride 2 needs a 3x3 / pad 1 filter and an even input size")
    if os.environ.get("DBEV_DGRAD_S2_MERGED", "1") != "0":
        # the four parity classes as work items of ONE launch per column split (csrc/conv2d_tc.cu: conv2d_tc_dgrad_s2)
        lib = _lib.load()
        dy_ld, o_ld = nhwc_ld(dy, "dy"), nhwc_ld(out, "out")
        c0 = 0
        with torch.cuda.device(dy.device):
            tiles = 4 * ((n * ho * wo + 127) // 128)        # (class, 128-pixel tile) pairs
            fit = [wdt for wdt in _TILE_N if c_in % wdt == 0 and tiles * (c_in // wdt) >= 100]
            parts = [(fit[0], c_in // fit[0])] if fit else _splits(c_in, n * ho * wo)
            for width, blocks in parts:
                rc = lib.dbev_conv2d_tc_dgrad_s2(_lib.ptr(dy), n, ho, wo, c_out, dy_ld, _lib.ptr(w_bwd), c_in, width, blocks,
                                                 _lib.ptr(out), o_ld, c0, 1 if accumulate else 0, _lib.stream_ptr(dy.device))
                _lib.check(rc, "dbev_conv2d_tc_dgrad_s2")
                c0 += width * blocks
        return out
    per = c_in * c_out
    for cls, off in enumerate((0, 1, 3, 5)):
        a, b = cls >> 1, cls & 1
        nky, nkx = 1 + a, 1 + b
        wm = w_bwd[off * per:(off + nky * nkx) * per].view(c_in, nky * nkx * c_out)
        _conv_launch(dy, wm, c_in, nky, nkx, 1, 0, out, out_mul=2, out_add=(a, b), force=(ho, wo), accumulate=accumulate)
    return out


def conv_weight_grad(x, dy, kh, kw, stride, pad, dw=None, accumulate=False):
    """dW [C_out, C_in, KH, KW] (+)= sum over pixels of dy (x) x on tcgen05 (C_in, C_out multiples of 128)."""
    lib = _lib.load()
    n, h, w, c_in = x.shape
    _, ho, wo, c_out = dy.shape
    x_ld, dy_ld = nhwc_ld(x, "x"), nhwc_ld(dy, "dy")
    if dw is None:
        dw = torch.empty((c_out, c_in, kh, kw), dtype=torch.float32, device=x.device)
        accumulate = False
    nbytes = lib.dbev_conv_wgrad_tc_workspace_bytes(n, ho, wo, c_in, c_out, kh, kw, stride)
    if nbytes == 0:
        raise RuntimeError("conv_weight_grad: C_in / C_out must be multiples of 128 (got %d, %d)" % (c_in, c_out))
    ws = _lib.workspace(nbytes, x.device)
    with torch.cuda.device(x.device):
        rc = lib.dbev_conv_wgrad_tc(_lib.ptr(x), n, h, w, c_in, x_ld, _lib.ptr(dy), ho, wo, c_out, dy_ld, kh, kw, stride, pad,
                                    _lib.ptr(dw), 1 if accumulate else 0, _lib.ptr(ws), nbytes, _lib.stream_ptr(x.device))
    _lib.check(rc, "dbev_conv_wgrad_tc")
    return dw


def stats_workspace(rows, c, device):
    """Per-block partial sums of a per-channel reduction. One workspace per concurrently running reduction."""
    return torch.empty(int(_lib.load().dbev_channel_stats_workspace_bytes(rows, c)), dtype=torch.uint8, device=device)


def _rows(t):
    return t.shape[0] * t.shape[1] * t.shape[2]


def bn_batch_stats(y, gamma, beta, eps, momentum=0.1, running_mean=None, running_var=None, ws=None):
    """[4, C] = (a, b, mean, invstd) of nn.BatchNorm2d in training mode over y NHWC; updates the running stats."""
    lib = _lib.load()
    c, rows, ld = y.shape[3], _rows(y), nhwc_ld(y, "y")
    ws = ws if ws is not None else stats_workspace(rows, c, y.device)
    out = torch.empty((4, c), dtype=torch.float32, device=y.device)
    with torch.cuda.device(y.device):
        rc = lib.dbev_bn_batch_stats(_lib.ptr(y), ld, rows, c, _lib.ptr(gamma), _lib.ptr(beta), float(eps), float(momentum),
                                     _lib.ptr(running_mean), _lib.ptr(running_var), _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                                     _lib.stream_ptr(y.device))
    _lib.check(rc, "dbev_bn_batch_stats")
    return out


def channel_sums(y, out=None, accumulate=False, ws=None):
    lib = _lib.load()
    c, rows, ld = y.shape[3], _rows(y), nhwc_ld(y, "y")
    ws = ws if ws is not None else stats_workspace(rows, c, y.device)
    if out is None:
        out, accumulate = torch.empty(c, dtype=torch.float32, device=y.device), False
    with torch.cuda.device(y.device):
        rc = lib.dbev_channel_sums(_lib.ptr(y), ld, rows, c, _lib.ptr(out), 1 if accumulate else 0, _lib.ptr(ws), ws.numel(),
                                   _lib.stream_ptr(y.device))
    _lib.check(rc, "dbev_channel_sums")
    return out


def bn_act(y, ab=None, residual=None, relu=True, out=None, want_mask=False):
    """out = relu?(a * y + b (+ residual)), NHWC; `out` may be a channel slice of a wider tensor. want_mask (with relu):
    also returns the ReLU mask [rows, C/4] uint8 (bit k of a byte = channel 4*quad + k was positive) that bn_backward /
    relu_backward take instead of the output - the backward then reads 1 byte instead of 16 per channel quad."""
    lib = _lib.load()
    c, rows, ld = y.shape[3], _rows(y), nhwc_ld(y, "y")
    if out is None:
        out = torch.empty(y.shape, dtype=torch.float32, device=y.device)
    r_ld = nhwc_ld(residual, "residual") if residual is not None else 0
    mask = torch.empty((rows, c // 4), dtype=torch.uint8, device=y.device) if (want_mask and relu) else None
    with torch.cuda.device(y.device):
        rc = lib.dbev_bn_act_forward(_lib.ptr(y), ld, _lib.ptr(ab), _lib.ptr(residual), r_ld, rows, c, 1 if relu else 0,
                                     _lib.ptr(out), nhwc_ld(out, "out"), _lib.ptr(mask), _lib.stream_ptr(y.device))
    _lib.check(rc, "dbev_bn_act_forward")
    return (out, mask) if want_mask else out


def bn_backward(dz, z, y, fwd, want_g=False, ws=None, mask=None):
    """Backward of z = relu?(BN(y) (+ identity)). z=None and mask=None: no ReLU; mask = bn_act(..., want_mask=True)[1]
    replaces z. Returns (dy, bwd[4, C] = dgamma, dbeta, .., g or None)."""
    lib = _lib.load()
    c, rows = y.shape[3], _rows(y)
    ws = ws if ws is not None else stats_workspace(rows, c, y.device)
    bwd = torch.empty((4, c), dtype=torch.float32, device=y.device)
    dy = torch.empty(y.shape, dtype=torch.float32, device=y.device)
    g = torch.empty(y.shape, dtype=torch.float32, device=y.device) if want_g else None
    if mask is not None:
        z = None
    with torch.cuda.device(y.device):
        rc = lib.dbev_bn_backward(_lib.ptr(dz), nhwc_ld(dz, "dz"), _lib.ptr(z), nhwc_ld(z, "z") if z is not None else 0,
                                  _lib.ptr(y), nhwc_ld(y, "y"), _lib.ptr(fwd), rows, c, _lib.ptr(bwd), _lib.ptr(dy), c,
                                  _lib.ptr(g), c, 0, _lib.ptr(mask), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(y.device))
    _lib.check(rc, "dbev_bn_backward")
    return dy, bwd, g


def relu_backward(dz, z=None, mask=None):
    """g = dz * (z > 0), from the forward output z or its ReLU mask."""
    lib = _lib.load()
    n, h, w, c = dz.shape
    rows = n * h * w
    g = torch.empty((n, h, w, c), dtype=torch.float32, device=dz.device)
    if mask is not None:
        z = None
    with torch.cuda.device(dz.device):
        rc = lib.dbev_relu_mask_backward(_lib.ptr(dz), nhwc_ld(dz, "dz"), _lib.ptr(z), nhwc_ld(z, "z") if z is not None else 0,
                                         rows, c, _lib.ptr(g), c, 0, _lib.ptr(mask), _lib.stream_ptr(dz.device))
    _lib.check(rc, "dbev_relu_mask_backward")
    return g


def upsample_bilinear(x, scale, out=None):
    """nn.Upsample(scale_factor=scale, mode='bilinear', align_corners=True) on NHWC."""
    lib = _lib.load()
    n, h, w, c = x.shape
    big_h, big_w = int(h * scale), int(w * scale)
    if out is None:
        out = torch.empty((n, big_h, big_w, c), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.dbev_upsample_bilinear_forward(_lib.ptr(x), nhwc_ld(x, "x"), n, h, w, c, big_h, big_w, _lib.ptr(out),
                                                nhwc_ld(out, "out"), _lib.stream_ptr(x.device))
    _lib.check(rc, "dbev_upsample_bilinear_forward")
    return out


def upsample_bilinear_backward(dout, in_hw, out=None, accumulate=False):
    lib = _lib.load()
    n, big_h, big_w, c = dout.shape
    h, w = in_hw
    if out is None:
        out, accumulate = torch.empty((n, h, w, c), dtype=torch.float32, device=dout.device), False
    with torch.cuda.device(dout.device):
        rc = lib.dbev_upsample_bilinear_backward(_lib.ptr(dout), nhwc_ld(dout, "dout"), n, h, w, c, big_h, big_w, _lib.ptr(out),
                                                 nhwc_ld(out, "out"), 1 if accumulate else 0, _lib.stream_ptr(dout.device))
    _lib.check(rc, "dbev_upsample_bilinear_backward")
    return out
