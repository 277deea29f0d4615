"""Drop-in for ``mmdet3d.ops.spconv`` on the sparse LiDAR teacher path.

Mirrors, with the same names / arguments / return shapes / ``state_dict`` keys:
  ops.get_conv_output_size, ops.get_indice_pairs, ops.indice_conv   mmdet3d/ops/spconv/ops.py:20-131
  SparseConvTensor (+ dense, find_indice_pair)                        mmdet3d/ops/spconv/structure.py:20-69
  SparseModule, SparseSequential                                      mmdet3d/ops/spconv/modules.py:26-202
  SparseConvolution, SubMConv3d, SparseConv3d                         mmdet3d/ops/spconv/conv.py:60-229,...
  SparseBasicBlock, make_sparse_convmodule                            mmdet3d/ops/sparse_block.py:68-186

All arithmetic runs in libdistill_bev_b200.so (csrc/spconv.cu). The rulebook lives on the device
as an output-major neighbour table; `get_indice_pairs` / `indice_conv` convert from / to the
reference's (indice_pairs, indice_pair_num) tensors for callers of the raw ops. In eval mode a
SparseSequential (conv, BatchNorm1d, ReLU) and a SparseBasicBlock run as ONE kernel per conv
(BN folded into the epilogue's scale / shift, residual + ReLU fused).

Forward only: on the DistillBEV path the sparse encoder belongs to the frozen teacher
(bevdet_distill.py:1591-1610, run under no_grad). Transposed / inverse convs (SparseUNet) are not
on that path and raise NotImplementedError. There is no CPU fallback.
"""
from collections import OrderedDict

import numpy as np
import torch
from torch import nn
from torch.nn import init
from torch.nn.parameter import Parameter
import math

from ... import _lib


def _triple(v, ndim=3):
    if isinstance(v, (list, tuple)):
        return [int(x) for x in v]
    return [int(v)] * ndim


def get_conv_output_size(input_size, kernel_size, stride, padding, dilation):
    """ops.py:20-31."""
    output_size = []
    for i in range(len(input_size)):
        size = (input_size[i] + 2 * padding[i] - dilation[i] * (kernel_size[i] - 1) - 1) // stride[i] + 1
        output_size.append(1 if kernel_size[i] == -1 else size)
    return output_size


def _geom(ksize, stride, padding, dilation, in_shape, out_shape, batch_size):
    vals = list(ksize) + list(stride) + list(padding) + list(dilation) + [int(v) for v in in_shape] + \
        [int(v) for v in out_shape] + [int(batch_size)]
    return _lib.host_ints(vals)


def _check_indices(indices):
    _lib.require_cuda(indices, "indices", torch.int32)
    if indices.dim() != 2 or indices.shape[1] != 4:
        raise RuntimeError("indices must be [N, 4] (batch, z, y, x); only 3D sparse convs are supported")
    return indices.contiguous()


class Rulebook(object):
    """Output-major rulebook of one sparse conv: out_indices [n_out,4], nbr [kvol, n_out]."""

    def __init__(self, out_indices, nbr, kvol, n_in, out_shape):
        self.out_indices, self.nbr, self.kvol, self.n_in = out_indices, nbr, kvol, n_in
        self.out_shape = list(out_shape)

    @property
    def n_out(self):
        return self.out_indices.shape[0]

    def pairs(self):
        """-> (indice_pairs [kvol, 2, n_in] int32 (-1 filled), indice_pair_num [kvol] int32): the
        tensors get_indice_pairs returns in the reference (spconv_ops.h:56-59,140)."""
        lib = _lib.load()
        dev = self.nbr.device
        stride = max(self.n_in, 1)
        pairs = torch.empty((self.kvol, 2, stride), dtype=torch.int32, device=dev)
        num = torch.empty((self.kvol,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.dbev_spconv_pairs_from_table(_lib.ptr(self.nbr), self.kvol, self.n_out, stride,
                                                  _lib.ptr(pairs), _lib.ptr(num), _lib.stream_ptr(dev))
        _lib.check(rc, "dbev_spconv_pairs_from_table")
        return pairs[:, :, :self.n_in], num


def build_rulebook(indices, batch_size, spatial_shape, ksize, stride, padding, dilation, subm):
    """Neighbour table of a SubMConv3d / SparseConv3d (replaces getIndicePair, spconv_ops.h:28-141)."""
    lib = _lib.load()
    indices = _check_indices(indices)
    dev = indices.device
    ksize, dilation = _triple(ksize), _triple(dilation)
    if subm:
        stride, padding = [1, 1, 1], [k // 2 for k in ksize]      # spconv_ops.h:75-78
        out_shape = list(spatial_shape)
    else:
        stride, padding = _triple(stride), _triple(padding)
        out_shape = get_conv_output_size(spatial_shape, ksize, stride, padding, dilation)
    for d, s in zip(dilation, stride):
        assert any([s == 1, d == 1]), "don't support this."       # conv.py:92-93
    kvol = int(np.prod(ksize))
    n_in = indices.shape[0]
    geom = _geom(ksize, stride, padding, dilation, spatial_shape, out_shape, batch_size)
    with torch.cuda.device(dev):
        sp = _lib.stream_ptr(dev)
        if subm:
            wsb = lib.dbev_spconv_workspace_bytes(n_in, n_in)
            ws = _lib.workspace(wsb, dev)
            nbr = torch.empty((kvol, n_in), dtype=torch.int32, device=dev)
            rc = lib.dbev_spconv_table(_lib.ptr(indices), n_in, _lib.ptr(indices), n_in, geom,
                                       _lib.ptr(nbr), _lib.ptr(ws), wsb, sp)
            _lib.check(rc, "dbev_spconv_table")
            return Rulebook(indices, nbr, kvol, n_in, out_shape)
        max_out = int(lib.dbev_spconv_max_out(n_in, geom))
        wsb = lib.dbev_spconv_workspace_bytes(n_in, max_out)
        ws = _lib.workspace(wsb, dev)
        keys = torch.empty((max(max_out, 1),), dtype=torch.int32, device=dev)
        count = torch.empty((1,), dtype=torch.int32, device=dev)
        rc = lib.dbev_spconv_out_candidates(_lib.ptr(indices), n_in, geom, _lib.ptr(keys), max_out,
                                            _lib.ptr(count), _lib.ptr(ws), wsb, sp)
        _lib.check(rc, "dbev_spconv_out_candidates")
        n_out = int(count.item())     # the reference synchronises here too (spconv_ops.h:130-137)
        out_indices = torch.empty((n_out, 4), dtype=torch.int32, device=dev)
        nbr = torch.empty((kvol, n_out), dtype=torch.int32, device=dev)
        rc = lib.dbev_spconv_out_table(_lib.ptr(indices), n_in, geom, _lib.ptr(keys), n_out,
                                       _lib.ptr(out_indices), _lib.ptr(nbr), _lib.ptr(ws), wsb, sp)
        _lib.check(rc, "dbev_spconv_out_table")
    return Rulebook(out_indices, nbr, kvol, n_in, out_shape)


def pack_weights(weight):
    """[*k, Cin, Cout] fp32 -> (wt_hi, wt_lo) [kvol, Cout, Cin]: transposed TF32 head / remainder,
    the operand format of the tensor-core kernel (done once per layer: the teacher is frozen)."""
    lib = _lib.load()
    _lib.require_cuda(weight, "filters", torch.float32)
    c_in, c_out = weight.shape[-2], weight.shape[-1]
    w = weight.detach().reshape(-1, c_in, c_out).contiguous()
    kvol = w.shape[0]
    hi = torch.empty((kvol, c_out, c_in), dtype=torch.float32, device=w.device)
    lo = torch.empty_like(hi)
    with torch.cuda.device(w.device):
        rc = lib.dbev_spconv_pack_weights(_lib.ptr(w), kvol, c_in, c_out, _lib.ptr(hi), _lib.ptr(lo),
                                          _lib.stream_ptr(w.device))
    _lib.check(rc, "dbev_spconv_pack_weights")
    return hi, lo


def conv_table(features, weight, nbr, n_out, scale=None, shift=None, residual=None, relu=False,
               impl=None, packed=None):
    """out = act((sum_k features[nbr[k]] @ weight[k]) * scale + shift + residual): one kernel.

    impl: None = automatic (tcgen05 3xTF32 kernel when C_in, C_out in {32, 64, 128}, else the fp32
    FMA kernels), "fma" / "tc" force one. `packed` = pack_weights(weight) cached by the caller."""
    lib = _lib.load()
    _lib.require_cuda(features, "features", torch.float32)
    _lib.require_cuda(weight, "filters", torch.float32)
    if torch.is_grad_enabled() and (features.requires_grad or weight.requires_grad):
        raise NotImplementedError(
            "distill_bev_b200 spconv is forward-only (frozen teacher); call it under torch.no_grad()")
    features = features.contiguous()
    c_in, c_out = weight.shape[-2], weight.shape[-1]
    if features.shape[1] != c_in:
        raise RuntimeError("features have %d channels, filters expect %d" % (features.shape[1], c_in))
    kvol = nbr.shape[0]
    out = torch.empty((n_out, c_out), dtype=torch.float32, device=features.device)
    for t in (scale, shift, residual):
        if t is not None:
            _lib.require_cuda(t, "epilogue operand", torch.float32)
    if residual is not None:
        residual = residual.contiguous()
    use_tc = bool(lib.dbev_spconv_tc_supported(c_in, c_out, kvol)) if impl is None else impl == "tc"
    with torch.cuda.device(features.device):
        sp = _lib.stream_ptr(features.device)
        if use_tc:
            hi, lo = packed if packed is not None else pack_weights(weight)
            rc = lib.dbev_spconv_forward_tc(_lib.ptr(features), c_in, _lib.ptr(hi), _lib.ptr(lo), c_out,
                                            _lib.ptr(nbr), kvol, n_out, _lib.ptr(scale), _lib.ptr(shift),
                                            _lib.ptr(residual), 1 if relu else 0, _lib.ptr(out), sp)
            _lib.check(rc, "dbev_spconv_forward_tc")
        else:
            w = weight.detach().reshape(kvol, c_in, c_out).contiguous()
            rc = lib.dbev_spconv_forward(_lib.ptr(features), c_in, _lib.ptr(w), c_out, _lib.ptr(nbr), kvol,
                                         n_out, _lib.ptr(scale), _lib.ptr(shift), _lib.ptr(residual),
                                         1 if relu else 0, _lib.ptr(out), sp)
            _lib.check(rc, "dbev_spconv_forward")
    return out


def get_indice_pairs(indices, batch_size, spatial_shape, ksize=3, stride=1, padding=0, dilation=1,
                     out_padding=0, subm=False, transpose=False, grid=None):
    """ops.py:50-107 -> (outids, indice_pairs, indice_pair_num). Output voxels come in
    lexicographic (b,z,y,x) order — the order of the reference's CUDA branch (torch::_unique,
    spconv_ops.h:131); pairs of one offset in ascending output row."""
    if transpose:
        raise NotImplementedError("transposed sparse conv is not on the DistillBEV teacher path")
    rb = build_rulebook(indices, batch_size, spatial_shape, ksize, stride, padding, dilation, subm)
    pairs, num = rb.pairs()
    return rb.out_indices, pairs, num


def indice_conv(features, filters, indice_pairs, indice_pair_num, num_activate_out, inverse=False,
                subm=False):
    """ops.py:110-131 / indiceConv (spconv_ops.h:261-361), fp32."""
    lib = _lib.load()
    if filters.dtype != torch.float32:
        raise NotImplementedError("only float32 filters are supported")
    _lib.require_cuda(indice_pairs, "indice_pairs", torch.int32)
    indice_pairs = indice_pairs.contiguous()
    indice_pair_num = indice_pair_num.to(device=indice_pairs.device, dtype=torch.int32).contiguous()
    kvol, _, stride = indice_pairs.shape
    nbr = torch.empty((kvol, int(num_activate_out)), dtype=torch.int32, device=indice_pairs.device)
    with torch.cuda.device(indice_pairs.device):
        rc = lib.dbev_spconv_table_from_pairs(_lib.ptr(indice_pairs), _lib.ptr(indice_pair_num), kvol,
                                              stride, 1 if inverse else 0, int(num_activate_out),
                                              _lib.ptr(nbr), _lib.stream_ptr(indice_pairs.device))
    _lib.check(rc, "dbev_spconv_table_from_pairs")
    return conv_table(features, filters, nbr, int(num_activate_out))


def dense_from_sparse(features, indices, spatial_shape, batch_size):
    """SparseConvTensor.dense() folded with .view(N, C*D, H, W): [B, C*D, H, W]."""
    lib = _lib.load()
    _lib.require_cuda(features, "features", torch.float32)
    indices = _check_indices(indices)
    features = features.contiguous()
    Z, Y, X = [int(v) for v in spatial_shape]
    C = features.shape[1]
    out = torch.empty((int(batch_size), C * Z, Y, X), dtype=torch.float32, device=features.device)
    with torch.cuda.device(features.device):
        rc = lib.dbev_spconv_dense(_lib.ptr(features), _lib.ptr(indices), features.shape[0], C,
                                   int(batch_size), Z, Y, X, _lib.ptr(out),
                                   _lib.stream_ptr(features.device))
    _lib.check(rc, "dbev_spconv_dense")
    return out


class SparseConvTensor(object):
    """structure.py:20-69."""

    def __init__(self, features, indices, spatial_shape, batch_size, grid=None):
        self.features = features
        self.indices = indices
        self.spatial_shape = spatial_shape
        self.batch_size = batch_size
        self.indice_dict = {}
        self.grid = grid

    @property
    def spatial_size(self):
        return np.prod(self.spatial_shape)

    def find_indice_pair(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key)

    def dense(self, channels_first=True):
        Z, Y, X = [int(v) for v in self.spatial_shape]
        C = self.features.shape[1]
        d = dense_from_sparse(self.features, self.indices, self.spatial_shape, self.batch_size)
        d = d.view(self.batch_size, C, Z, Y, X)
        return d if channels_first else d.permute(0, 2, 3, 4, 1).contiguous()

    @property
    def sparity(self):
        return self.indices.shape[0] / np.prod(self.spatial_shape) / self.batch_size


class SparseModule(nn.Module):
    """modules.py:78-82: marker base class."""
    pass


def _fold_bn(bn, bias=None):
    """eval-mode BatchNorm1d (+ conv bias) as per-channel scale / shift. Cached on the module and
    rebuilt only when a parameter / running statistic changes (the teacher is frozen)."""
    ts = [t for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var, bias) if t is not None]
    key = tuple((t.data_ptr(), t._version) for t in ts)
    cache = getattr(bn, "_dbev_folded", None)
    if cache is not None and cache[0] == key:
        return cache[1]
    with torch.no_grad():
        scale = 1.0 / torch.sqrt(bn.running_var + bn.eps)
        if bn.affine:
            scale = scale * bn.weight
        shift = -bn.running_mean * scale
        if bn.affine:
            shift = shift + bn.bias
        if bias is not None:
            shift = shift + bias * scale
        folded = (scale.float().contiguous(), shift.float().contiguous())
    bn._dbev_folded = (key, folded)
    return folded


class SparseConvolution(SparseModule):
    """conv.py:60-229 (same constructor and parameters: weight [*k, Cin, Cout], optional bias)."""

    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0,
                 dilation=1, groups=1, bias=True, subm=False, output_padding=0, transposed=False,
                 inverse=False, indice_key=None, fused_bn=False):
        super(SparseConvolution, self).__init__()
        assert groups == 1
        if ndim != 3:
            raise NotImplementedError("only 3D sparse convolutions are on the DistillBEV path")
        self.ndim = ndim
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _triple(kernel_size)
        self.conv1x1 = np.prod(self.kernel_size) == 1
        self.stride, self.padding = _triple(stride), _triple(padding)
        self.dilation, self.output_padding = _triple(dilation), _triple(output_padding)
        for d, s in zip(self.dilation, self.stride):
            assert any([s == 1, d == 1]), "don't support this."
        if transposed or inverse:
            raise NotImplementedError("transposed / inverse sparse convs are not on the DistillBEV path")
        self.transposed, self.inverse = transposed, inverse
        self.groups, self.subm, self.indice_key, self.fused_bn = groups, subm, indice_key, fused_bn
        self.weight = Parameter(torch.Tensor(*self.kernel_size, in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.weight.shape[-2] * int(np.prod(self.kernel_size))   # _calculate_fan_in_and_fan_out_hwio
            bound = 1 / math.sqrt(fan_in)
            init.uniform_(self.bias, -bound, bound)

    impl = None     # None = automatic kernel choice; "fma" / "tc" force one (tests, benchmarks)

    def _packed_weights(self, kvol):
        """TF32 hi/lo operands of the tensor-core kernel, rebuilt only when the weight changes."""
        lib = _lib.load()
        if self.impl == "fma" or not self.weight.is_cuda or not (
                self.impl == "tc" or lib.dbev_spconv_tc_supported(self.in_channels, self.out_channels, kvol)):
            return None
        key = (self.weight.data_ptr(), self.weight._version)
        cache = getattr(self, "_packed", None)
        if cache is None or cache[0] != key:
            cache = (key, pack_weights(self.weight))
            self._packed = cache
        return cache[1]

    def rulebook(self, input):
        datas = input.find_indice_pair(self.indice_key)
        if self.indice_key is not None and datas is not None:
            return datas
        # A submanifold rulebook depends only on the active coordinates and the kernel, so un-keyed
        # SubM convs over the same voxel set (SparseBasicBlock's conv1 / conv2, and every block of a
        # stage) share one table; the reference rebuilds it per conv (indice_key=None, conv.py:165-183).
        auto_key = None
        if self.subm and self.indice_key is None:
            auto_key = ("subm", tuple(self.kernel_size), tuple(self.dilation), input.indices.data_ptr(),
                        int(input.indices.shape[0]))
            rb = input.indice_dict.get(auto_key)
            if rb is not None:
                return rb
        rb = build_rulebook(input.indices, input.batch_size, input.spatial_shape, self.kernel_size,
                            self.stride, self.padding, self.dilation, self.subm)
        input.indice_dict[self.indice_key if auto_key is None else auto_key] = rb
        return rb

    def forward(self, input, scale=None, shift=None, residual=None, relu=False):
        """`scale/shift/residual/relu` are the fused epilogue (used by SparseSequential and
        SparseBasicBlock in eval mode); a plain call is exactly SparseConvolution.forward."""
        assert isinstance(input, SparseConvTensor)
        if self.conv1x1:
            feats = torch.mm(input.features, self.weight.view(self.in_channels, self.out_channels))
            if self.bias is not None:
                feats = feats + self.bias
            out = SparseConvTensor(feats, input.indices, input.spatial_shape, input.batch_size)
            out.indice_dict, out.grid = input.indice_dict, input.grid
            return out
        rb = self.rulebook(input)
        if scale is None and shift is None and self.bias is not None:
            shift = self.bias.detach()
        feats = conv_table(input.features, self.weight, rb.nbr, rb.n_out, scale, shift, residual, relu,
                           impl=self.impl, packed=self._packed_weights(rb.kvol))
        out = SparseConvTensor(feats, rb.out_indices, rb.out_shape, input.batch_size)
        out.indice_dict, out.grid = input.indice_dict, input.grid
        return out


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, indice_key=None):
        super(SparseConv3d, self).__init__(3, in_channels, out_channels, kernel_size, stride, padding,
                                           dilation, groups, bias, indice_key=indice_key)


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, indice_key=None):
        super(SubMConv3d, self).__init__(3, in_channels, out_channels, kernel_size, stride, padding,
                                         dilation, groups, bias, True, indice_key=indice_key)


CONV_TYPES = {"SparseConv3d": SparseConv3d, "SubMConv3d": SubMConv3d}


class SparseSequential(SparseModule):
    """modules.py:85-202 (same container semantics and child names)."""

    def __init__(self, *args, **kwargs):
        super(SparseSequential, self).__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)
        for name, module in kwargs.items():
            if name in self._modules:
                raise ValueError("name exists.")
            self.add_module(name, module)

    def __getitem__(self, idx):
        if not (-len(self) <= idx < len(self)):
            raise IndexError("index {} is out of range".format(idx))
        if idx < 0:
            idx += len(self)
        return list(self._modules.values())[idx]

    def __len__(self):
        return len(self._modules)

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError("name exists")
        self.add_module(name, module)

    def forward(self, input):
        mods = list(self._modules.values())
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, SparseConvolution) and isinstance(input, SparseConvTensor) and not m.conv1x1:
                # eval-mode conv -> BatchNorm1d -> ReLU runs as one kernel
                bn = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm1d) and \
                    not mods[i + 1].training else None
                if bn is not None:
                    relu = i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
                    scale, shift = _fold_bn(bn, m.bias)
                    input = m(input, scale=scale, shift=shift, relu=relu)
                    i += 3 if relu else 2
                    continue
            if isinstance(m, SparseModule):
                assert isinstance(input, SparseConvTensor)
                input = m(input)
            elif isinstance(input, SparseConvTensor):
                if input.indices.shape[0] != 0:
                    input.features = m(input.features)
            else:
                input = m(input)
            i += 1
        return input


class SparseBasicBlock(SparseModule):
    """ops/sparse_block.py:68-121 on mmdet 2.24 BasicBlock's module layout (conv1, bn1, conv2,
    bn2, relu, downsample — third party, parity unpinned at that boundary)."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, conv_cfg=None, norm_cfg=None):
        super(SparseBasicBlock, self).__init__()
        conv_cfg = dict(conv_cfg or dict(type="SubMConv3d"))
        cls = CONV_TYPES[conv_cfg.pop("type")]
        norm_cfg = dict(norm_cfg or dict(type="BN1d"))
        assert norm_cfg.pop("type") in ("BN1d", "BN")
        norm_cfg.pop("requires_grad", None)
        self.conv1 = cls(inplanes, planes, 3, stride=stride, padding=1, dilation=1, bias=False, **conv_cfg)
        self.bn1 = nn.BatchNorm1d(planes, **norm_cfg)
        self.conv2 = cls(planes, planes, 3, padding=1, bias=False, **conv_cfg)
        self.bn2 = nn.BatchNorm1d(planes, **norm_cfg)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    @property
    def norm1(self):
        return self.bn1

    @property
    def norm2(self):
        return self.bn2

    def forward(self, x):
        identity = x.features
        assert x.features.dim() == 2, "x.features.dim()=%d" % x.features.dim()
        if self.downsample is not None:
            identity = self.downsample(x)
        if not self.bn1.training and not self.bn2.training:
            s1, b1 = _fold_bn(self.bn1)
            s2, b2 = _fold_bn(self.bn2)
            out = self.conv1(x, scale=s1, shift=b1, relu=True)
            return self.conv2(out, scale=s2, shift=b2, residual=identity, relu=True)
        out = self.conv1(x)
        out.features = self.relu(self.bn1(out.features))
        out = self.conv2(out)
        out.features = self.relu(self.bn2(out.features) + identity)
        return out


def make_sparse_convmodule(in_channels, out_channels, kernel_size, indice_key, stride=1, padding=0,
                           conv_type="SubMConv3d", norm_cfg=None, order=("conv", "norm", "act")):
    """ops/sparse_block.py:124-186."""
    assert isinstance(order, tuple) and len(order) <= 3
    assert set(order) | {"conv", "norm", "act"} == {"conv", "norm", "act"}
    layers = []
    for layer in order:
        if layer == "conv":
            layers.append(CONV_TYPES[conv_type](in_channels, out_channels, kernel_size, stride=stride,
                                                padding=padding, bias=False, indice_key=indice_key))
        elif layer == "norm":
            cfg = dict(norm_cfg)
            assert cfg.pop("type") in ("BN1d", "BN")
            cfg.pop("requires_grad", None)
            layers.append(nn.BatchNorm1d(out_channels, **cfg))
        elif layer == "act":
            layers.append(nn.ReLU(inplace=True))
    return SparseSequential(*layers)
