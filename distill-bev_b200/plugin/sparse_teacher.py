"""Sparse LiDAR teacher modules (LidarFormer / MVPFormer front end), same registry names,
constructor arguments and ``state_dict`` keys as the reference:

  SparseEncoder        mmdet3d/models/middle_encoders/sparse_encoder.py:11-204
  HardSimpleVFE        mmdet3d/models/voxel_encoders/voxel_encoder.py:13-45
  DynamicVoxelEncoder  mmdet3d/models/voxel_encoders/dynamic_voxel_encoder.py:70-102

Every op is a kernel of libdistill_bev_b200.so (no CPU path, forward only — the teacher is frozen).
"""
import numpy as np
import torch
from torch import nn

from .. import _lib
from .ops import spconv
from .ops.spconv import SparseBasicBlock, make_sparse_convmodule
from .ops.voxel import voxel_layer


class SparseEncoder(nn.Module):
    """sparse_encoder.py:11-128. forward(voxel_features [N,C], coors [N,4] (b,z,y,x), batch_size)
    -> [B, C_out * D', H', W'] dense BEV features."""

    def __init__(self, in_channels, sparse_shape, order=("conv", "norm", "act"),
                 norm_cfg=dict(type="BN1d", eps=1e-3, momentum=0.01), base_channels=16,
                 output_channels=128,
                 encoder_channels=((16,), (32, 32, 32), (64, 64, 64), (64, 64, 64)),
                 encoder_paddings=((1,), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1)),
                 block_type="conv_module"):
        super().__init__()
        assert block_type in ["conv_module", "basicblock"]
        self.sparse_shape = sparse_shape
        self.in_channels = in_channels
        self.order = order
        self.base_channels = base_channels
        self.output_channels = output_channels
        self.encoder_channels = encoder_channels
        self.encoder_paddings = encoder_paddings
        self.stage_num = len(self.encoder_channels)
        self.fp16_enabled = False
        assert isinstance(order, tuple) and len(order) == 3
        assert set(order) == {"conv", "norm", "act"}
        if self.order[0] != "conv":  # pre activate
            self.conv_input = make_sparse_convmodule(in_channels, self.base_channels, 3,
                                                     norm_cfg=norm_cfg, padding=1, indice_key="subm1",
                                                     conv_type="SubMConv3d", order=("conv",))
        else:
            self.conv_input = make_sparse_convmodule(in_channels, self.base_channels, 3,
                                                     norm_cfg=norm_cfg, padding=1, indice_key="subm1",
                                                     conv_type="SubMConv3d")
        encoder_out_channels = self.make_encoder_layers(make_sparse_convmodule, norm_cfg,
                                                        self.base_channels, block_type=block_type)
        self.conv_out = make_sparse_convmodule(encoder_out_channels, self.output_channels,
                                               kernel_size=(3, 1, 1), stride=(2, 1, 1),
                                               norm_cfg=norm_cfg, padding=0,
                                               indice_key="spconv_down2", conv_type="SparseConv3d")

    @torch.no_grad()
    def forward(self, voxel_features, coors, batch_size):
        coors = coors.int()
        x = spconv.SparseConvTensor(voxel_features, coors, self.sparse_shape, int(batch_size))
        x = self.conv_input(x)
        encode_features = []
        for encoder_layer in self.encoder_layers:
            x = encoder_layer(x)
            encode_features.append(x)
        out = self.conv_out(encode_features[-1])
        # out.dense() + view(N, C*D, H, W) in one kernel (:121-127)
        return spconv.dense_from_sparse(out.features, out.indices, out.spatial_shape, out.batch_size)

    def make_encoder_layers(self, make_block, norm_cfg, in_channels, block_type="conv_module",
                            conv_cfg=dict(type="SubMConv3d")):
        """sparse_encoder.py:130-204."""
        assert block_type in ["conv_module", "basicblock"]
        self.encoder_layers = spconv.SparseSequential()
        for i, blocks in enumerate(self.encoder_channels):
            blocks_list = []
            for j, out_channels in enumerate(tuple(blocks)):
                padding = tuple(self.encoder_paddings[i])[j]
                if i != 0 and j == 0 and block_type == "conv_module":
                    blocks_list.append(make_block(in_channels, out_channels, 3, norm_cfg=norm_cfg,
                                                  stride=2, padding=padding,
                                                  indice_key="spconv%d" % (i + 1),
                                                  conv_type="SparseConv3d"))
                elif block_type == "basicblock":
                    if j == len(blocks) - 1 and i != len(self.encoder_channels) - 1:
                        blocks_list.append(make_block(in_channels, out_channels, 3, norm_cfg=norm_cfg,
                                                      stride=2, padding=padding,
                                                      indice_key="spconv%d" % (i + 1),
                                                      conv_type="SparseConv3d"))
                    else:
                        blocks_list.append(SparseBasicBlock(out_channels, out_channels,
                                                            norm_cfg=norm_cfg, conv_cfg=conv_cfg))
                else:
                    blocks_list.append(make_block(in_channels, out_channels, 3, norm_cfg=norm_cfg,
                                                  padding=padding, indice_key="subm%d" % (i + 1),
                                                  conv_type="SubMConv3d"))
                in_channels = out_channels
            self.encoder_layers.add_module("encoder_layer%d" % (i + 1),
                                           spconv.SparseSequential(*blocks_list))
        return out_channels


class HardSimpleVFE(nn.Module):
    """voxel_encoder.py:13-45: mean of the points of every voxel."""

    def __init__(self, num_features=4):
        super(HardSimpleVFE, self).__init__()
        self.num_features = num_features
        self.fp16_enabled = False

    def forward(self, features, num_points, coors=None):
        lib = _lib.load()
        _lib.require_cuda(features, "features", torch.float32)
        features = features.contiguous()
        num_points = num_points.to(device=features.device, dtype=torch.int32).contiguous()
        m, maxp, f = features.shape
        out = torch.empty((m, self.num_features), dtype=torch.float32, device=features.device)
        with torch.cuda.device(features.device):
            rc = lib.dbev_hard_simple_vfe(_lib.ptr(features), _lib.ptr(num_points), m, maxp, f,
                                          self.num_features, _lib.ptr(out),
                                          _lib.stream_ptr(features.device))
        _lib.check(rc, "dbev_hard_simple_vfe")
        return out


class DynamicVoxelEncoder(nn.Module):
    """dynamic_voxel_encoder.py:70-102. forward(points: list[Tensor [Np, 5 | 17]]) ->
    (voxels [M, 5 | 23] fp32, coors [M, 4] int64 (b,z,y,x), shape_np). The whole batch goes through
    one kernel sequence (coords -> sort -> segment mean) instead of a per-sample Python loop of
    boolean-mask compactions, unique(dim=0) and scatter_mean."""

    def __init__(self, pc_range, voxel_size, virtual=False):
        super(DynamicVoxelEncoder, self).__init__()
        self.pc_range = torch.tensor(pc_range)
        self.voxel_size = torch.tensor(voxel_size)
        self.shape = torch.round((self.pc_range[3:] - self.pc_range[:3]) / self.voxel_size)
        self.shape_np = self.shape.numpy().astype(np.int32)
        self.virtual = virtual

    @torch.no_grad()
    def forward(self, points):
        lib = _lib.load()
        batch = len(points)
        dev = points[0].device
        for p in points:
            _lib.require_cuda(p, "points", torch.float32)
        pts = torch.cat(points, dim=0).contiguous() if batch > 1 else points[0].contiguous()
        n, f = pts.shape
        offs = [0]
        for p in points:
            offs.append(offs[-1] + p.shape[0])
        offsets = _lib.h2d_async(torch.tensor(offs, dtype=torch.int32), dev)
        coors = torch.empty((n, 4), dtype=torch.int32, device=dev)
        pr = _lib.host_floats(self.pc_range.tolist())
        vs = _lib.host_floats(self.voxel_size.tolist())
        with torch.cuda.device(dev):
            sp = _lib.stream_ptr(dev)
            rc = lib.dbev_dynvoxel_coords(_lib.ptr(pts), n, f, _lib.ptr(offsets), batch, pr, vs,
                                          1 if self.virtual else 0, _lib.ptr(coors), sp)
            _lib.check(rc, "dbev_dynvoxel_coords")
            rows = pts
            if self.virtual:
                rows = torch.empty((n, 24), dtype=torch.float32, device=dev)
                rc = lib.dbev_dynvoxel_virtual_rows(_lib.ptr(pts), n, f, _lib.ptr(rows), sp)
                _lib.check(rc, "dbev_dynvoxel_virtual_rows")
        gx, gy, gz = [int(v) for v in self.shape_np]
        # both range ends are inclusive (:9-11): a point on the upper border gets index == size
        dims = [batch, gz + 1, gy + 1, gx + 1]
        voxels, out_coors, _, _ = voxel_layer.dynamic_point_to_voxel_forward(rows, coors, "mean", dims)
        if self.virtual:
            m = voxels.shape[0]
            fixed = torch.empty((m, 23), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                rc = lib.dbev_dynvoxel_virtual_fix(_lib.ptr(voxels.contiguous()), None, m,
                                                   _lib.ptr(fixed), _lib.stream_ptr(dev))
            _lib.check(rc, "dbev_dynvoxel_virtual_fix")
            voxels = fixed
        return voxels, out_coors.long(), self.shape_np
