"""Teacher pillar path on the B200 kernels (csrc/pillar.cu).

Mirrors (same constructor arguments, parameter names and return shapes):
  * ``DynamicPillarFeatureNet``  mmdet3d/models/voxel_encoders/pillar_encoder.py:165-338
  * ``PointPillarsScatter``      mmdet3d/models/middle_encoders/pillar_scatter.py:10-102
  * ``DynamicCenterPoint.voxelize`` + ``extract_pts_feat`` up to the canvas
    (mmdet3d/models/detectors/dynamic_centerpoint.py:43-93) as ``pillar_canvas``.

In eval mode (the DistillBEV teacher is always frozen / eval,
bevdet_distill.py:1591-1597) the whole encoder is one fused kernel sequence; in
training mode (batch-norm needs batch statistics over all points) the forward is
composed from this package's DynamicScatter kernels plus the module's own
Linear / BatchNorm1d / ReLU, exactly the reference's data flow.
"""
import torch
from torch import nn

from .. import _lib
from .ops.voxel import DynamicScatter, grid_size as _grid_size


def fold_bn(bn):
    """eval-mode BatchNorm1d -> per-channel (scale, shift)."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    shift = bn.bias - bn.running_mean * scale
    return scale.detach().float().contiguous(), shift.detach().float().contiguous()


def pillar_encode(points, weight, bn_scale, bn_shift, voxel_size, point_cloud_range, batch_size,
                  coors=None, batch_offsets=None, return_point_coors=False, sync=True):
    """Fused eval-mode DynamicPillarFeatureNet. points [N, F] of the whole batch; either
    ``coors`` [N, 4] int32 (b, z, y, x) or ``batch_offsets`` [B + 1] int32 (voxelized on the fly).
    -> voxel_feats [M, nout], voxel_coors [M, 4] int32 (and point coors / device count)."""
    lib = _lib.load()
    _lib.require_cuda(points, "points", torch.float32)
    points = points.contiguous()
    n, f = points.shape
    nout = weight.shape[0]
    if weight.shape[1] != f + 5:
        raise RuntimeError("PFN weight must be [%d, %d] (raw + cluster(3) + center(2)), got %s"
                           % (nout, f + 5, tuple(weight.shape)))
    dev = points.device
    w = weight.detach().float().contiguous()
    vf = torch.empty((max(n, 1), nout), dtype=torch.float32, device=dev)
    vc = torch.empty((max(n, 1), 4), dtype=torch.int32, device=dev)
    cnt = torch.empty(1, dtype=torch.int32, device=dev)
    pc = torch.empty((max(n, 1), 4), dtype=torch.int32, device=dev) if return_point_coors else None
    if coors is not None:
        _lib.require_cuda(coors, "coors", torch.int32)
        coors = coors.contiguous()
    if batch_offsets is not None:
        batch_offsets = batch_offsets.to(torch.int32)
        batch_offsets = batch_offsets.contiguous() if batch_offsets.is_cuda else _lib.h2d_async(batch_offsets, dev)
    x_off = float(voxel_size[0]) / 2 + float(point_cloud_range[0])   # python doubles, :88-89
    y_off = float(voxel_size[1]) / 2 + float(point_cloud_range[1])
    with torch.cuda.device(dev):
        wsb = lib.dbev_pillar_encode_workspace_bytes(n)
        ws = _lib.workspace(wsb, dev)
        rc = lib.dbev_pillar_encode(
            _lib.ptr(points), _lib.ptr(batch_offsets), _lib.ptr(coors), int(batch_size), n, f,
            _lib.host_floats(voxel_size), _lib.host_floats(point_cloud_range), x_off, y_off,
            _lib.ptr(w), nout, _lib.ptr(bn_scale), _lib.ptr(bn_shift), _lib.ptr(vf), _lib.ptr(vc),
            _lib.ptr(cnt), _lib.ptr(pc), _lib.ptr(ws), wsb, _lib.stream_ptr(dev))
    _lib.check(rc, "dbev_pillar_encode")
    if not sync:
        return vf, vc, cnt, pc
    m = int(cnt.item())
    out = (vf[:m], vc[:m])
    return out + ((pc[:n],) if return_point_coors else ())


def pillar_scatter(voxel_features, coors, batch_size, ny, nx, channels_last=False, count=None):
    """canvas [B, C, ny, nx] (NCHW, or the same shape in channels_last memory format)."""
    lib = _lib.load()
    _lib.require_cuda(voxel_features, "voxel_features", torch.float32)
    _lib.require_cuda(coors, "coors")
    voxel_features = voxel_features.contiguous()
    coors = coors.to(torch.int32).contiguous()
    m, c = voxel_features.shape
    dev = voxel_features.device
    fmt = torch.channels_last if channels_last else torch.contiguous_format
    canvas = torch.empty((batch_size, c, ny, nx), dtype=torch.float32, device=dev, memory_format=fmt)
    with torch.cuda.device(dev):
        rc = lib.dbev_pillar_scatter(_lib.ptr(voxel_features), _lib.ptr(coors), _lib.ptr(count), m, c,
                                     batch_size, ny, nx, int(channels_last), 1, _lib.ptr(canvas),
                                     _lib.stream_ptr(dev))
    _lib.check(rc, "dbev_pillar_scatter")
    return canvas


class PFNLayer(nn.Module):
    """voxel_encoders/utils.py:107-181 (sub-module names ``norm`` / ``linear`` as in the reference)."""

    def __init__(self, in_channels, out_channels, norm_cfg=dict(type='BN1d', eps=1e-3, momentum=0.01), last_layer=False,
                 mode='max'):
        super().__init__()
        self.name = 'PFNLayer'
        self.last_vfe = last_layer
        if not self.last_vfe:
            out_channels = out_channels // 2
        self.units = out_channels
        self.norm = nn.BatchNorm1d(self.units, eps=norm_cfg.get('eps', 1e-5), momentum=norm_cfg.get('momentum', 0.1))
        self.linear = nn.Linear(in_channels, self.units, bias=False)
        assert mode in ['max', 'avg']
        self.mode = mode


def _rewrite_virtual_flag(features):
    """MVP virtual points (pillar_encoder.py:108-113 / :294-299): the second-to-last point feature is the virtual
    label, -1 for a virtual point; the encoder sees 1 for virtual and 0 for real points. In place, like the reference."""
    flag = features[..., -2]
    flag.copy_((flag == -1).to(features.dtype))
    return features


class PillarFeatureNet(nn.Module):
    """pillar_encoder.py:14-162: the hard-voxel pillar encoder of the shipped CenterPoint teacher
    (configs/_base_/models/centerpoint_02pillar_second_secfpn_nus.py:6-13). Same constructor arguments and
    ``state_dict`` keys (``pfn_layers.0.linear.weight``, ``pfn_layers.0.norm.*``). Eval mode (the DistillBEV teacher is
    frozen, bevdet_distill.py:1591-1597): one fused kernel (csrc/pillar_hard.cu). Training mode needs batch statistics
    over all points of all pillars and raises (train the teacher with the reference; distillation never does)."""

    def __init__(self, in_channels=4, feat_channels=(64,), with_distance=False, with_cluster_center=True,
                 with_voxel_center=True, voxel_size=(0.2, 0.2, 4), point_cloud_range=(0, -40, -3, 70.4, 40, 1),
                 norm_cfg=dict(type='BN1d', eps=1e-3, momentum=0.01), mode='max', legacy=True, virtual=False):
        super(PillarFeatureNet, self).__init__()
        assert len(feat_channels) > 0
        if len(feat_channels) != 1 or with_distance or not with_cluster_center or not with_voxel_center or mode != 'max':
            raise NotImplementedError("PillarFeatureNet: only the shipped teacher configuration is implemented: one PFN "
                                      "layer, cluster + voxel centre decorations, max pooling")
        self.legacy = legacy
        self.in_channels = in_channels + 5
        self._with_distance, self._with_cluster_center, self._with_voxel_center = False, True, True
        self.pfn_layers = nn.ModuleList([PFNLayer(self.in_channels, feat_channels[0], norm_cfg=norm_cfg, last_layer=True,
                                                  mode=mode)])
        self.vx, self.vy = voxel_size[0], voxel_size[1]
        self.x_offset = self.vx / 2 + point_cloud_range[0]
        self.y_offset = self.vy / 2 + point_cloud_range[1]
        self.point_cloud_range = point_cloud_range
        self.virtual = virtual

    def forward(self, features, num_points, coors, count=None):
        """features [M, max_points, F], num_points [M], coors [M, 4] (b, z, y, x) -> [M, nout]
        (``count``: optional device int = number of valid voxels when the buffers are over-allocated)."""
        if self.training:
            raise NotImplementedError("PillarFeatureNet: training mode (batch statistics) is not implemented; the "
                                      "distillation teacher is always eval()")
        lib = _lib.load()
        _lib.require_cuda(features, "features", torch.float32)
        if self.virtual:
            _rewrite_virtual_flag(features)
        features = features.contiguous()
        m, t, f = features.shape
        layer = self.pfn_layers[0]
        if layer.linear.weight.shape[1] != f + 5:
            raise RuntimeError("features have %d channels, the PFN layer expects %d" % (f, layer.linear.weight.shape[1] - 5))
        scale, shift = fold_bn(layer.norm)
        out = torch.empty((m, layer.units), dtype=torch.float32, device=features.device)
        npts = num_points.to(torch.int32).contiguous()
        co = coors.to(torch.int32).contiguous()
        with torch.cuda.device(features.device):
            rc = lib.dbev_hard_pillar_encode(_lib.ptr(features), _lib.ptr(npts), _lib.ptr(co), _lib.ptr(count), m, t, f,
                                             _lib.host_floats([self.vx, self.vy]), float(self.x_offset), float(self.y_offset),
                                             _lib.ptr(layer.linear.weight.detach().float().contiguous()), layer.units,
                                             _lib.ptr(scale), _lib.ptr(shift), 1 if self.legacy else 0, _lib.ptr(out),
                                             _lib.stream_ptr(features.device))
        _lib.check(rc, "dbev_hard_pillar_encode")
        return out.squeeze()


class DynamicPillarFeatureNet(nn.Module):
    """pillar_encoder.py:165-338 with one PFN layer (the only case the reference supports,
    ':219 TODO: currently only support one PFNLayer')."""

    def __init__(self, in_channels=4, feat_channels=(64,), with_distance=False,
                 with_cluster_center=True, with_voxel_center=True, voxel_size=(0.2, 0.2, 4),
                 point_cloud_range=(0, -40, -3, 70.4, 40, 1),
                 norm_cfg=dict(type='BN1d', eps=1e-3, momentum=0.01), mode='max', virtual=False,
                 act_cfg=dict(type='ReLU', inplace=True), use_checkpoint=False):
        super(DynamicPillarFeatureNet, self).__init__()
        if len(feat_channels) != 1 or with_distance or not with_cluster_center or not with_voxel_center or mode != 'max':
            raise NotImplementedError("only the shipped teacher configuration is implemented: one PFN "
                                      "layer, cluster + voxel centre decorations, max pooling")
        self.virtual = virtual
        self.in_channels = in_channels + 5
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.vx, self.vy = voxel_size[0], voxel_size[1]
        self.x_offset = self.vx / 2 + point_cloud_range[0]
        self.y_offset = self.vy / 2 + point_cloud_range[1]
        bn = nn.BatchNorm1d(feat_channels[0], eps=norm_cfg.get('eps', 1e-5),
                            momentum=norm_cfg.get('momentum', 0.1))
        self.pfn_layers = nn.ModuleList([nn.Sequential(
            nn.Linear(self.in_channels, feat_channels[0], bias=False), bn, nn.ReLU(inplace=True))])
        self.pfn_scatter = DynamicScatter(voxel_size, point_cloud_range, False)
        self.cluster_scatter = DynamicScatter(voxel_size, point_cloud_range, True)

    def forward(self, features, coors):
        batch_size = int(coors[-1, 0] + 1)
        if self.virtual:
            _rewrite_virtual_flag(features)
        lin, bn = self.pfn_layers[0][0], self.pfn_layers[0][1]
        if not self.training:
            scale, shift = fold_bn(bn)
            return pillar_encode(features, lin.weight, scale, shift, self.voxel_size,
                                 self.point_cloud_range, batch_size, coors=coors.to(torch.int32))
        # training mode: reference data flow on this package's scatter kernels
        coors = coors.to(torch.int32)
        voxel_mean, mean_coors = self.cluster_scatter(features, coors, batch_size)
        gx, gy, gz = _grid_size(self.voxel_size, self.point_cloud_range)
        key = lambda c: ((c[:, 0].long() * gz + c[:, 1].long()) * gy + c[:, 2].long()) * gx + c[:, 3].long()
        valid = (coors >= 0).all(dim=1)
        idx = torch.searchsorted(key(mean_coors), key(coors.clamp(min=0)))
        points_mean = voxel_mean[idx.clamp(max=max(voxel_mean.shape[0] - 1, 0))]
        points_mean = torch.where(valid[:, None], points_mean, torch.zeros_like(points_mean))
        f_cluster = features[:, :3] - points_mean[:, :3]
        f_center = torch.stack([features[:, 0] - (coors[:, 3].type_as(features) * self.vx + self.x_offset),
                                features[:, 1] - (coors[:, 2].type_as(features) * self.vy + self.y_offset)], 1)
        point_feats = self.pfn_layers[0](torch.cat([features, f_cluster, f_center], dim=-1))
        return self.pfn_scatter(point_feats, coors, batch_size)


class PointPillarsScatter(nn.Module):
    """pillar_scatter.py:10-102. ``channels_last=True`` returns the same [B, C, ny, nx] tensor in
    NHWC memory order (one contiguous 4*C-byte row per pillar; what cuDNN prefers downstream)."""

    def __init__(self, in_channels, output_shape, channels_last=False):
        super().__init__()
        self.output_shape = output_shape
        self.ny, self.nx = output_shape[0], output_shape[1]
        self.in_channels = in_channels
        self.channels_last = channels_last

    def forward(self, voxel_features, coors, batch_size=None):
        if batch_size is not None:
            return pillar_scatter(voxel_features, coors, batch_size, self.ny, self.nx, self.channels_last)
        # forward_single: coors = (y, x) at columns 1, 2 (:47); returns a list like the reference
        c4 = torch.zeros((coors.shape[0], 4), dtype=torch.int32, device=coors.device)
        c4[:, 2], c4[:, 3] = coors[:, 1], coors[:, 2]
        return [pillar_scatter(voxel_features, c4, 1, self.ny, self.nx, self.channels_last)]


_OFFSETS_CACHE = {}


def _device_offsets(offs, device):
    """Batch offsets [B + 1] int32 on the device; cached per (offsets, device): fixed-size clouds
    then cost no copy per step, and the call is CUDA-graph capturable."""
    key = (offs, str(device))
    t = _OFFSETS_CACHE.get(key)
    if t is None:
        if len(_OFFSETS_CACHE) > 256:
            _OFFSETS_CACHE.clear()
        t = _lib.h2d_async(torch.tensor(offs, dtype=torch.int32), device)
        _OFFSETS_CACHE[key] = t
    return t


def pillar_canvas(points_list, encoder, scatter):
    """points (list of [N_b, F] CUDA tensors) -> teacher pseudo image [B, C, ny, nx] with no host
    synchronisation: voxelization, pillar encoding and scatter are enqueued back to back
    (DynamicCenterPoint.voxelize + extract_pts_feat, dynamic_centerpoint.py:43-93)."""
    B = len(points_list)
    offs = [0]
    for p in points_list:
        offs.append(offs[-1] + p.shape[0])
    points = torch.cat(points_list, 0) if B > 1 else points_list[0]
    offsets = _device_offsets(tuple(offs), points.device)
    lin, bn = encoder.pfn_layers[0][0], encoder.pfn_layers[0][1]
    if encoder.training:
        raise RuntimeError("pillar_canvas is the frozen-teacher (eval) path")
    scale, shift = fold_bn(bn)
    lib = _lib.load()
    _lib.require_cuda(points, "points", torch.float32)
    points = points.contiguous()
    n, f = points.shape
    dev = points.device
    w = lin.weight.detach().float().contiguous()
    nout = w.shape[0]
    gx, gy, gz = _grid_size(encoder.voxel_size, encoder.point_cloud_range)
    if (gy, gx) != (scatter.ny, scatter.nx) or gz != 1:
        raise RuntimeError("pillar_canvas: encoder grid %s does not match the scatter output_shape %s"
                           % ((gz, gy, gx), (scatter.ny, scatter.nx)))
    fmt = torch.channels_last if scatter.channels_last else torch.contiguous_format
    canvas = torch.empty((B, nout, gy, gx), dtype=torch.float32, device=dev, memory_format=fmt)
    cnt = torch.empty(1, dtype=torch.int32, device=dev)
    x_off = float(encoder.voxel_size[0]) / 2 + float(encoder.point_cloud_range[0])
    y_off = float(encoder.voxel_size[1]) / 2 + float(encoder.point_cloud_range[1])
    with torch.cuda.device(dev):
        wsb = lib.dbev_pillar_encode_workspace_bytes(n)
        ws = _lib.workspace(wsb, dev)
        rc = lib.dbev_pillar_canvas(
            _lib.ptr(points), _lib.ptr(offsets), None, B, n, f,
            _lib.host_floats(encoder.voxel_size), _lib.host_floats(encoder.point_cloud_range), x_off, y_off,
            _lib.ptr(w), nout, _lib.ptr(scale), _lib.ptr(shift), int(scatter.channels_last), 1,
            _lib.ptr(canvas), _lib.ptr(cnt), _lib.ptr(ws), wsb, _lib.stream_ptr(dev))
    _lib.check(rc, "dbev_pillar_canvas")
    return canvas
