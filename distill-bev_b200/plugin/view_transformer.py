"""LSS view transformer on the B200 kernels.

Mirrors ``ViewTransformerLiftSplatShoot`` (mmdet3d/models/necks/view_transformer_mine.py:59-264):
same constructor arguments, same ``dx/bx/nx/frustum/D`` attributes and the methods the
DistillBEV detectors call directly (bevdet_distill_more.py:398-421): ``get_geometry``,
``voxel_pooling``, ``voxel_pooling_accelerated`` — plus ``lift_splat`` (fused, no volume).
``ViewTransformerLSSBEVDepth`` (:283-344) adds the BEVDepth sub-modules the detectors call one by one
(featnet / se / extra_depthnet / dcn / depthnet): small image-branch modules kept as torch modules, except
``extra_depthnet`` which is this package's ResNetForBEVDet.
"""
import torch
from torch import nn

from .. import _lib
from .ops import bev_pool as _bp


def gen_dx_bx(xbound, ybound, zbound):
    dx = torch.Tensor([row[2] for row in [xbound, ybound, zbound]])
    bx = torch.Tensor([row[0] + row[2] / 2.0 for row in [xbound, ybound, zbound]])
    nx = torch.Tensor([(row[1] - row[0]) / row[2] for row in [xbound, ybound, zbound]])
    return dx, bx, nx


def lss_geometry(frustum, rots, trans, intrins, post_rots, post_trans):
    """get_geometry (:111-139) as one kernel: -> [B, N, D, fH, fW, 3] fp32."""
    lib = _lib.load()
    for name, t in (("frustum", frustum), ("rots", rots), ("trans", trans), ("intrins", intrins),
                    ("post_rots", post_rots), ("post_trans", post_trans)):
        _lib.require_cuda(t, name, torch.float32)
    B, N, _ = trans.shape
    if intrins.shape[-1] != 3:
        raise NotImplementedError("4x4 (KITTI) intrinsics are not supported")
    D, fH, fW, _ = frustum.shape
    dev = frustum.device
    geom = torch.empty((B, N, D, fH, fW, 3), dtype=torch.float32, device=dev)
    mats = torch.empty((B * N, 18), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.dbev_lss_geometry(
            _lib.ptr(frustum.contiguous()), D * fH * fW, _lib.ptr(rots.contiguous()),
            _lib.ptr(trans.contiguous()), _lib.ptr(intrins.contiguous()),
            _lib.ptr(post_rots.contiguous()), _lib.ptr(post_trans.contiguous()), B * N,
            _lib.ptr(mats), _lib.ptr(geom), _lib.stream_ptr(dev))
    _lib.check(rc, "dbev_lss_geometry")
    return geom


class ViewTransformerLiftSplatShoot(nn.Module):

    def __init__(self, grid_config=None, data_config=None, numC_input=512, numC_Trans=64,
                 downsample=16, accelerate=True, **kwargs):
        super(ViewTransformerLiftSplatShoot, self).__init__()
        if grid_config is None:
            grid_config = {'xbound': [-51.2, 51.2, 0.8], 'ybound': [-51.2, 51.2, 0.8],
                           'zbound': [-10.0, 10.0, 20.0], 'dbound': [1.0, 60.0, 1.0]}
        self.grid_config = grid_config
        dx, bx, nx = gen_dx_bx(grid_config['xbound'], grid_config['ybound'], grid_config['zbound'])
        self.dx = nn.Parameter(dx, requires_grad=False)
        self.bx = nn.Parameter(bx, requires_grad=False)
        self.nx = nn.Parameter(nx, requires_grad=False)
        if data_config is None:
            data_config = {'input_size': (256, 704)}
        self.data_config = data_config
        self.downsample = downsample
        self.frustum = self.create_frustum()
        self.D = self.frustum.shape[0]
        self.numC_input = numC_input
        self.numC_Trans = numC_Trans
        self.depthnet = nn.Conv2d(numC_input, self.D + numC_Trans, kernel_size=1, padding=0)
        self.accelerate = accelerate
        self._grid = _bp.GridSpec(bx, dx, nx)   # host copy: no device read-back per call
        # dx / bx / nx are state_dict entries (as in the reference, which reads them on every call): a checkpoint
        # with another grid must refresh the host copy too
        self.register_load_state_dict_post_hook(ViewTransformerLiftSplatShoot._refresh_grid)

    @staticmethod
    def _refresh_grid(module, incompatible_keys):
        module._grid = _bp.GridSpec(module.bx.detach().cpu(), module.dx.detach().cpu(), module.nx.detach().cpu())

    def get_depth_dist(self, x):
        return x.softmax(dim=1)

    def create_frustum(self):
        ogfH, ogfW = self.data_config['input_size']
        fH, fW = ogfH // self.downsample, ogfW // self.downsample
        ds = torch.arange(*self.grid_config['dbound'], dtype=torch.float).view(-1, 1, 1).expand(-1, fH, fW)
        D = ds.shape[0]
        xs = torch.linspace(0, ogfW - 1, fW, dtype=torch.float).view(1, 1, fW).expand(D, fH, fW)
        ys = torch.linspace(0, ogfH - 1, fH, dtype=torch.float).view(1, fH, 1).expand(D, fH, fW)
        return nn.Parameter(torch.stack((xs, ys, ds), -1), requires_grad=False)

    def get_geometry(self, rots, trans, intrins, post_rots, post_trans, offset=None):
        if offset is not None:
            raise NotImplementedError("per-pixel depth offsets")
        return lss_geometry(self.frustum, rots, trans, intrins, post_rots, post_trans)

    def make_plan(self, geom, batch, with_point_cell=True):
        """Sort the frustum over the BEV grid once; reuse for every tensor sharing `geom`."""
        return _bp.bev_plan_from_geom(geom, batch, fast_axis=0, with_point_cell=with_point_cell,
                                      grid=self._grid)

    def make_cells(self, geom, batch, frames=1):
        """Geometry -> cell of every frustum point, no sort: input of the sort-free ``lift_splat``.
        ``frames`` > 1: batch = samples * frames and the splat returns the frames concatenated along the channels."""
        return _bp.bev_point_cells(geom, batch, fast_axis=0, grid=self._grid, frames=frames)

    def voxel_pooling(self, geom_feats, x, plan=None, channels_last=False):
        return _bp.voxel_pooling(geom_feats, x, plan=plan, grid=self._grid, channels_last=channels_last)

    voxel_pooling_accelerated = voxel_pooling   # same result as the scatter_sum path (:184-240)

    def lift_splat(self, geom, depth_prob, img_feat, batch, plan=None):
        if plan is None:
            plan = self.make_plan(geom, batch)
        return _bp.lift_splat(depth_prob, img_feat, plan)

    def forward(self, input):
        x, rots, trans, intrins, post_rots, post_trans = input[:6]
        B, N, C, H, W = x.shape
        x = self.depthnet(x.view(B * N, C, H, W))
        depth = self.get_depth_dist(x[:, :self.D])
        geom = self.get_geometry(rots, trans, intrins, post_rots, post_trans)
        return self.lift_splat(geom, depth, x[:, self.D:(self.D + self.numC_Trans)].contiguous(), B)


class SELikeModule(nn.Module):
    """view_transformer_mine.py:267-280 (camera-parameter gating of the depth branch): a 1x1 conv and a tiny MLP on
    16 x 44 maps - plain torch modules, same sub-module names / state_dict keys."""

    def __init__(self, in_channel=512, feat_channel=256, intrinsic_channel=33):
        super(SELikeModule, self).__init__()
        self.input_conv = nn.Conv2d(in_channel, feat_channel, kernel_size=1, padding=0)
        self.fc = nn.Sequential(nn.BatchNorm1d(intrinsic_channel), nn.Linear(intrinsic_channel, feat_channel), nn.Sigmoid())

    def forward(self, x, cam_params):
        x = self.input_conv(x)
        b, c, _, _ = x.shape
        y = self.fc(cam_params).view(b, c, 1, 1)
        return x * y.expand_as(x)


class ModulatedDeformConv2dPack(nn.Module):
    """mmcv 1.6.0 'DCNv2' (mmcv/ops/modulated_deform_conv.py; third party, parity unpinned): parameters ``weight``,
    ``bias``, ``conv_offset.{weight,bias}`` (zero-initialised offsets), forward through
    torchvision.ops.deform_conv2d with the sigmoid mask - same offset channel order (y, x interleaved per tap)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, deform_groups=1, bias=True):
        super().__init__()
        k = kernel_size
        self.stride, self.padding, self.dilation, self.deform_groups = stride, padding, dilation, deform_groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, k, k))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        n = in_channels * k * k
        nn.init.uniform_(self.weight, -1.0 / n ** 0.5, 1.0 / n ** 0.5)       # mmcv: uniform(-stdv, stdv), stdv = 1/sqrt(n)
        self.conv_offset = nn.Conv2d(in_channels, deform_groups * 3 * k * k, kernel_size=k, stride=stride, padding=padding,
                                     dilation=dilation, bias=True)
        nn.init.zeros_(self.conv_offset.weight)
        nn.init.zeros_(self.conv_offset.bias)

    def forward(self, x):
        from torchvision.ops import deform_conv2d
        out = self.conv_offset(x)
        o1, o2, mask = torch.chunk(out, 3, dim=1)
        return deform_conv2d(x, torch.cat((o1, o2), dim=1), self.weight, self.bias, stride=self.stride,
                             padding=self.padding, dilation=self.dilation, mask=torch.sigmoid(mask))


class ViewTransformerLSSBEVDepth(ViewTransformerLiftSplatShoot):
    """view_transformer_mine.py:283-344. Same constructor arguments, attributes (``featnet``, ``se``, ``extra_depthnet``,
    ``dcn``, ``depthnet``, ``loss_depth_weight`` - the detectors call them one by one, bevdet_distill_more.py:398-421)
    and state_dict keys. ``extra_depthnet`` is this package's ResNetForBEVDet (tcgen05 training kernels); the 1x1 convs,
    the SE gate and the deformable conv act on 16 x 44 image-feature maps and stay torch modules (image branch, outside
    SURVEY §8 (a)); geometry and lift+splat are the kernels of the base class - ``forward`` never builds the
    [B, N, D, fH, fW, C] volume."""

    def __init__(self, extra_depth_net, loss_depth_weight, se_config=dict(), dcn_config=dict(bias=True), **kwargs):
        super(ViewTransformerLSSBEVDepth, self).__init__(**kwargs)
        from .bev_encoder import ResNetForBEVDet
        self.loss_depth_weight = loss_depth_weight
        cfg = dict(extra_depth_net)
        if cfg.pop("type", "ResNetForBEVDet") != "ResNetForBEVDet":
            raise NotImplementedError("extra_depth_net type %r" % extra_depth_net.get("type"))
        self.extra_depthnet = ResNetForBEVDet(**cfg)
        ch = extra_depth_net['num_channels'][0]
        self.featnet = nn.Conv2d(self.numC_input, self.numC_Trans, kernel_size=1, padding=0)
        self.depthnet = nn.Conv2d(ch, self.D, kernel_size=1, padding=0)
        self.dcn = nn.Sequential(ModulatedDeformConv2dPack(ch, ch, kernel_size=3, stride=1, padding=1, dilation=1,
                                                           deform_groups=1, **dcn_config), nn.BatchNorm2d(ch))
        self.se = SELikeModule(self.numC_input, feat_channel=ch, **se_config)

    def forward(self, input):
        x, rots, trans, intrins, post_rots, post_trans, depth_gt = input
        B, N, C, H, W = x.shape
        x = x.view(B * N, C, H, W)
        img_feat = self.featnet(x)
        cam_params = torch.cat([intrins.reshape(B * N, -1), post_rots.reshape(B * N, -1), post_trans.reshape(B * N, -1),
                                rots.reshape(B * N, -1), trans.reshape(B * N, -1)], dim=1)
        depth_feat = self.se(x, cam_params)
        depth_feat = self.extra_depthnet(depth_feat)[0]
        depth_feat = self.dcn(depth_feat)
        depth_digit = self.depthnet(depth_feat)
        depth_prob = self.get_depth_dist(depth_digit)
        geom = self.get_geometry(rots, trans, intrins, post_rots, post_trans)
        bev_feat = self.lift_splat(geom, depth_prob, img_feat.contiguous(), B)      # lift + splat fused
        return bev_feat, depth_digit
