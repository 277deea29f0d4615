"""LSS view transformer on the B200 kernels.

Mirrors ``ViewTransformerLiftSplatShoot`` (mmdet3d/models/necks/view_transformer_mine.py:59-264):
same constructor arguments, same ``dx/bx/nx/frustum/D`` attributes and the methods the
DistillBEV detectors call directly (bevdet_distill_more.py:398-421): ``get_geometry``,
``voxel_pooling``, ``voxel_pooling_accelerated`` — plus ``lift_splat`` (fused, no volume).
The learned sub-modules of the BEVDepth variant (featnet / depthnet / SE / DCN, :283-344) are
dense convolutions outside this package's kernels and are not re-implemented here.
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
