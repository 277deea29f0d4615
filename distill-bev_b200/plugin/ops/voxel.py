"""`mmdet3d.ops.voxel` on the B200 kernels (distill-bev_b200/csrc/voxelize.cu).

Mirrors, with the same names / argument meaning / return shapes:
  * ``Voxelization``, ``voxelization``            mmdet3d/ops/voxel/voxelize.py:13-148
  * ``DynamicScatter``, ``dynamic_scatter``       mmdet3d/ops/voxel/scatter_points.py:9-107
  * ``voxel_layer.{hard_voxelize, dynamic_voxelize, dynamic_point_to_voxel_forward,
    dynamic_point_to_voxel_backward}``            mmdet3d/ops/voxel/src/voxelization.cpp:6-11

CUDA only: a CPU tensor raises (the reference itself raises "do not support
cpu yet" for dynamic scatter, voxelization.h:118). Integer outputs are
bit-exact with the reference's CPU build, which tests/ pin through oracle/.
"""
import torch
from torch import nn
from torch.autograd import Function

from ... import _lib

_REDUCE = {"sum": 0, "mean": 1, "max": 2}


def _reduce_id(reduce_type):
    if reduce_type not in _REDUCE:
        raise RuntimeError("do not support reduce type " + str(reduce_type))  # voxelization.h:103
    return _REDUCE[reduce_type]


def _prep_points(points):
    _lib.require_cuda(points, "points")
    if points.dim() != 2 or points.shape[1] < 3:
        raise RuntimeError("points must be [N, >=3], got %s" % (tuple(points.shape),))
    if points.dtype != torch.float32:
        raise RuntimeError("points must be float32 (got %s)" % points.dtype)
    return points.contiguous()


def grid_size(voxel_size, coors_range):
    """(gx, gy, gz) = round((max - min) / voxel) in fp32, as the reference computes it."""
    lib = _lib.load()
    out = (_lib.ctypes.c_int * 3)()
    rc = lib.dbev_voxel_grid_size(_lib.host_floats(voxel_size), _lib.host_floats(coors_range), out)
    _lib.check(rc, "dbev_voxel_grid_size")
    return int(out[0]), int(out[1]), int(out[2])


class _VoxelLayer(object):
    """Tensor-level twin of the pybind module ``mmdet3d.ops.voxel.voxel_layer``."""

    @staticmethod
    def dynamic_voxelize(points, coors, voxel_size, coors_range, NDim=3):
        lib = _lib.load()
        if NDim != 3:
            raise RuntimeError("only NDim=3 is supported")
        points = _prep_points(points)
        _lib.require_cuda(coors, "coors", torch.int32)
        if not coors.is_contiguous() or tuple(coors.shape) != (points.shape[0], 3):
            raise RuntimeError("coors must be a contiguous [N, 3] int32 tensor")
        with torch.cuda.device(points.device):
            rc = lib.dbev_dynamic_voxelize(
                _lib.ptr(points), points.shape[0], points.shape[1], _lib.host_floats(voxel_size),
                _lib.host_floats(coors_range), _lib.ptr(coors), _lib.stream_ptr(points.device))
        _lib.check(rc, "dbev_dynamic_voxelize")

    @staticmethod
    def hard_voxelize_device(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range,
                             max_points, max_voxels):
        """Like hard_voxelize but returns the voxel count as a device tensor (no sync)."""
        lib = _lib.load()
        points = _prep_points(points)
        _lib.require_cuda(voxels, "voxels", torch.float32)
        _lib.require_cuda(coors, "coors", torch.int32)
        _lib.require_cuda(num_points_per_voxel, "num_points_per_voxel", torch.int32)
        n, f = points.shape
        if (not voxels.is_contiguous() or voxels.shape[0] < max_voxels or voxels.shape[1] != max_points
                or voxels.shape[2] != f or not coors.is_contiguous() or coors.shape[0] < max_voxels
                or not num_points_per_voxel.is_contiguous() or num_points_per_voxel.shape[0] < max_voxels):
            raise RuntimeError("hard_voxelize: output buffers must be contiguous and sized "
                               "[max_voxels, max_points, F] / [max_voxels, 3] / [max_voxels]")
        count = torch.empty(1, dtype=torch.int32, device=points.device)
        with torch.cuda.device(points.device):
            wsb = lib.dbev_hard_voxelize_workspace_bytes(n)
            ws = _lib.workspace(wsb, points.device)
            rc = lib.dbev_hard_voxelize(
                _lib.ptr(points), n, f, _lib.host_floats(voxel_size), _lib.host_floats(coors_range),
                int(max_points), int(max_voxels), _lib.ptr(voxels), _lib.ptr(coors),
                _lib.ptr(num_points_per_voxel), _lib.ptr(count), _lib.ptr(ws), wsb,
                _lib.stream_ptr(points.device))
        _lib.check(rc, "dbev_hard_voxelize")
        return count

    @staticmethod
    def hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range,
                      max_points, max_voxels, NDim=3, deterministic=True):
        if NDim != 3:
            raise RuntimeError("only NDim=3 is supported")
        count = _VoxelLayer.hard_voxelize_device(points, voxels, coors, num_points_per_voxel,
                                                 voxel_size, coors_range, max_points, max_voxels)
        return int(count.item())

    @staticmethod
    def dynamic_point_to_voxel_forward(feats, coors, reduce_type, dims=None):
        """-> [reduced_feats, out_coors, coors_map, reduce_count] (scatter_points_cuda.cu:183-239).

        ``dims`` (exclusive upper bound per coors column) avoids a device->host
        read of coors.max(); DynamicScatter passes it from its voxel grid.
        """
        lib = _lib.load()
        _lib.require_cuda(feats, "feats", torch.float32)
        _lib.require_cuda(coors, "coors", torch.int32)
        if not feats.is_contiguous():
            raise RuntimeError("feats must be contiguous")       # CHECK_CONTIGUOUS, :11-15
        if not coors.is_contiguous():
            raise RuntimeError("coors must be contiguous")
        rid = _reduce_id(reduce_type)
        n, c = feats.shape
        ncol = coors.shape[1]
        if n == 0:
            return [feats.clone().detach(), coors.clone().detach(),
                    coors.new_empty((0,), dtype=torch.int32), coors.new_empty((0,), dtype=torch.int32)]
        if dims is None:
            dims = (coors.max(dim=0)[0] + 1).clamp_(min=1).tolist()
        reduced = torch.empty((n, c), dtype=torch.float32, device=feats.device)
        out_coors = torch.empty((n, ncol), dtype=torch.int32, device=feats.device)
        coors_map = torch.empty((n,), dtype=torch.int32, device=feats.device)
        reduce_count = torch.empty((n,), dtype=torch.int32, device=feats.device)
        num_out = torch.empty(1, dtype=torch.int32, device=feats.device)
        with torch.cuda.device(feats.device):
            wsb = lib.dbev_dynamic_scatter_workspace_bytes(n)
            ws = _lib.workspace(wsb, feats.device)
            rc = lib.dbev_dynamic_scatter_forward(
                _lib.ptr(feats), _lib.ptr(coors), n, c, ncol, _lib.host_ints(dims), rid,
                _lib.ptr(reduced), _lib.ptr(out_coors), _lib.ptr(coors_map), _lib.ptr(reduce_count),
                _lib.ptr(num_out), _lib.ptr(ws), wsb, _lib.stream_ptr(feats.device))
        _lib.check(rc, "dbev_dynamic_scatter_forward")
        m = int(num_out.item())
        return [reduced[:m], out_coors[:m], coors_map, reduce_count[:m]]

    @staticmethod
    def dynamic_point_to_voxel_backward(grad_feats, grad_reduced_feats, feats, reduced_feats,
                                        coors_idx, reduce_count, reduce_type):
        lib = _lib.load()
        for name, t in (("grad_feats", grad_feats), ("grad_reduced_feats", grad_reduced_feats),
                        ("feats", feats), ("reduced_feats", reduced_feats)):
            _lib.require_cuda(t, name, torch.float32)
            if not t.is_contiguous():
                raise RuntimeError(name + " must be contiguous")
        rid = _reduce_id(reduce_type)
        n, c = feats.shape
        m = reduced_feats.shape[0]
        coors_idx = coors_idx.contiguous()
        reduce_count = reduce_count.contiguous()
        ws = None
        if rid == 2 and m > 0:
            ws = torch.empty((m * c,), dtype=torch.int32, device=feats.device)
        with torch.cuda.device(feats.device):
            rc = lib.dbev_dynamic_scatter_backward(
                _lib.ptr(grad_reduced_feats), _lib.ptr(feats), _lib.ptr(reduced_feats),
                _lib.ptr(coors_idx), _lib.ptr(reduce_count), n, m, c, rid, _lib.ptr(grad_feats),
                _lib.ptr(ws), _lib.stream_ptr(feats.device))
        _lib.check(rc, "dbev_dynamic_scatter_backward")


voxel_layer = _VoxelLayer()


class _Voxelization(Function):
    """voxelize.py:13-70."""

    @staticmethod
    def forward(ctx, points, voxel_size, coors_range, max_points=35, max_voxels=20000,
                deterministic=True):
        if max_points == -1 or max_voxels == -1:
            coors = torch.empty((points.size(0), 3), dtype=torch.int32, device=points.device)
            voxel_layer.dynamic_voxelize(points, coors, voxel_size, coors_range, 3)
            return coors
        # the kernel writes every slot of the voxels it returns: no zero-fill of the
        # max_voxels x max_points x F buffer (12-18 MB per call in the reference, :57-61)
        voxels = torch.empty((max_voxels, max_points, points.size(1)), dtype=points.dtype,
                             device=points.device)
        coors = torch.empty((max_voxels, 3), dtype=torch.int32, device=points.device)
        num_points_per_voxel = torch.empty((max_voxels,), dtype=torch.int32, device=points.device)
        voxel_num = voxel_layer.hard_voxelize(points, voxels, coors, num_points_per_voxel,
                                              voxel_size, coors_range, max_points, max_voxels, 3,
                                              deterministic)
        return voxels[:voxel_num], coors[:voxel_num], num_points_per_voxel[:voxel_num]


voxelization = _Voxelization.apply


def _pair(v):
    return v if isinstance(v, tuple) else (v, v)


class Voxelization(nn.Module):
    """voxelize.py:76-148 (same constructor, attributes and train/eval max_voxels switch)."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000,
                 deterministic=True):
        super(Voxelization, self).__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.max_num_points = max_num_points
        self.max_voxels = _pair(max_voxels)
        self.deterministic = deterministic
        pcr = torch.tensor(point_cloud_range, dtype=torch.float32)
        vs = torch.tensor(voxel_size, dtype=torch.float32)
        grid = torch.round((pcr[3:] - pcr[:3]) / vs).long()
        self.grid_size = grid
        self.pcd_shape = [*grid[:2], 1][::-1]

    def forward(self, input):
        max_voxels = self.max_voxels[0] if self.training else self.max_voxels[1]
        return voxelization(input, self.voxel_size, self.point_cloud_range, self.max_num_points,
                            max_voxels, self.deterministic)

    def __repr__(self):
        return (self.__class__.__name__ + "(voxel_size=" + str(self.voxel_size) +
                ", point_cloud_range=" + str(self.point_cloud_range) + ", max_num_points=" +
                str(self.max_num_points) + ", max_voxels=" + str(self.max_voxels) +
                ", deterministic=" + str(self.deterministic) + ")")


class _dynamic_scatter(Function):
    """scatter_points.py:9-47."""

    @staticmethod
    def forward(ctx, feats, coors, reduce_type="max", dims=None):
        voxel_feats, voxel_coors, point2voxel_map, voxel_points_count = \
            voxel_layer.dynamic_point_to_voxel_forward(feats, coors, reduce_type, dims)
        ctx.reduce_type = reduce_type
        ctx.save_for_backward(feats, voxel_feats, point2voxel_map, voxel_points_count)
        ctx.mark_non_differentiable(voxel_coors)
        return voxel_feats, voxel_coors

    @staticmethod
    def backward(ctx, grad_voxel_feats, grad_voxel_coors=None):
        feats, voxel_feats, point2voxel_map, voxel_points_count = ctx.saved_tensors
        grad_feats = torch.empty_like(feats)
        voxel_layer.dynamic_point_to_voxel_backward(
            grad_feats, grad_voxel_feats.contiguous(), feats, voxel_feats, point2voxel_map,
            voxel_points_count, ctx.reduce_type)
        return grad_feats, None, None, None


def dynamic_scatter(feats, coors, reduce_type="max", dims=None):
    return _dynamic_scatter.apply(feats, coors, reduce_type, dims)


class DynamicScatter(nn.Module):
    """scatter_points.py:53-107. Batched coors [N, 4] = (b, z, y, x) are handled by ONE
    kernel sequence (batch is the most significant sort digit), which returns exactly the
    per-sample concatenation the reference builds in its Python loop (:86-97)."""

    def __init__(self, voxel_size, point_cloud_range, average_points):
        super(DynamicScatter, self).__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.average_points = average_points
        self._grid = None

    def _dims3(self):
        if self._grid is None:
            gx, gy, gz = grid_size(self.voxel_size, self.point_cloud_range)
            self._grid = (gz, gy, gx)
        return self._grid

    def forward_single(self, points, coors):
        reduce = "mean" if self.average_points else "max"
        return dynamic_scatter(points.contiguous(), coors.contiguous(), reduce, list(self._dims3()))

    def forward(self, points, coors, batch_size=None):
        if coors.size(-1) == 3:
            return self.forward_single(points, coors)
        if batch_size is None:
            batch_size = int(coors[-1, 0] + 1)      # same rule as the reference (:84)
        reduce = "mean" if self.average_points else "max"
        dims = [batch_size] + list(self._dims3())
        return dynamic_scatter(points.contiguous(), coors.contiguous(), reduce, dims)

    def __repr__(self):
        return (self.__class__.__name__ + "(voxel_size=" + str(self.voxel_size) +
                ", point_cloud_range=" + str(self.point_cloud_range) + ", average_points=" +
                str(self.average_points) + ")")
