"""`mmdet3d.ops.bev_pool` and the LSS splat on the B200 plan/gather kernels.

Mirrors (same names, argument meaning and shapes):
  * ``bev_pool(feats, coords, B, D, H, W)``      mmdet3d/ops/bev_pool/bev_pool.py:83-97
  * ``bev_pool_ext.bev_pool_forward/backward``   mmdet3d/ops/bev_pool/src/bev_pool.cpp:22-87
  * ``voxel_pooling(geom_feats, x)``             mmdet3d/models/necks/view_transformer_mine.py:141-181
    (same result as ``voxel_pooling_accelerated`` :184-240)

All of them run the same CUDA kernels (distill-bev_b200/csrc/bev_pool.cu)
through the C-ABI; nothing here computes on the CPU.
"""
import torch

from ... import _lib


class BevPlan(object):
    """Sorted view of the frustum points over the BEV grid (geometry only).

    Reusable for every feature tensor that shares the geometry (e.g. cached
    calibration at inference time, or fwd + bwd of one training step).
    """

    __slots__ = ("order", "cell_start", "cell_end", "items", "n_items", "rows_per_item", "n_points",
                 "batch", "nz", "nslow", "nfast", "fast_axis", "device", "point_cell")

    def __init__(self, order, cell_start, cell_end, items, n_items, rows_per_item, n_points, batch,
                 nz, nslow, nfast, fast_axis):
        self.order = order
        self.cell_start = cell_start
        self.cell_end = cell_end
        self.items = items            # [max_items, 4] int32 work list (tile, c0 | c1 << 8, row_lo, row_hi)
        self.n_items = n_items        # [1] int32, device
        self.rows_per_item = rows_per_item
        self.n_points = n_points
        self.batch = batch
        self.nz = nz
        self.nslow = nslow
        self.nfast = nfast
        self.fast_axis = fast_axis
        self.device = order.device
        self.point_cell = None        # [n_points] int32 cell id / -1 (lift_splat backward)

    @property
    def n_cells(self):
        return self.batch * self.nz * self.nslow * self.nfast

    def num_kept(self):
        """Number of points inside the grid (device sync; for tests/diagnostics)."""
        tail = int(self.cell_start[self.n_cells].item())
        end = int(self.cell_end[self.n_cells].item())
        return self.n_points - (end - tail)


def _alloc_plan(n_points, batch, n0, n1, nz, fast_axis, device, rows_per_item=0):
    lib = _lib.load()
    n_cells = batch * n0 * n1 * nz
    if min(batch, n0, n1, nz) < 1:
        raise RuntimeError("bev plan: empty grid (batch=%d, %dx%dx%d)" % (batch, n0, n1, nz))
    order = torch.empty(max(n_points, 1), dtype=torch.int32, device=device)
    cell_start = torch.empty(n_cells + 1, dtype=torch.int32, device=device)
    cell_end = torch.empty(n_cells + 1, dtype=torch.int32, device=device)
    nslow, nfast = (n1, n0) if fast_axis == 0 else (n0, n1)
    max_items = lib.dbev_bev_plan_max_items(n_points, n_cells, nfast, rows_per_item)
    items = torch.empty((max_items, 4), dtype=torch.int32, device=device)
    n_items = torch.empty(1, dtype=torch.int32, device=device)
    return BevPlan(order, cell_start, cell_end, items, n_items, rows_per_item, n_points, batch, nz,
                   nslow, nfast, fast_axis)


class GridSpec(object):
    """Host copy of the BEV grid constants (off = bx - dx/2, dx, nx as fp32 and nx.to(long)).

    Built once per view transformer: reading ``bx/dx/nx`` CUDA parameters back on every call
    would put a device synchronisation on the hot path."""

    __slots__ = ("off", "dx", "nx_f", "nx_i")

    def __init__(self, bx, dx, nx):
        bx32 = torch.as_tensor(bx, dtype=torch.float32).detach().cpu()
        dx32 = torch.as_tensor(dx, dtype=torch.float32).detach().cpu()
        nx32 = torch.as_tensor(nx, dtype=torch.float32).detach().cpu()
        self.off = (bx32 - dx32 / 2.0).tolist()   # fp32, exactly as (self.bx - self.dx / 2.)
        self.dx = dx32.tolist()
        self.nx_f = nx32.tolist()
        self.nx_i = nx32.to(torch.long).tolist()   # nx.to(torch.long) truncates


def bev_plan_from_geom(geom, batch, bx=None, dx=None, nx=None, fast_axis=0, rows_per_item=0,
                       with_point_cell=False, grid=None):
    """Plan from ego-frame frustum coordinates.

    geom: [..., 3] fp32 CUDA tensor, batch-major (e.g. [B, N, D, fH, fW, 3]);
    bx, dx, nx: the float triples of ``gen_dx_bx`` (view_transformer_mine.py:14-18), or a
    cached ``GridSpec``. Index math, bounds test and batch index follow voxel_pooling :150-161.
    """
    lib = _lib.load()
    _lib.require_cuda(geom, "geom", torch.float32)
    geom = geom.contiguous()
    n_points = geom.numel() // 3
    if grid is None:
        grid = GridSpec(bx, dx, nx)
    nx_i = grid.nx_i
    plan = _alloc_plan(n_points, batch, int(nx_i[0]), int(nx_i[1]), int(nx_i[2]), fast_axis,
                       geom.device, rows_per_item)
    if with_point_cell:
        plan.point_cell = torch.empty(max(n_points, 1), dtype=torch.int32, device=geom.device)
    with torch.cuda.device(geom.device):
        ws_bytes = lib.dbev_bev_plan_workspace_bytes(n_points, plan.n_cells)
        ws = _lib.workspace(ws_bytes, geom.device)
        rc = lib.dbev_bev_plan_from_geom(
            _lib.ptr(geom), n_points, batch, _lib.host_f3(grid.off), _lib.host_f3(grid.dx),
            _lib.host_f3(grid.nx_f), _lib.host_i3(grid.nx_i), fast_axis, rows_per_item,
            _lib.ptr(plan.order), _lib.ptr(plan.cell_start), _lib.ptr(plan.cell_end),
            _lib.ptr(plan.items), plan.items.shape[0], _lib.ptr(plan.n_items),
            _lib.ptr(plan.point_cell), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(geom.device))
    _lib.check(rc, "dbev_bev_plan_from_geom")
    return plan


def bev_plan_from_coords(coords, B, D, H, W, fast_axis=1, rows_per_item=0):
    """Plan from integer coords [n, 4] = (c0 < H, c1 < W, c2 < D, batch < B)."""
    lib = _lib.load()
    _lib.require_cuda(coords, "coords")
    if coords.dim() != 2 or coords.shape[1] != 4:
        raise RuntimeError("coords must be [n, 4], got %s" % (tuple(coords.shape),))
    if coords.dtype not in (torch.int64, torch.int32):
        coords = coords.long()
    coords = coords.contiguous()
    n = coords.shape[0]
    plan = _alloc_plan(n, B, H, W, D, fast_axis, coords.device, rows_per_item)
    with torch.cuda.device(coords.device):
        ws_bytes = lib.dbev_bev_plan_workspace_bytes(n, plan.n_cells)
        ws = _lib.workspace(ws_bytes, coords.device)
        rc = lib.dbev_bev_plan_from_coords(
            _lib.ptr(coords), 1 if coords.dtype == torch.int64 else 0, n, B, H, W, D, fast_axis,
            rows_per_item, _lib.ptr(plan.order), _lib.ptr(plan.cell_start), _lib.ptr(plan.cell_end),
            _lib.ptr(plan.items), plan.items.shape[0], _lib.ptr(plan.n_items),
            _lib.ptr(ws), ws_bytes, _lib.stream_ptr(coords.device))
    _lib.check(rc, "dbev_bev_plan_from_coords")
    return plan


def _out_strides(plan, C, layout):
    plane = plan.nslow * plan.nfast
    if layout == "bz_c":      # [B, nz*C, slow, fast]   (voxel_pooling: cat(unbind(2), 1))
        return (plan.batch, plan.nz * C, plan.nslow, plan.nfast), plan.nz * C * plane, C * plane, plane
    if layout == "b_c_z":     # [B, C, nz, slow, fast]  (bev_pool: permute(0,4,1,2,3))
        return (plan.batch, C, plan.nz, plan.nslow, plan.nfast), C * plan.nz * plane, plane, plan.nz * plane
    if layout == "cl":        # [B, nz, slow, fast, C]  cells-major / channels-last rows (what the NHWC conv kernels read)
        return (plan.batch, plan.nz, plan.nslow, plan.nfast, C), plan.nz * plane * C, plane * C, 1
    raise ValueError(layout)


class _BevPoolGather(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, plan, layout):
        lib = _lib.load()
        _lib.require_cuda(x, "x", torch.float32)
        if x.dim() != 2 or x.shape[0] != plan.n_points:
            raise RuntimeError("x must be [%d, C], got %s" % (plan.n_points, tuple(x.shape)))
        x = x.contiguous()
        C = x.shape[1]
        shape, sB, sZ, sC = _out_strides(plan, C, layout)
        out = torch.empty(shape, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.dbev_bev_pool_gather_forward(
                _lib.ptr(x), C, _lib.ptr(plan.order), _lib.ptr(plan.cell_start),
                _lib.ptr(plan.cell_end), _lib.ptr(plan.items), _lib.ptr(plan.n_items),
                plan.batch, plan.nz, plan.nslow, plan.nfast, sB, sZ, sC,
                _lib.ptr(out), _lib.stream_ptr(x.device))
        _lib.check(rc, "dbev_bev_pool_gather_forward")
        ctx.plan = plan
        ctx.layout = layout
        ctx.C = C
        return out

    @staticmethod
    def backward(ctx, out_grad):
        lib = _lib.load()
        plan, C = ctx.plan, ctx.C
        out_grad = out_grad.contiguous().float()
        _, sB, sZ, sC = _out_strides(plan, C, ctx.layout)
        x_grad = torch.empty((plan.n_points, C), dtype=torch.float32, device=out_grad.device)
        if ctx.layout == "cl" and not (plan.point_cell is not None and C % 4 == 0):
            # no point -> cell map in the plan: back to the [B, nz*C, slow, fast] layout of the gather-backward kernel
            out_grad = out_grad.permute(0, 1, 4, 2, 3).reshape(plan.batch, plan.nz * C, plan.nslow, plan.nfast).contiguous()
            _, sB, sZ, sC = _out_strides(plan, C, "bz_c")
        elif plan.point_cell is not None and ctx.layout in ("bz_c", "cl") and C % 4 == 0:
            # point-centric: sequential row writes, cell rows gathered from the cells-major gradient
            g_cl = out_grad if ctx.layout == "cl" else transpose_batched(out_grad, plan.batch * plan.nz, C,
                                                                        plan.nslow * plan.nfast)
            with torch.cuda.device(out_grad.device):
                rc = lib.dbev_bev_pool_point_backward(_lib.ptr(g_cl), _lib.ptr(plan.point_cell), plan.n_points,
                                                      C, _lib.ptr(x_grad), _lib.stream_ptr(out_grad.device))
            _lib.check(rc, "dbev_bev_pool_point_backward")
            return x_grad, None, None
        with torch.cuda.device(out_grad.device):
            rc = lib.dbev_bev_pool_gather_backward(
                _lib.ptr(out_grad), C, _lib.ptr(plan.order), _lib.ptr(plan.cell_start),
                _lib.ptr(plan.cell_end), _lib.ptr(plan.items), _lib.ptr(plan.n_items),
                plan.batch, plan.nz, plan.nslow, plan.nfast, sB, sZ, sC,
                _lib.ptr(x_grad), _lib.stream_ptr(out_grad.device))
        _lib.check(rc, "dbev_bev_pool_gather_backward")
        return x_grad, None, None


def transpose_batched(x, batch, rows, cols):
    """[batch, rows, cols] -> [batch, cols, rows] (NCHW <-> channels-last helper kernel)."""
    lib = _lib.load()
    _lib.require_cuda(x, "x", torch.float32)
    x = x.contiguous()
    out = torch.empty((batch, cols, rows), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.dbev_transpose_batched(_lib.ptr(x), _lib.ptr(out), batch, rows, cols,
                                        _lib.stream_ptr(x.device))
    _lib.check(rc, "dbev_transpose_batched")
    return out


class _LiftSplat(torch.autograd.Function):

    @staticmethod
    def forward(ctx, depth, feat, plan):
        lib = _lib.load()
        _lib.require_cuda(depth, "depth", torch.float32)
        _lib.require_cuda(feat, "feat", torch.float32)
        BN, D, fH, fW = depth.shape
        C = feat.shape[1]
        if tuple(feat.shape) != (BN, C, fH, fW):
            raise RuntimeError("feat must be [%d, C, %d, %d], got %s" % (BN, fH, fW, tuple(feat.shape)))
        if plan.n_points != BN * D * fH * fW:
            raise RuntimeError("plan was built for %d points, frustum has %d"
                               % (plan.n_points, BN * D * fH * fW))
        depth = depth.contiguous()
        fhw = fH * fW
        feat_cl = transpose_batched(feat, BN, C, fhw)            # [BN, fH*fW, C]
        shape, sB, sZ, sC = _out_strides(plan, C, "bz_c")
        out = torch.empty(shape, dtype=torch.float32, device=depth.device)
        with torch.cuda.device(depth.device):
            rc = lib.dbev_lift_splat_forward(
                _lib.ptr(depth), _lib.ptr(feat_cl), C, D, fhw, _lib.ptr(plan.order),
                _lib.ptr(plan.cell_start), _lib.ptr(plan.cell_end), _lib.ptr(plan.items),
                _lib.ptr(plan.n_items), plan.batch, plan.nz, plan.nslow, plan.nfast, sB, sZ, sC,
                _lib.ptr(out), _lib.stream_ptr(depth.device))
        _lib.check(rc, "dbev_lift_splat_forward")
        ctx.plan = plan
        ctx.dims = (BN, C, D, fH, fW)
        ctx.save_for_backward(depth, feat_cl)
        return out

    @staticmethod
    def backward(ctx, out_grad):
        lib = _lib.load()
        plan = ctx.plan
        if plan.point_cell is None:
            raise RuntimeError("lift_splat backward needs a plan built with with_point_cell=True")
        depth, feat_cl = ctx.saved_tensors
        BN, C, D, fH, fW = ctx.dims
        fhw = fH * fW
        plane = plan.nslow * plan.nfast
        # [B, nz*C, ny, nx] -> cells-major [B*nz, plane, C]
        g_cl = transpose_batched(out_grad.contiguous().float(), plan.batch * plan.nz, C, plane)
        d_depth = torch.empty_like(depth)
        d_feat_cl = torch.empty_like(feat_cl)
        with torch.cuda.device(depth.device):
            rc = lib.dbev_lift_splat_backward(
                _lib.ptr(g_cl), _lib.ptr(depth), _lib.ptr(feat_cl), _lib.ptr(plan.point_cell),
                BN * fhw, C, D, fhw, _lib.ptr(d_depth), _lib.ptr(d_feat_cl),
                _lib.stream_ptr(depth.device))
        _lib.check(rc, "dbev_lift_splat_backward")
        d_feat = transpose_batched(d_feat_cl, BN, fhw, C).view(BN, C, fH, fW)
        return d_depth, d_feat, None


class PointCells(object):
    """Geometry-only companion of BevPlan for the sort-free lift+splat: ``point_cell`` [n_points]
    int32 = output cell of every frustum point (-1 = dropped), no sort / bounds / work list."""

    def __init__(self, point_cell, n_points, batch, nz, nslow, nfast, frames=1):
        self.point_cell, self.n_points, self.batch = point_cell, n_points, batch
        self.nz, self.nslow, self.nfast = nz, nslow, nfast
        self.n_cells = batch * nz * nslow * nfast
        self.frames = frames      # > 1: cells numbered [sample][y][x][frame] (frames concatenated along the channels)


def bev_point_cells(geom, batch, bx=None, dx=None, nx=None, fast_axis=0, grid=None, frames=1):
    """Index math + bounds test of voxel_pooling (:150-161) only: PointCells for ``lift_splat``.
    ``frames`` > 1 (BEVDepth4D, batch = samples * frames): the sort-free lift_splat then returns
    [samples, frames * C, ny, nx] - ``torch.cat`` of the per-frame BEV maps along the channels (bevdet.py:300-320)
    falls out of the cell numbering, no concat pass."""
    lib = _lib.load()
    _lib.require_cuda(geom, "geom", torch.float32)
    geom = geom.contiguous()
    n_points = geom.numel() // 3
    if grid is None:
        grid = GridSpec(bx, dx, nx)
    n0, n1, nz = int(grid.nx_i[0]), int(grid.nx_i[1]), int(grid.nx_i[2])
    pc = torch.empty(max(n_points, 1), dtype=torch.int32, device=geom.device)
    with torch.cuda.device(geom.device):
        rc = lib.dbev_bev_point_cells_frames(_lib.ptr(geom), n_points, batch, int(frames), _lib.host_f3(grid.off),
                                             _lib.host_f3(grid.dx), _lib.host_f3(grid.nx_f), _lib.host_i3(grid.nx_i),
                                             fast_axis, _lib.ptr(pc), _lib.stream_ptr(geom.device))
    _lib.check(rc, "dbev_bev_point_cells_frames")
    nslow, nfast = (n1, n0) if fast_axis == 0 else (n0, n1)
    return PointCells(pc, n_points, batch, nz, nslow, nfast, frames=int(frames))


class _LiftSplatAtomic(torch.autograd.Function):
    """Sort-free lift+splat (vector float reductions into the channels-last BEV map)."""

    @staticmethod
    def forward(ctx, depth, feat, cells):
        lib = _lib.load()
        _lib.require_cuda(depth, "depth", torch.float32)
        _lib.require_cuda(feat, "feat", torch.float32)
        BN, D, fH, fW = depth.shape
        C = feat.shape[1]
        if tuple(feat.shape) != (BN, C, fH, fW):
            raise RuntimeError("feat must be [%d, C, %d, %d], got %s" % (BN, fH, fW, tuple(feat.shape)))
        if cells.n_points != BN * D * fH * fW:
            raise RuntimeError("cells were built for %d points, frustum has %d" % (cells.n_points, BN * D * fH * fW))
        if cells.nz != 1:
            raise NotImplementedError("sort-free lift_splat needs a single z bin (BEV)")
        depth = depth.contiguous()
        fhw = fH * fW
        feat_cl = transpose_batched(feat, BN, C, fhw)
        out_cl = torch.empty((cells.batch, cells.nslow, cells.nfast, C), dtype=torch.float32, device=depth.device)
        with torch.cuda.device(depth.device):
            rc = lib.dbev_lift_splat_atomic_forward(_lib.ptr(depth), _lib.ptr(feat_cl), _lib.ptr(cells.point_cell),
                                                    BN * fhw, C, D, fhw, cells.n_cells, _lib.ptr(out_cl),
                                                    _lib.stream_ptr(depth.device))
        _lib.check(rc, "dbev_lift_splat_atomic_forward")
        ctx.cells = cells
        ctx.dims = (BN, C, D, fH, fW)
        ctx.save_for_backward(depth, feat_cl)
        if cells.frames > 1:                       # memory is [samples, ny, nx, frames, C]
            out_cl = out_cl.view(cells.batch // cells.frames, cells.nslow, cells.nfast, cells.frames * C)
        return out_cl.permute(0, 3, 1, 2)          # [B, C, ny, nx] in channels_last memory

    @staticmethod
    def backward(ctx, out_grad):
        lib = _lib.load()
        cells = ctx.cells
        depth, feat_cl = ctx.saved_tensors
        BN, C, D, fH, fW = ctx.dims
        fhw = fH * fW
        g_nhwc = out_grad.float().permute(0, 2, 3, 1)
        if g_nhwc.is_contiguous():
            g_cl = g_nhwc                             # channels_last upstream gradient: free
        else:
            g_cl = transpose_batched(out_grad.contiguous().float(), cells.batch // cells.frames, C * cells.frames,
                                     cells.nslow * cells.nfast)
        d_depth = torch.empty_like(depth)
        d_feat_cl = torch.empty_like(feat_cl)
        with torch.cuda.device(depth.device):
            rc = lib.dbev_lift_splat_backward(_lib.ptr(g_cl), _lib.ptr(depth), _lib.ptr(feat_cl),
                                              _lib.ptr(cells.point_cell), BN * fhw, C, D, fhw, _lib.ptr(d_depth),
                                              _lib.ptr(d_feat_cl), _lib.stream_ptr(depth.device))
        _lib.check(rc, "dbev_lift_splat_backward")
        d_feat = transpose_batched(d_feat_cl, BN, fhw, C).view(BN, C, fH, fW)
        return d_depth, d_feat, None


def lift_splat(depth_prob, img_feat, plan):
    """Fused lift + splat: depth_prob [B*N, D, fH, fW] (softmax depth), img_feat
    [B*N, C, fH, fW] -> BEV [B, C*nz, ny, nx], identical to
    ``voxel_pooling(geom, (depth.unsqueeze(1) * feat.unsqueeze(2)).view(B,N,C,D,fH,fW)
    .permute(0,1,3,4,5,2))`` (bevdet_distill_more.py:413-421) without materialising the
    volume. ``plan`` = bev_plan_from_geom(geom, B, ..., with_point_cell=True) (sorted plan: fixed
    summation order, bit-reproducible) or bev_point_cells(geom, B, ...) (sort-free: the splat uses
    vector float reductions, no plan is built; result in channels_last memory, sums in arbitrary order)."""
    if isinstance(plan, PointCells):
        return _LiftSplatAtomic.apply(depth_prob, img_feat, plan)
    return _LiftSplat.apply(depth_prob, img_feat, plan)


def bev_pool_gather(x, plan, layout="bz_c"):
    """Sum rows of x [n_points, C] into the plan's BEV cells (differentiable)."""
    return _BevPoolGather.apply(x, plan, layout)


def bev_pool(feats, coords, B, D, H, W):
    """Drop-in for ``mmdet3d.ops.bev_pool`` (bev_pool.py:83-97).

    feats [n, C] fp32, coords [n, 4] (x < H, y < W, z < D, b < B) -> [B, C, D, H, W].
    """
    assert feats.shape[0] == coords.shape[0]
    plan = bev_plan_from_coords(coords, B, D, H, W, fast_axis=1)
    return bev_pool_gather(feats, plan, layout="b_c_z")


def voxel_pooling(geom_feats, x, bx=None, dx=None, nx=None, plan=None, grid=None, channels_last=False):
    """Drop-in for ``ViewTransformerLiftSplatShoot.voxel_pooling`` (:141-181).

    geom_feats [B, N, D, H, W, 3] fp32 ego-frame xyz, x [B, N, D, H, W, C] ->
    [B, C * nz, ny, nx]. ``plan`` may carry a cached BevPlan for this geometry.
    ``channels_last=True`` (single z bin): the same values in ``torch.channels_last`` memory - every pooled cell is
    one coalesced row store of the gather kernel (no transposing epilogue) and the BEV encoder's NHWC conv kernels
    read it as is.
    """
    B, N, D, H, W, C = x.shape
    Nprime = B * N * D * H * W
    if plan is None:
        plan = bev_plan_from_geom(geom_feats, B, bx, dx, nx, fast_axis=0, grid=grid,
                                  with_point_cell=bool(x.requires_grad and torch.is_grad_enabled()))
    x = x.reshape(Nprime, C)
    if channels_last:
        if plan.nz != 1:
            raise NotImplementedError("voxel_pooling(channels_last=True) needs a single z bin (BEV)")
        out = bev_pool_gather(x, plan, layout="cl")                       # [B, 1, ny, nx, C]
        return out.view(plan.batch, plan.nslow, plan.nfast, C).permute(0, 3, 1, 2)
    return bev_pool_gather(x, plan, layout="bz_c")


# --- the pybind module the reference builds (setup.py:246-252) ---------------

class _BevPoolExt(object):
    """Tensor-level twin of ``mmdet3d.ops.bev_pool.bev_pool_ext``."""

    @staticmethod
    def bev_pool_forward(x, geom_feats, interval_lengths, interval_starts, b, d, h, w):
        lib = _lib.load()
        _lib.require_cuda(x, "x", torch.float32)
        _lib.require_cuda(geom_feats, "geom_feats", torch.int32)
        _lib.require_cuda(interval_lengths, "interval_lengths", torch.int32)
        _lib.require_cuda(interval_starts, "interval_starts", torch.int32)
        x, geom_feats = x.contiguous(), geom_feats.contiguous()
        interval_lengths, interval_starts = interval_lengths.contiguous(), interval_starts.contiguous()
        n, c = x.shape
        out = torch.empty((b, d, h, w, c), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.dbev_bev_pool_forward(
                b, d, h, w, n, c, interval_lengths.shape[0], _lib.ptr(x), _lib.ptr(geom_feats),
                _lib.ptr(interval_starts), _lib.ptr(interval_lengths), _lib.ptr(out), 1,
                _lib.stream_ptr(x.device))
        _lib.check(rc, "dbev_bev_pool_forward")
        return out

    @staticmethod
    def bev_pool_backward(out_grad, geom_feats, interval_lengths, interval_starts, b, d, h, w):
        lib = _lib.load()
        _lib.require_cuda(out_grad, "out_grad", torch.float32)
        _lib.require_cuda(geom_feats, "geom_feats", torch.int32)
        out_grad, geom_feats = out_grad.contiguous(), geom_feats.contiguous()
        interval_lengths, interval_starts = interval_lengths.contiguous(), interval_starts.contiguous()
        n, c = geom_feats.shape[0], out_grad.shape[4]
        x_grad = torch.empty((n, c), dtype=out_grad.dtype, device=out_grad.device)
        with torch.cuda.device(out_grad.device):
            rc = lib.dbev_bev_pool_backward(
                b, d, h, w, n, c, interval_lengths.shape[0], _lib.ptr(out_grad),
                _lib.ptr(geom_feats), _lib.ptr(interval_starts), _lib.ptr(interval_lengths),
                _lib.ptr(x_grad), 1, _lib.stream_ptr(out_grad.device))
        _lib.check(rc, "dbev_bev_pool_backward")
        return x_grad


bev_pool_ext = _BevPoolExt()


def intervals_from_sorted_ranks(ranks):
    """(interval_starts, interval_lengths) int32 of the runs of equal values in a sorted
    rank vector - the bookkeeping of QuickCumsumCuda.forward (bev_pool.py:40-46)."""
    n = ranks.shape[0]
    is_head = torch.ones(n, dtype=torch.bool, device=ranks.device)
    if n > 1:
        torch.ne(ranks[1:], ranks[:-1], out=is_head[1:])
    starts = is_head.nonzero(as_tuple=False).flatten().to(torch.int32)
    ends = torch.cat([starts[1:], starts.new_tensor([n])])
    return starts, ends - starts


class QuickCumsumCuda(torch.autograd.Function):
    """API twin of bev_pool.py:37-80 for callers that bring pre-sorted rows + ranks."""

    @staticmethod
    def forward(ctx, x, geom_feats, ranks, B, D, H, W):
        starts, lengths = intervals_from_sorted_ranks(ranks)
        geom_i32 = geom_feats.to(torch.int32)
        ctx.save_for_backward(starts, lengths, geom_i32)
        ctx.grid = (B, D, H, W)
        return bev_pool_ext.bev_pool_forward(x, geom_i32, lengths, starts, B, D, H, W)

    @staticmethod
    def backward(ctx, out_grad):
        starts, lengths, geom_i32 = ctx.saved_tensors
        grad = bev_pool_ext.bev_pool_backward(out_grad.contiguous(), geom_i32, lengths, starts,
                                              *ctx.grid)
        return (grad,) + (None,) * 6
