"""BEVDepth4D steps on either side of the view transform, same argument
meaning as the reference methods:

  shift_feature    BEVDetSequentialES.shift_feature   mmdet3d/models/detectors/bevdet.py:267-321
  get_depth_loss   BEVDepth_Base.get_depth_loss       mmdet3d/models/detectors/bevdet.py:397-417
"""
import torch
from torch.autograd import Function

from .. import _lib


def feature_transform(h, w, trans, rots, dx, bx, dtype=torch.float32):
    """tf [n, 3, 3] of shift_feature (:279-313): current-frame feature pixel -> adjacent-frame feature
    pixel. A handful of 4x4 / 3x3 matrix ops per sample, kept in torch exactly as the reference."""
    rots0, rots1 = rots
    trans0, trans1 = trans
    n, v = trans0.shape[0], trans0.shape[1]
    dev = trans0.device
    c02l0 = torch.zeros((n, v, 4, 4), dtype=dtype, device=dev)
    c02l0[:, :, :3, :3] = rots0
    c02l0[:, :, :3, 3] = trans0
    c02l0[:, :, 3, 3] = 1
    c12l0 = torch.zeros((n, v, 4, 4), dtype=dtype, device=dev)
    c12l0[:, :, :3, :3] = rots1
    c12l0[:, :, :3, 3] = trans1
    c12l0[:, :, 3, 3] = 1
    l02l1 = c02l0.matmul(torch.inverse(c12l0))[:, 0, :, :].view(n, 4, 4)
    keep = [0, 1, 3]                                   # drop z: align in the BEV plane only (:303)
    l02l1 = l02l1[:, keep, :][:, :, keep]
    feat2bev = torch.zeros((3, 3), dtype=dtype, device=dev)
    feat2bev[0, 0] = float(dx[0])
    feat2bev[1, 1] = float(dx[1])
    feat2bev[0, 2] = float(bx[0]) - float(dx[0]) / 2.
    feat2bev[1, 2] = float(bx[1]) - float(dx[1]) / 2.
    feat2bev[2, 2] = 1
    feat2bev = feat2bev.view(1, 3, 3)
    return torch.inverse(feat2bev).matmul(l02l1).matmul(feat2bev).contiguous()


class _ShiftFeature(Function):
    @staticmethod
    def forward(ctx, input, tf):
        lib = _lib.load()
        _lib.require_cuda(input, "input", torch.float32)
        input = input.contiguous()
        n, c, h, w = input.shape
        out = torch.empty_like(input)
        with torch.cuda.device(input.device):
            rc = lib.dbev_shift_feature_forward(_lib.ptr(input), _lib.ptr(tf), n, c, h, w, _lib.ptr(out),
                                                _lib.stream_ptr(input.device))
        _lib.check(rc, "dbev_shift_feature_forward")
        ctx.save_for_backward(tf)
        ctx.shape = (n, c, h, w)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        (tf,) = ctx.saved_tensors
        n, c, h, w = ctx.shape
        grad_out = grad_out.contiguous()
        grad_in = torch.empty(ctx.shape, dtype=torch.float32, device=grad_out.device)
        with torch.cuda.device(grad_out.device):
            rc = lib.dbev_shift_feature_backward(_lib.ptr(grad_out), _lib.ptr(tf), n, c, h, w, _lib.ptr(grad_in),
                                                 _lib.stream_ptr(grad_out.device))
        _lib.check(rc, "dbev_shift_feature_backward")
        return grad_in, None


def shift_feature(input, trans, rots, dx, bx, interpolation_mode="bilinear"):
    """input [n, c, h, w] adjacent-frame BEV feature; trans / rots = (current, adjacent) camera->lidar
    translations [n, v, 3] / rotations [n, v, 3, 3]; dx / bx of the view transformer."""
    if interpolation_mode != "bilinear":
        raise NotImplementedError("interpolation_mode=%r" % interpolation_mode)
    _lib.require_cuda(input, "input", torch.float32)
    n, c, h, w = input.shape
    tf = feature_transform(h, w, trans, rots, dx, bx, input.dtype)
    return _ShiftFeature.apply(input, tf)


class _DepthLoss(Function):
    @staticmethod
    def forward(ctx, logits, depth_gt, D, dmin, dstep, weight):
        lib = _lib.load()
        _lib.require_cuda(logits, "depth", torch.float32)
        _lib.require_cuda(depth_gt, "depth_gt", torch.float32)
        logits, depth_gt = logits.contiguous(), depth_gt.contiguous()
        hw = depth_gt.shape[-2] * depth_gt.shape[-1]
        bn = depth_gt.numel() // hw
        if logits.numel() != bn * D * hw:
            raise RuntimeError("depth logits %s do not match depth_gt %s x D=%d"
                               % (tuple(logits.shape), tuple(depth_gt.shape), D))
        loss = torch.empty((1,), dtype=torch.float32, device=logits.device)
        with torch.cuda.device(logits.device):
            wsb = lib.dbev_depth_loss_workspace_bytes()
            ws = _lib.workspace(wsb, logits.device)
            rc = lib.dbev_depth_loss_forward(_lib.ptr(logits), _lib.ptr(depth_gt), bn, D, hw, float(dmin),
                                             float(dstep), float(weight), _lib.ptr(loss), _lib.ptr(ws), wsb,
                                             _lib.stream_ptr(logits.device))
        _lib.check(rc, "dbev_depth_loss_forward")
        ctx.save_for_backward(logits, depth_gt)
        ctx.cfg = (bn, D, hw, float(dmin), float(dstep), float(weight))
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_loss):
        lib = _lib.load()
        logits, depth_gt = ctx.saved_tensors
        bn, D, hw, dmin, dstep, weight = ctx.cfg
        g = grad_loss.to(torch.float32).reshape(1).contiguous()
        grad = torch.empty_like(logits)
        with torch.cuda.device(logits.device):
            rc = lib.dbev_depth_loss_backward(_lib.ptr(logits), _lib.ptr(depth_gt), bn, D, hw, dmin, dstep, weight,
                                              _lib.ptr(g), _lib.ptr(grad), _lib.stream_ptr(logits.device))
        _lib.check(rc, "dbev_depth_loss_backward")
        return grad, None, None, None, None, None


def get_depth_loss(depth_gt, depth, D, dbound, loss_depth_weight):
    """depth_gt [B, N, H, W] (0 = no LiDAR return), depth = depth logits [B*N, D, H, W]; dbound =
    grid_config['dbound'] = (min, max, step)."""
    return _DepthLoss.apply(depth, depth_gt.float(), int(D), float(dbound[0]), float(dbound[2]),
                            float(loss_depth_weight))
