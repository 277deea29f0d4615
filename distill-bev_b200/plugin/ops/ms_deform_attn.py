"""Multi-scale deformable attention on the B200 kernels (csrc/ms_deform_attn.cu): drop-in for
``MultiScaleDeformableAttnFunction_fp32`` (mmdet3d/models/transformer_modules/
multi_scale_deformable_attn_function.py:90-165), the op under BEVFormer's spatial cross attention
and temporal self attention. Same argument order, ``im2col_step`` accepted and ignored."""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from ... import _lib


def _check(value, shapes, starts, loc, attn):
    _lib.require_cuda(value, "value", torch.float32)
    _lib.require_cuda(loc, "sampling_locations", torch.float32)
    _lib.require_cuda(attn, "attention_weights", torch.float32)
    shapes = shapes.to(device=value.device, dtype=torch.int64).contiguous()
    starts = starts.to(device=value.device, dtype=torch.int64).contiguous()
    bs, nk, heads, dim = value.shape
    _, nq, h2, L, P, two = loc.shape
    if h2 != heads or two != 2 or tuple(attn.shape) != (bs, nq, heads, L, P) or shapes.shape[0] != L:
        raise RuntimeError("ms_deform_attn: inconsistent shapes value %s loc %s attn %s"
                           % (tuple(value.shape), tuple(loc.shape), tuple(attn.shape)))
    return shapes, starts, (bs, nk, heads, dim, nq, L, P)


class MultiScaleDeformableAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step=64):
        lib = _lib.load()
        value, loc, attn = value.contiguous(), sampling_locations.contiguous(), attention_weights.contiguous()
        shapes, starts, (bs, nk, heads, dim, nq, L, P) = _check(value, value_spatial_shapes,
                                                                value_level_start_index, loc, attn)
        out = torch.empty((bs, nq, heads * dim), dtype=torch.float32, device=value.device)
        with torch.cuda.device(value.device):
            rc = lib.dbev_ms_deform_attn_forward(_lib.ptr(value), _lib.ptr(shapes), _lib.ptr(starts), _lib.ptr(loc),
                                                 _lib.ptr(attn), bs, nk, heads, dim, nq, L, P, _lib.ptr(out),
                                                 _lib.stream_ptr(value.device))
        _lib.check(rc, "dbev_ms_deform_attn_forward")
        ctx.save_for_backward(value, shapes, starts, loc, attn)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        lib = _lib.load()
        value, shapes, starts, loc, attn = ctx.saved_tensors
        bs, nk, heads, dim = value.shape
        nq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
        grad_output = grad_output.contiguous().float()
        gv, gl, ga = torch.empty_like(value), torch.empty_like(loc), torch.empty_like(attn)
        with torch.cuda.device(value.device):
            rc = lib.dbev_ms_deform_attn_backward(_lib.ptr(value), _lib.ptr(shapes), _lib.ptr(starts), _lib.ptr(loc),
                                                  _lib.ptr(attn), _lib.ptr(grad_output), bs, nk, heads, dim, nq, L, P,
                                                  _lib.ptr(gv), _lib.ptr(gl), _lib.ptr(ga),
                                                  _lib.stream_ptr(value.device))
        _lib.check(rc, "dbev_ms_deform_attn_backward")
        return gv, None, None, gl, ga, None


MultiScaleDeformableAttnFunction_fp32 = MultiScaleDeformableAttnFunction


def multi_scale_deformable_attn(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                                attention_weights, im2col_step=64):
    return MultiScaleDeformableAttnFunction.apply(value, value_spatial_shapes, value_level_start_index,
                                                  sampling_locations, attention_weights, im2col_step)
