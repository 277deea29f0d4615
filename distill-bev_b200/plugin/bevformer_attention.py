"""BEVFormer student: spatial cross attention on the B200 kernels.

Mirrors (same constructor arguments, parameter names and forward signature):
  SpatialCrossAttention     mmdet3d/models/transformer_modules/spatial_cross_attention.py:30-174
  MSDeformableAttention3D   mmdet3d/models/transformer_modules/spatial_cross_attention.py:177-399

The reference re-batches the BEV queries per camera with Python double loops over (batch, camera) and nonzero() index
lists (:128-146), calls mmcv's ms_deform_attn extension (third party) and adds the results back with another double
loop (:160-167). Here the index lists are built once on the device (one read-back of the maximum list length, where
the reference synchronises once per camera), re-batching and the slot accumulation + count division are row kernels
(csrc/sca_rebatch.cu: gather / reduce, no atomics, camera order fixed), and the attention itself is
csrc/ms_deform_attn.cu. The Linear layers (sampling_offsets, attention_weights, value_proj, output_proj) stay torch
modules (plain library GEMMs). CUDA fp32 only.
"""
import math

import torch
from torch import nn

from .. import _lib
from .ops.ms_deform_attn import MultiScaleDeformableAttnFunction_fp32


def _rows_call(fn_name, src, index, scale, bs, cams, max_len, nq, C, out, strides=None):
    lib = _lib.load()
    with torch.cuda.device(src.device):
        if fn_name == "gather":
            cstride, bstride, qstride = strides
            rc = lib.dbev_sca_gather_rows(_lib.ptr(src), _lib.ptr(index), _lib.ptr(scale), bs, cams, max_len, nq, C, cstride,
                                          bstride, qstride, _lib.ptr(out), _lib.stream_ptr(src.device))
        else:
            rc = lib.dbev_sca_reduce_rows(_lib.ptr(src), _lib.ptr(index), _lib.ptr(scale), bs, cams, max_len, nq, C,
                                          _lib.ptr(out), _lib.stream_ptr(src.device))
    _lib.check(rc, "dbev_sca_%s_rows" % fn_name)
    return out


class _Rebatch(torch.autograd.Function):
    """queries_rebatch[b, cam, k] = query[b, idx[cam, k]] (zero padding); backward sums over the cameras."""

    @staticmethod
    def forward(ctx, query, idx, pos):
        bs, nq, C = query.shape
        cams, max_len = idx.shape
        q = query.contiguous()
        out = torch.empty((bs, cams, max_len, C), dtype=torch.float32, device=q.device)
        _rows_call("gather", q, idx, None, bs, cams, max_len, nq, C, out, strides=(0, nq * C, C))
        ctx.save_for_backward(idx, pos)
        ctx.dims = (bs, nq, C, cams, max_len)
        return out

    @staticmethod
    def backward(ctx, g):
        idx, pos = ctx.saved_tensors
        bs, nq, C, cams, max_len = ctx.dims
        out = torch.empty((bs, nq, C), dtype=torch.float32, device=g.device)
        _rows_call("reduce", g.contiguous(), pos, None, bs, cams, max_len, nq, C, out)
        return out, None, None


class _Slots(torch.autograd.Function):
    """slots[b, q] = sum over cameras of queries[b, cam, pos[cam, q]] / count[b, q]."""

    @staticmethod
    def forward(ctx, queries, idx, pos, inv_count):
        bs, cams, max_len, C = queries.shape
        nq = pos.shape[1]
        out = torch.empty((bs, nq, C), dtype=torch.float32, device=queries.device)
        _rows_call("reduce", queries.contiguous(), pos, inv_count, bs, cams, max_len, nq, C, out)
        ctx.save_for_backward(idx, inv_count)
        ctx.dims = (bs, nq, C, cams, max_len)
        return out

    @staticmethod
    def backward(ctx, g):
        idx, inv_count = ctx.saved_tensors
        bs, nq, C, cams, max_len = ctx.dims
        out = torch.empty((bs, cams, max_len, C), dtype=torch.float32, device=g.device)
        _rows_call("gather", g.contiguous(), idx, inv_count, bs, cams, max_len, nq, C, out, strides=(0, nq * C, C))
        return out, None, None, None


def camera_query_lists(bev_mask):
    """bev_mask [cams, bs, nq, D] -> (idx [cams, max_len] int32, pos [cams, nq] int32, inv_count [bs, nq] fp32, max_len).
    Like the reference (:130-133) the lists come from batch element 0's mask; the count (:169-171) from every
    element's own mask."""
    hit = bev_mask[:, 0].sum(-1) > 0                                   # [cams, nq]
    rank = torch.cumsum(hit.to(torch.int32), dim=1) - 1
    pos = torch.where(hit, rank, torch.full_like(rank, -1)).to(torch.int32).contiguous()
    max_len = int(hit.sum(1).max().item())                             # the one host read-back (the reference: one per camera)
    cams, nq = hit.shape
    idx = torch.full((cams, max(max_len, 1)), -1, dtype=torch.int32, device=hit.device)
    cam_ids, q_ids = torch.nonzero(hit, as_tuple=True)
    idx[cam_ids, rank[cam_ids, q_ids].long()] = q_ids.to(torch.int32)
    count = (bev_mask.sum(-1) > 0).permute(1, 2, 0).sum(-1)
    inv_count = (1.0 / torch.clamp(count, min=1.0)).to(torch.float32).contiguous()
    return idx[:, :max(max_len, 1)].contiguous(), pos, inv_count, max_len


class MSDeformableAttention3D(nn.Module):
    """spatial_cross_attention.py:177-399 (one z-anchor group per query; reference_points [bs, nq, D, 2])."""

    def __init__(self, embed_dims=256, num_heads=8, num_levels=4, num_points=8, im2col_step=64, dropout=0.1,
                 batch_first=True, norm_cfg=None, init_cfg=None):
        super().__init__()
        if embed_dims % num_heads != 0:
            raise ValueError("embed_dims must be divisible by num_heads, but got %d and %d" % (embed_dims, num_heads))
        self.norm_cfg, self.batch_first, self.output_proj = norm_cfg, batch_first, None
        self.im2col_step, self.embed_dims, self.num_levels = im2col_step, embed_dims, num_levels
        self.num_heads, self.num_points = num_heads, num_points
        self.sampling_offsets = nn.Linear(embed_dims, num_heads * num_levels * num_points * 2)
        self.attention_weights = nn.Linear(embed_dims, num_heads * num_levels * num_points)
        self.value_proj = nn.Linear(embed_dims, embed_dims)
        self.init_weights()

    def init_weights(self):
        nn.init.constant_(self.sampling_offsets.weight, 0.)
        thetas = torch.arange(self.num_heads, dtype=torch.float32) * (2.0 * math.pi / self.num_heads)
        grid_init = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid_init = (grid_init / grid_init.abs().max(-1, keepdim=True)[0]).view(self.num_heads, 1, 1, 2).repeat(
            1, self.num_levels, self.num_points, 1)
        for i in range(self.num_points):
            grid_init[:, :, i, :] *= i + 1
        self.sampling_offsets.bias.data = grid_init.view(-1)
        nn.init.constant_(self.attention_weights.weight, 0.)
        nn.init.constant_(self.attention_weights.bias, 0.)
        nn.init.xavier_uniform_(self.value_proj.weight)
        nn.init.constant_(self.value_proj.bias, 0.)

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_padding_mask=None,
                reference_points=None, spatial_shapes=None, level_start_index=None, **kwargs):
        if value is None:
            value = query
        if query_pos is not None:
            query = query + query_pos
        if not self.batch_first:
            query, value = query.permute(1, 0, 2), value.permute(1, 0, 2)
        bs, num_query, _ = query.shape
        _, num_value, _ = value.shape
        value = self.value_proj(value)
        if key_padding_mask is not None:
            value = value.masked_fill(key_padding_mask[..., None], 0.0)
        value = value.view(bs, num_value, self.num_heads, -1)
        sampling_offsets = self.sampling_offsets(query).view(bs, num_query, self.num_heads, self.num_levels, self.num_points, 2)
        attention_weights = self.attention_weights(query).view(bs, num_query, self.num_heads, self.num_levels * self.num_points)
        attention_weights = attention_weights.softmax(-1).view(bs, num_query, self.num_heads, self.num_levels, self.num_points)
        if reference_points.shape[-1] != 2:
            raise ValueError("Last dim of reference_points must be 2, but get %d instead." % reference_points.shape[-1])
        offset_normalizer = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
        _, _, num_Z_anchors, _ = reference_points.shape
        reference_points = reference_points[:, :, None, None, None, :, :]
        sampling_offsets = sampling_offsets / offset_normalizer[None, None, None, :, None, :]
        num_all_points = sampling_offsets.shape[4]
        assert num_all_points % num_Z_anchors == 0
        sampling_offsets = sampling_offsets.view(bs, num_query, self.num_heads, self.num_levels,
                                                 num_all_points // num_Z_anchors, num_Z_anchors, 2)
        sampling_locations = (reference_points + sampling_offsets).view(bs, num_query, self.num_heads, self.num_levels,
                                                                        num_all_points, 2)
        output = MultiScaleDeformableAttnFunction_fp32.apply(value, spatial_shapes, level_start_index, sampling_locations,
                                                             attention_weights, self.im2col_step)
        if not self.batch_first:
            output = output.permute(1, 0, 2)
        return output


class SpatialCrossAttention(nn.Module):
    """spatial_cross_attention.py:30-174."""

    def __init__(self, embed_dims=256, num_cams=6, pc_range=None, dropout=0.1, init_cfg=None, batch_first=False,
                 deformable_attention=dict(type='MSDeformableAttention3D', embed_dims=256, num_levels=4), **kwargs):
        super(SpatialCrossAttention, self).__init__()
        self.init_cfg = init_cfg
        self.dropout = nn.Dropout(dropout)
        self.pc_range = pc_range
        cfg = dict(deformable_attention)
        if cfg.pop("type", "MSDeformableAttention3D") != "MSDeformableAttention3D":
            raise NotImplementedError("deformable_attention type %r" % deformable_attention.get("type"))
        self.deformable_attention = MSDeformableAttention3D(**cfg)
        self.embed_dims, self.num_cams, self.batch_first = embed_dims, num_cams, batch_first
        self.output_proj = nn.Linear(embed_dims, embed_dims)
        nn.init.xavier_uniform_(self.output_proj.weight)
        nn.init.constant_(self.output_proj.bias, 0.)

    def forward(self, query, key, value, residual=None, query_pos=None, key_padding_mask=None, reference_points=None,
                spatial_shapes=None, reference_points_cam=None, bev_mask=None, level_start_index=None, flag='encoder',
                **kwargs):
        _lib.require_cuda(query, "query", torch.float32)
        if key is None:
            key = query
        if value is None:
            value = key
        if residual is not None:
            raise NotImplementedError("SpatialCrossAttention: residual must be None (the reference leaves `slots` undefined otherwise)")
        inp_residual = query
        if query_pos is not None:
            query = query + query_pos
        bs, num_query, _ = query.size()
        D = reference_points_cam.size(3)
        idx, pos, inv_count, max_len = camera_query_lists(bev_mask)
        L = idx.shape[1]
        queries_rebatch = _Rebatch.apply(query, idx, pos)                                        # [bs, cams, L, C]
        # reference points of the listed queries: [cams, bs, nq, D, 2] -> [bs, cams, L, D, 2] (no gradient)
        rp = reference_points_cam.detach().to(torch.float32).contiguous()
        rp_rebatch = torch.empty((bs, self.num_cams, L, D * 2), dtype=torch.float32, device=query.device)
        _rows_call("gather", rp, idx, None, bs, self.num_cams, L, num_query, D * 2, rp_rebatch,
                   strides=(bs * num_query * D * 2, num_query * D * 2, D * 2))
        num_cams, l, bs2, embed_dims = key.shape
        key = key.permute(2, 0, 1, 3).reshape(bs * self.num_cams, l, self.embed_dims)
        value = value.permute(2, 0, 1, 3).reshape(bs * self.num_cams, l, self.embed_dims)
        queries = self.deformable_attention(
            query=queries_rebatch.view(bs * self.num_cams, L, self.embed_dims), key=key, value=value,
            reference_points=rp_rebatch.view(bs * self.num_cams, L, D, 2), spatial_shapes=spatial_shapes,
            level_start_index=level_start_index).view(bs, self.num_cams, L, self.embed_dims)
        slots = _Slots.apply(queries, idx, pos, inv_count)
        slots = self.output_proj(slots)
        return self.dropout(slots) + inp_residual
