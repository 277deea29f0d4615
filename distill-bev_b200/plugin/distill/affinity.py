"""Affinity distillation loss on the B200 kernels (csrc/affinity.cu).

Mirrors ``BEVDetDistill.affinity_distill_loss`` (list branch,
mmdet3d/models/detectors/bevdet_distill.py:735-748) together with the masked-cell gather of
``fgd_distill_loss`` that feeds it (:1294-1321): for every sample the cells selected by
``affinity_mask`` become rows [K_b, C] of teacher and student, and

    kd_affinity_loss = sum_b weight * mean(criterion(T_b T_b^T, S_b S_b^T))

with criterion = mmdet SmoothL1Loss (beta 1) / L1Loss / MSELoss, reduction 'mean'. The K x K
gram matrices are never written. ``affinity_split`` > 1 partitions every sample's rows at random
(torch.randperm drawn like the reference, :741-743, or caller-given permutations) and sums the per-part losses.
"""
import torch
from torch.autograd import Function

from ... import _lib

_KIND = {"SmoothL1Loss": 0, "L1Loss": 1, "MSELoss": 2}


def select_rows(mask_a, mask_b=None):
    """Cells with mask_a != 0 (or mask_b != 0) per sample, ascending cell order ->
    (row_cell [K_total] int32 cuda, row_offsets [B+1] int32 cuda, offsets as a host list).
    One device->host read of B+1 ints (the reference's boolean-mask indexing synchronises too)."""
    lib = _lib.load()
    _lib.require_cuda(mask_a, "affinity_mask", torch.float32)
    B = mask_a.shape[0]
    hw = mask_a[0].numel()
    mask_a = mask_a.contiguous()
    if mask_b is not None:
        _lib.require_cuda(mask_b, "affinity_mask (second)", torch.float32)
        mask_b = mask_b.contiguous()
    dev = mask_a.device
    row_cell = torch.empty((B * hw,), dtype=torch.int32, device=dev)
    row_offsets = torch.empty((B + 1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        wsb = lib.dbev_affinity_select_workspace_bytes(B, hw)
        ws = _lib.workspace(wsb, dev)
        rc = lib.dbev_affinity_select(_lib.ptr(mask_a), _lib.ptr(mask_b), B, hw, _lib.ptr(row_cell),
                                      _lib.ptr(row_offsets), _lib.ptr(ws), wsb, _lib.stream_ptr(dev))
    _lib.check(rc, "dbev_affinity_select")
    offs = row_offsets.tolist()
    return row_cell[:offs[-1]], row_offsets, offs


def gather_rows(feat, row_cell, row_offsets, k_total):
    """feat [B, C, H, W] -> rows [K_total, C] (the feat[c][mask] gather, :1317-1320)."""
    lib = _lib.load()
    _lib.require_cuda(feat, "feat", torch.float32)
    feat = feat.contiguous()
    B, C = feat.shape[0], feat.shape[1]
    hw = feat[0, 0].numel()
    rows = torch.empty((k_total, C), dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        rc = lib.dbev_affinity_gather_rows(_lib.ptr(feat), _lib.ptr(row_cell), _lib.ptr(row_offsets),
                                           B, C, hw, k_total, _lib.ptr(rows),
                                           _lib.stream_ptr(feat.device))
    _lib.check(rc, "dbev_affinity_gather_rows")
    return rows


class _AffinityRows(Function):
    """loss(t_rows, s_rows): gradient flows to the student rows only (teacher is detached)."""

    @staticmethod
    def forward(ctx, t_rows, s_rows, offs, kind, beta, weight):
        lib = _lib.load()
        dev = s_rows.device
        B, C = len(offs) - 1, s_rows.shape[1]
        hoffs = _lib.host_ints(offs)
        partial = torch.empty((lib.dbev_affinity_partial_floats(hoffs, B),), dtype=torch.float32,
                              device=dev)
        loss = torch.empty((1,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.dbev_affinity_forward(_lib.ptr(t_rows), _lib.ptr(s_rows), hoffs, B, C, kind,
                                           float(beta), float(weight), _lib.ptr(partial),
                                           _lib.ptr(loss), _lib.stream_ptr(dev))
        _lib.check(rc, "dbev_affinity_forward")
        ctx.save_for_backward(t_rows, s_rows)
        ctx.cfg = (offs, kind, float(beta), float(weight))
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_loss):
        lib = _lib.load()
        t_rows, s_rows = ctx.saved_tensors
        offs, kind, beta, weight = ctx.cfg
        dev = s_rows.device
        B, C = len(offs) - 1, s_rows.shape[1]
        g = grad_loss.to(torch.float32).reshape(1).contiguous()
        d_rows = torch.empty_like(s_rows)
        with torch.cuda.device(dev):
            rc = lib.dbev_affinity_backward(_lib.ptr(t_rows), _lib.ptr(s_rows), _lib.host_ints(offs), B,
                                            C, kind, beta, weight, _lib.ptr(g), _lib.ptr(d_rows),
                                            _lib.stream_ptr(dev))
        _lib.check(rc, "dbev_affinity_backward")
        return None, d_rows, None, None, None, None


class _GatherRows(Function):
    """Differentiable feat -> rows gather; backward scatters the row gradients into a zero map."""

    @staticmethod
    def forward(ctx, feat, row_cell, row_offsets, k_total):
        ctx.save_for_backward(row_cell, row_offsets)
        ctx.shape = tuple(feat.shape)
        return gather_rows(feat, row_cell, row_offsets, k_total)

    @staticmethod
    def backward(ctx, d_rows):
        lib = _lib.load()
        row_cell, row_offsets = ctx.saved_tensors
        B, C = ctx.shape[0], ctx.shape[1]
        hw = 1
        for s in ctx.shape[2:]:
            hw *= s
        d_rows = d_rows.contiguous()
        grad = torch.empty(ctx.shape, dtype=torch.float32, device=d_rows.device)
        with torch.cuda.device(d_rows.device):
            rc = lib.dbev_affinity_scatter_rows(_lib.ptr(d_rows), _lib.ptr(row_cell),
                                                _lib.ptr(row_offsets), B, C, hw, d_rows.shape[0],
                                                _lib.ptr(grad), _lib.stream_ptr(d_rows.device))
        _lib.check(rc, "dbev_affinity_scatter_rows")
        return grad, None, None, None


def affinity_distill_loss(teacher_feat, student_feat, affinity_mask, affinity_mask_2=None,
                          weight=1.0, criterion=dict(type="SmoothL1Loss"), split=1, perms=None):
    """teacher_feat / student_feat [B, C, H, W], affinity_mask [B, 1, H, W] (non-zero = selected;
    affinity_mask_2 is OR-ed in: the 'foreground+fp' mode, :1296-1299) -> dict(kd_affinity_loss).
    split > 1 (:738-747): the K rows of a sample are partitioned by a random permutation into
    perm[j::split], one gram pair per part, loss averaged over the parts. perms (one index tensor per
    sample) fixes the permutation; by default it is torch.randperm(K) drawn per sample from torch's CPU
    generator in sample order - the reference's own call, so equal seeds give equal partitions."""
    split = int(split)
    if split < 1:
        raise ValueError("affinity_split must be >= 1")
    cfg = dict(criterion)
    kind = _KIND[cfg.get("type", "SmoothL1Loss")]
    beta = float(cfg.get("beta", 1.0))
    weight = float(weight) * float(cfg.get("loss_weight", 1.0))
    if teacher_feat.shape != student_feat.shape:
        raise RuntimeError("teacher and (adapted) student features must have the same shape")
    if teacher_feat.shape[1] % 4 != 0:
        raise RuntimeError("channel count must be a multiple of 4")
    row_cell, row_offsets, offs = select_rows(affinity_mask.float(), None if affinity_mask_2 is None
                                              else affinity_mask_2.float())
    k_total = offs[-1]
    t_rows = gather_rows(teacher_feat.detach(), row_cell, row_offsets, k_total)
    s_rows = _GatherRows.apply(student_feat, row_cell, row_offsets, k_total)
    if split > 1:
        # every (sample, part) becomes one segment of the row list; mean per segment, weight / split
        order, seg = [], [0]
        for b in range(len(offs) - 1):
            k = offs[b + 1] - offs[b]
            perm = torch.randperm(k) if perms is None else torch.as_tensor(perms[b], dtype=torch.long).cpu()
            if perm.numel() != k:
                raise RuntimeError("perms[%d] has %d entries, sample has %d rows" % (b, perm.numel(), k))
            for j in range(split):
                part = perm[j::split] + offs[b]
                order.append(part)
                seg.append(seg[-1] + int(part.numel()))
        index = torch.cat(order).to(t_rows.device) if order else torch.zeros(0, dtype=torch.long, device=t_rows.device)
        t_rows, s_rows = t_rows.index_select(0, index).contiguous(), s_rows.index_select(0, index).contiguous()
        offs, weight = seg, weight / split
    loss = _AffinityRows.apply(t_rows, s_rows, offs, kind, beta, weight)
    return dict(kd_affinity_loss=loss)
