"""DistillBEV feature-distillation loss on the B200 kernels (csrc/distill_loss.cu).

Mirrors the loss part of ``BEVDetDistill`` (mmdet3d/models/detectors/bevdet_distill.py):
  * ``foreground_scale_mask``  :755-843  (BEVFormer variant bevformer_distill.py:391-482)
  * ``add_fp_as_fg``           :846-970
  * ``fgd_distill_loss``       :973-1324 — everything after the adaptation layers
with the same ``distill_params`` keys and the same loss-dict keys
(kd_fg_feat_loss, kd_bg_feat_loss, kd_channel_loss, kd_spatial_loss,
kd_fp_bg_feat_loss). Masks are rasterised on the device (the reference builds
them with numpy/numba on the host for every sample and distill position).
"""
import ctypes

import torch

from ... import _lib

_SCALE = {None: 0, False: 0, "": 0, "combine_gt": 1, "separate_gt": 2, "bg_only": 3}
_FP_MODE = {"teacher": 0, "student": 1, "teacher_selected_student": 2,
            "teacher+teacher_selected_student": 3}
LOSS_KEYS = ("kd_fg_feat_loss", "kd_bg_feat_loss", "kd_fp_bg_feat_loss", "kd_channel_loss",
             "kd_spatial_loss")


def _box_tensor(b):
    t = getattr(b, "tensor", b)          # LiDARInstance3DBoxes or a plain tensor
    return torch.as_tensor(t, dtype=torch.float32)


class PackedBoxes(object):
    """Ground-truth boxes of a batch already on the device: ``boxes`` [capacity, D >= 7] fp32 with
    the samples back to back, ``offsets`` [B + 1] int32 (device), ``max_per_sample`` an upper bound
    of boxes in one sample (host int; sizes the rasteriser's shared memory). Accepted wherever
    ``gt_bboxes_3d`` is; lets a step be CUDA-graph captured: refill the two tensors in place."""

    def __init__(self, boxes, offsets, max_per_sample):
        _lib.require_cuda(boxes, "boxes", torch.float32)
        _lib.require_cuda(offsets, "offsets", torch.int32)
        if boxes.dim() != 2 or boxes.shape[1] < 7:
            raise RuntimeError("boxes must be [n, >=7], got %s" % (tuple(boxes.shape),))
        self.boxes, self.offsets, self.max_per_sample = boxes.contiguous(), offsets.contiguous(), int(max_per_sample)

    def __len__(self):
        return self.offsets.numel() - 1


def pack_boxes(gt_bboxes_3d, device):
    """list of per-sample boxes [M_b, >=7] -> (boxes [sum M, D] cuda, offsets [B+1] cuda int32, max M)."""
    if isinstance(gt_bboxes_3d, PackedBoxes):
        return gt_bboxes_3d.boxes, gt_bboxes_3d.offsets, gt_bboxes_3d.max_per_sample
    ts = [_box_tensor(b).reshape(-1, _box_tensor(b).shape[-1] if _box_tensor(b).numel() else 9)
          for b in gt_bboxes_3d]
    dim = max([t.shape[1] for t in ts if t.shape[0] > 0] + [7])
    ts = [t if t.shape[0] > 0 else t.new_zeros((0, dim)) for t in ts]
    counts = [t.shape[0] for t in ts]
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    boxes = torch.cat(ts, 0) if sum(counts) else torch.zeros((1, dim))
    if boxes.is_cuda:
        boxes = boxes.contiguous()
    else:
        boxes = _lib.h2d_async(boxes, device)
    return boxes, _lib.h2d_async(torch.tensor(offs, dtype=torch.int32), device), max(counts + [0])


def foreground_scale_mask(student_H, student_W, gt_bboxes_3d, grid_size, point_cloud_range,
                          voxel_size, device, transpose_mask=False, cell_center=False,
                          float_out_size_factor=False, return_counts=False):
    """-> foreground_mask, fg_scale_mask, bg_scale_mask, each [B, 1, H, W] fp32 on `device`."""
    lib = _lib.load()
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("foreground_scale_mask: device must be CUDA (no CPU path)")
    assert grid_size[0] == grid_size[1] and student_W == student_H
    if float_out_size_factor:
        osf = float(grid_size[0]) / student_W        # bevformer_distill.py:397
    else:
        assert grid_size[0] % student_W == 0
        osf = float(int(grid_size[0]) // student_W)  # bevdet_distill.py:764
    B = len(gt_bboxes_3d)
    boxes, offs, max_m = pack_boxes(gt_bboxes_3d, device)
    fg = torch.empty((B, 1, student_H, student_W), dtype=torch.float32, device=device)
    fg_scale = torch.empty_like(fg)
    fg_count = torch.empty((B,), dtype=torch.int32, device=device)
    vs = torch.tensor(voxel_size, dtype=torch.float32)
    pcr = torch.tensor(point_cloud_range, dtype=torch.float32)
    with torch.cuda.device(device):
        rc = lib.dbev_fgd_foreground_mask(
            _lib.ptr(boxes), boxes.shape[1], _lib.ptr(offs), max_m, B, student_H, student_W,
            float(vs[0]), float(vs[1]), osf, float(pcr[0]), float(pcr[1]), int(bool(cell_center)),
            int(bool(transpose_mask)), _lib.ptr(fg), _lib.ptr(fg_scale), _lib.ptr(fg_count),
            _lib.stream_ptr(device))
    _lib.check(rc, "dbev_fgd_foreground_mask")
    hw = float(student_H * student_W)
    bg_scale = (1.0 / (hw - fg_count.to(torch.float32))).view(B, 1, 1, 1).expand_as(fg)
    if return_counts:
        return fg, fg_scale, bg_scale, fg_count
    return fg, fg_scale, bg_scale


def heatmap_class_max(heatmaps, apply_clip_sigmoid=False):
    """[B, K, H, W] (tensor or list of per-task tensors) -> [B, 1, H, W] max over classes."""
    lib = _lib.load()
    if isinstance(heatmaps, (list, tuple)):
        heatmaps = torch.cat(list(heatmaps), dim=1)
    _lib.require_cuda(heatmaps, "heatmaps", torch.float32)
    heatmaps = heatmaps.contiguous()
    B, K, H, W = heatmaps.shape
    out = torch.empty((B, 1, H, W), dtype=torch.float32, device=heatmaps.device)
    with torch.cuda.device(heatmaps.device):
        rc = lib.dbev_heatmap_class_max(_lib.ptr(heatmaps), B, K, H, W, int(apply_clip_sigmoid),
                                        _lib.ptr(out), _lib.stream_ptr(heatmaps.device))
    _lib.check(rc, "dbev_heatmap_class_max")
    return out


def fp_dfs_scale(fp_mask):
    """fp_scale_mode 'dfs' (bevdet_distill.py:926-966): [B,1,H,W] FP mask -> per-component scale map."""
    lib = _lib.load()
    _lib.require_cuda(fp_mask, "fp_mask", torch.float32)
    fp_mask = fp_mask.contiguous()
    B, _, H, W = fp_mask.shape
    scale = torch.empty_like(fp_mask)
    ws = torch.empty(int(lib.dbev_fgd_fp_dfs_workspace_bytes(B, H, W)) // 8 + 1, dtype=torch.float64,
                     device=fp_mask.device)
    with torch.cuda.device(fp_mask.device):
        rc = lib.dbev_fgd_fp_dfs_scale(_lib.ptr(fp_mask), B, H, W, _lib.ptr(scale), _lib.ptr(ws), ws.numel() * 8,
                                       _lib.stream_ptr(fp_mask.device))
    _lib.check(rc, "dbev_fgd_fp_dfs_scale")
    return scale


def add_fp_as_fg(mode, fg_mask, gt_hm_max, teacher_hm_max, student_hm_max, thres, gt_thres=None,
                 return_counts=False, scale_mode="average"):
    """Class-max maps [B,1,S,S] -> fp_mask, fp_scale_mask [B,1,H,W], fp count [B] (float).
    scale_mode = distill_params['fp_scale_mode']: 'average' (1 / #fp of the sample) or 'dfs'."""
    if scale_mode not in ("average", "dfs"):
        raise NotImplementedError("fp_scale_mode=%r" % (scale_mode,))
    lib = _lib.load()
    if mode not in _FP_MODE:
        raise NotImplementedError(mode)
    if gt_thres is None:
        gt_thres = thres
    _lib.require_cuda(fg_mask, "fg_mask", torch.float32)
    B, _, R, _ = fg_mask.shape
    g, t = gt_hm_max.contiguous(), teacher_hm_max.contiguous()
    s = student_hm_max.contiguous() if student_hm_max is not None else None
    fp = torch.empty_like(fg_mask)
    cnt = torch.empty((B,), dtype=torch.int32, device=fg_mask.device)
    with torch.cuda.device(fg_mask.device):
        rc = lib.dbev_fgd_fp_mask(_lib.ptr(g), g.shape[2], _lib.ptr(t), t.shape[2], _lib.ptr(s),
                                  s.shape[2] if s is not None else 0, _lib.ptr(fg_mask.contiguous()),
                                  R, B, _FP_MODE[mode], float(thres), float(gt_thres), _lib.ptr(fp),
                                  _lib.ptr(cnt), _lib.stream_ptr(fg_mask.device))
    _lib.check(rc, "dbev_fgd_fp_mask")
    cntf = cnt.to(torch.float32)
    if scale_mode == "dfs":
        scale = fp_dfs_scale(fp)
    else:
        scale = torch.where(cntf > 0, 1.0 / cntf.clamp(min=1.0), torch.zeros_like(cntf)).view(B, 1, 1, 1) * fp
    if return_counts:
        return fp, scale, cntf, cnt
    return fp, scale, cntf


def make_config(B, C, H, W, distill_params, index=0, use_fp=None, epoch=0):
    """dbev_fgd_config from the reference's distill_params dict (…r50.py:50-92)."""
    p = distill_params

    def pick(key):
        v = p[key]
        return v[index] if len(v) > 1 else v[0]
    fp_mode = p.get("fp_as_foreground", "none")
    if isinstance(fp_mode, (list, tuple)):
        fp_mode = fp_mode[index] if len(fp_mode) > 1 else fp_mode[0]
    if use_fp is None:
        use_fp = fp_mode != "none" and epoch >= p.get("fp_epoch", 0)
    att = pick("spatial_attentions")
    if att not in ("teacher", "teacher_student"):
        raise NotImplementedError(att)
    # options that change the reference loss and are not implemented raise (they must never be ignored silently):
    # the non-empty background term (bevdet_distill.py:1137-1165, :1287-1291), the enlarged-foreground context
    # (:803-817), criteria other than the shipped MSE / L1 / L1 with reduction='none' (:997-999)
    if float(p.get("non_empty_weight", 0) or 0) != 0:
        raise NotImplementedError("distill_params['non_empty_weight'] != 0 (kd_non_empty_bg_feat_loss) is not implemented")
    if float(p.get("context_length", 0) or 0) > 0 and float(p.get("context_weight", 0) or 0) > 0:
        raise NotImplementedError("distill_params context_length / context_weight > 0 is not implemented")
    for key, want in (("feat_criterion", "MSELoss"), ("spatial_criterion", "L1Loss"), ("channel_criterion", "L1Loss")):
        crit = p.get(key)
        if crit is not None and (crit.get("type") != want or crit.get("reduction", "none") != "none"
                                 or float(crit.get("loss_weight", 1.0)) != 1.0):
            raise NotImplementedError("distill_params[%r] = %r: only dict(type=%r, reduction='none') is implemented"
                                      % (key, crit, want))
    if p.get("scale_mask") not in _SCALE:
        raise NotImplementedError(p.get("scale_mask"))
    if p.get("background_mask", "logical_not") not in ("logical_not", "1minus"):
        raise NotImplementedError(p.get("background_mask"))
    cfg = _lib.FgdConfig(
        B, C, H, W, float(p["spatial_t"]), float(p["channel_t"]), float(p["spatial_student_ratio"]),
        float(pick("fg_feat_loss_weights")), float(pick("bg_feat_loss_weights")),
        float(p.get("fp_weight", 0.0)), float(pick("channel_loss_weights")),
        float(pick("spatial_loss_weights")), 0 if att == "teacher" else 1,
        int(bool(p["spatial_mask"])), int(bool(p["channel_mask"])), _SCALE[p.get("scale_mask")],
        int(bool(use_fp)))
    return cfg, fp_mode


def _loss_forward(student, teacher, conv_w, conv_b, cfg, fg, fg_scale, fg_count, fp, fp_count):
    lib = _lib.load()
    _lib.require_cuda(student, "student_feat", torch.float32)
    _lib.require_cuda(teacher, "teacher_feat", torch.float32)
    if student.shape != teacher.shape:
        raise RuntimeError("student %s and teacher %s must have the same shape after adaptation"
                           % (tuple(student.shape), tuple(teacher.shape)))
    if cfg.spatial_mask and conv_w is None:
        raise RuntimeError("distill_params['spatial_mask'] is set: the spatial loss needs spatial_adaptation "
                           "(spatial_wise_adaptations[index], bevdet_distill.py:348-351, :1272-1278)")
    student, teacher = student.contiguous(), teacher.contiguous()
    dev = student.device
    if conv_w is None:
        conv_w = torch.zeros(9, device=dev)
        conv_b = torch.zeros(1, device=dev)
    cw = conv_w.detach().reshape(-1).contiguous().float()
    cb = conv_b.detach().reshape(-1).contiguous().float()
    nbytes = lib.dbev_fgd_state_bytes(ctypes.byref(cfg))
    state = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
    losses = torch.empty(5, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.dbev_fgd_loss_forward(
            ctypes.byref(cfg), _lib.ptr(student), _lib.ptr(teacher), _lib.ptr(fg.contiguous()),
            _lib.ptr(fg_scale.contiguous()), _lib.ptr(fg_count), _lib.ptr(fp), _lib.ptr(fp_count),
            _lib.ptr(cw), _lib.ptr(cb), _lib.ptr(state), nbytes, _lib.ptr(losses),
            _lib.stream_ptr(dev))
    _lib.check(rc, "dbev_fgd_loss_forward")
    return losses, state, student, teacher, cw, cb, (conv_w.shape, conv_b.shape)


def _loss_backward(cfg, state, student, teacher, cw, cb, grad_losses, conv_shape, channel_sum=False):
    lib = _lib.load()
    dev = student.device
    gl = grad_losses.contiguous().float()
    gs = torch.empty_like(student)
    gw = torch.empty(9, dtype=torch.float32, device=dev)
    gb = torch.empty(1, dtype=torch.float32, device=dev)
    gc = torch.empty(student.shape[1], dtype=torch.float32, device=dev) if channel_sum else None
    with torch.cuda.device(dev):
        rc = lib.dbev_fgd_loss_backward(
            ctypes.byref(cfg), _lib.ptr(student), _lib.ptr(teacher), _lib.ptr(cw), _lib.ptr(cb),
            _lib.ptr(state), state.numel() * 4, _lib.ptr(gl), _lib.ptr(gs), _lib.ptr(gw),
            _lib.ptr(gb), _lib.ptr(gc), _lib.stream_ptr(dev))
    _lib.check(rc, "dbev_fgd_loss_backward")
    return gs, gw.view(conv_shape[0]), gb.view(conv_shape[1]), gc


class _FGDLoss(torch.autograd.Function):

    @staticmethod
    def forward(ctx, student, teacher, conv_w, conv_b, cfg, fg, fg_scale, fg_count, fp, fp_count,
                teacher_ready=None):
        if teacher_ready is not None:
            torch.cuda.current_stream(student.device).wait_event(teacher_ready)
        losses, state, student, teacher, cw, cb, shp = _loss_forward(
            student, teacher, conv_w, conv_b, cfg, fg, fg_scale, fg_count, fp, fp_count)
        ctx.cfg, ctx.state, ctx.conv_shape, ctx.has_conv = cfg, state, shp, conv_w is not None
        ctx.save_for_backward(student, teacher, cw, cb)
        return losses

    @staticmethod
    def backward(ctx, grad_losses):
        student, teacher, cw, cb = ctx.saved_tensors
        gs, gw, gb, _ = _loss_backward(ctx.cfg, ctx.state, student, teacher, cw, cb, grad_losses,
                                       ctx.conv_shape)
        if not ctx.has_conv:      # no spatial_adaptation module: autograd accepts no gradient for a None input
            gw = gb = None
        return (gs, None, gw, gb, None, None, None, None, None, None, None)


class _AdaptFGDLoss(torch.autograd.Function):
    """channel_wise_adaptations[index] (1x1 conv, :1004) + the loss as ONE autograd node. Fused path (C_in % 128 == 0,
    C_out in 128 / 256 / 384 / 512): the adapted student exists only as tiles in tensor memory - the loss sums are
    taken in the GEMM epilogue (dbev_fgd_adapt_loss_forward), the backward recomputes the tiles and emits
    d loss / d adapted as channels-last rows for the tcgen05 input- / weight-gradient GEMMs, and the conv's bias
    gradient falls out as per-channel sums. Other shapes: adaptation GEMM -> NCHW map -> loss kernels."""

    @staticmethod
    def forward(ctx, x, weight, bias, teacher, conv_w, conv_b, cfg, fg, fg_scale, fg_count, fp, fp_count,
                teacher_ready=None):
        lib = _lib.load()
        ctx.fused = (x.shape[1] % 128 == 0 and weight.shape[0] % 128 == 0
                     and bool(lib.dbev_fgd_adapt_supported(ctypes.byref(cfg), int(x.shape[1]))))
        if ctx.fused:
            return _AdaptFGDLoss._forward_fused(ctx, lib, x, weight, bias, teacher, conv_w, conv_b, cfg, fg, fg_scale,
                                                fg_count, fp, fp_count, teacher_ready)
        from .adaptation import conv1x1_forward
        adapted, x = conv1x1_forward(x, weight, bias)
        if teacher_ready is not None:   # the adaptation conv above does not read the teacher
            torch.cuda.current_stream(x.device).wait_event(teacher_ready)
        losses, state, adapted, teacher, cw, cb, shp = _loss_forward(
            adapted, teacher, conv_w, conv_b, cfg, fg, fg_scale, fg_count, fp, fp_count)
        ctx.cfg, ctx.state, ctx.conv_shape, ctx.has_bias = cfg, state, shp, bias is not None
        ctx.has_conv = conv_w is not None
        ctx.save_for_backward(x, weight, adapted, teacher, cw, cb)
        return losses

    @staticmethod
    def _forward_fused(ctx, lib, x, weight, bias, teacher, conv_w, conv_b, cfg, fg, fg_scale, fg_count, fp, fp_count,
                       teacher_ready):
        from ..ops import conv_train as ct
        _lib.require_cuda(x, "student_feat", torch.float32)
        _lib.require_cuda(teacher, "teacher_feat", torch.float32)
        B, Cin, H, W = x.shape
        C = weight.shape[0]
        if tuple(teacher.shape) != (B, C, H, W):
            raise RuntimeError("student %s adapted to %d channels and teacher %s must have the same shape"
                               % (tuple(x.shape), C, tuple(teacher.shape)))
        if cfg.spatial_mask and conv_w is None:
            raise RuntimeError("distill_params['spatial_mask'] is set: the spatial loss needs spatial_adaptation "
                               "(spatial_wise_adaptations[index], bevdet_distill.py:348-351, :1272-1278)")
        dev = x.device
        x_cl = ct.as_nhwc(x.detach())
        if not x_cl.is_contiguous():                 # a channel slice: the GEMM's TMA map wants dense rows
            x_cl = x_cl.contiguous()
        teacher = teacher.detach().contiguous()
        w2 = weight.detach().reshape(C, Cin).contiguous()
        b1 = bias.detach().contiguous() if bias is not None else None
        has_conv = conv_w is not None
        if not has_conv:
            conv_w, conv_b = torch.zeros(9, device=dev), torch.zeros(1, device=dev)
        cw = conv_w.detach().reshape(-1).contiguous().float()
        cb = conv_b.detach().reshape(-1).contiguous().float()
        nbytes = lib.dbev_fgd_state_bytes(ctypes.byref(cfg))
        state = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
        losses = torch.empty(5, dtype=torch.float32, device=dev)
        if teacher_ready is not None:
            torch.cuda.current_stream(dev).wait_event(teacher_ready)
        with torch.cuda.device(dev):
            rc = lib.dbev_fgd_adapt_loss_forward(
                ctypes.byref(cfg), _lib.ptr(x_cl), Cin, _lib.ptr(w2), _lib.ptr(b1), _lib.ptr(teacher),
                _lib.ptr(fg.contiguous()), _lib.ptr(fg_scale.contiguous()), _lib.ptr(fg_count), _lib.ptr(fp),
                _lib.ptr(fp_count), _lib.ptr(cw), _lib.ptr(cb), _lib.ptr(state), nbytes, _lib.ptr(losses),
                _lib.stream_ptr(dev))
        _lib.check(rc, "dbev_fgd_adapt_loss_forward")
        ctx.cfg, ctx.state, ctx.has_bias, ctx.has_conv = cfg, state, bias is not None, has_conv
        ctx.conv_shape = (conv_w.shape, conv_b.shape)
        ctx.save_for_backward(x_cl, weight, w2, b1 if b1 is not None else cb, teacher, cw, cb)
        return losses

    @staticmethod
    def _backward_fused(ctx, grad_losses):
        from ..ops import conv_train as ct
        lib = _lib.load()
        x_cl, weight, w2, b1, teacher, cw, cb = ctx.saved_tensors
        dev = x_cl.device
        B, H, W, Cin = x_cl.shape
        C = w2.shape[0]
        want_bias = ctx.has_bias and ctx.needs_input_grad[2]
        gl = grad_losses.contiguous().float()
        g_cl = torch.empty((B, H, W, C), dtype=torch.float32, device=dev)
        gw = torch.empty(9, dtype=torch.float32, device=dev)
        gb = torch.empty(1, dtype=torch.float32, device=dev)
        gsum = torch.empty(C, dtype=torch.float32, device=dev) if want_bias else None
        with torch.cuda.device(dev):
            rc = lib.dbev_fgd_adapt_loss_backward(
                ctypes.byref(ctx.cfg), _lib.ptr(x_cl), Cin, _lib.ptr(w2), _lib.ptr(b1 if ctx.has_bias else None),
                _lib.ptr(teacher), _lib.ptr(cw), _lib.ptr(cb), _lib.ptr(ctx.state), ctx.state.numel() * 4, _lib.ptr(gl),
                _lib.ptr(g_cl), _lib.ptr(gw), _lib.ptr(gb), _lib.ptr(gsum), _lib.stream_ptr(dev))
        _lib.check(rc, "dbev_fgd_adapt_loss_backward")
        gx = gwt = None
        if ctx.needs_input_grad[0]:
            gx = ct.as_nchw(ct.conv_input_grad(g_cl, ct.pack_weights(weight, 1), Cin, 1, 1, 1, 0, (H, W)))
        if ctx.needs_input_grad[1]:
            gwt = ct.conv_weight_grad(x_cl, g_cl, 1, 1, 1, 0)
        gw, gb = (gw.view(ctx.conv_shape[0]), gb.view(ctx.conv_shape[1])) if ctx.has_conv else (None, None)
        return (gx, gwt, gsum, None, gw, gb, None, None, None, None, None, None, None)

    @staticmethod
    def backward(ctx, grad_losses):
        if ctx.fused:
            return _AdaptFGDLoss._backward_fused(ctx, grad_losses)
        x, weight, adapted, teacher, cw, cb = ctx.saved_tensors
        want_bias = ctx.has_bias and ctx.needs_input_grad[2]
        gs, gw, gb, gsum = _loss_backward(ctx.cfg, ctx.state, adapted, teacher, cw, cb, grad_losses,
                                          ctx.conv_shape, channel_sum=want_bias)
        from .adaptation import conv1x1_backward
        gx, gwt = conv1x1_backward(x, weight, gs, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        if not ctx.has_conv:
            gw = gb = None
        return (gx, gwt, gsum, None, gw, gb, None, None, None, None, None, None, None)


def fgd_loss_terms(student_feat, teacher_feat, cfg, fg, fg_scale, fg_count, fp=None, fp_count=None,
                   conv_weight=None, conv_bias=None, adapt_weight=None, adapt_bias=None, teacher_ready=None):
    """losses[5] tensor (order LOSS_KEYS), differentiable w.r.t. student_feat / conv. With
    ``adapt_weight`` [Ct, Cs, 1, 1] (+ ``adapt_bias``) the 1x1 adaptation conv is applied to
    ``student_feat`` [B, Cs, H, W] inside the same autograd node."""
    if adapt_weight is not None:
        return _AdaptFGDLoss.apply(student_feat, adapt_weight, adapt_bias, teacher_feat, conv_weight,
                                   conv_bias, cfg, fg, fg_scale, fg_count, fp, fp_count, teacher_ready)
    return _FGDLoss.apply(student_feat, teacher_feat, conv_weight, conv_bias, cfg, fg, fg_scale,
                          fg_count, fp, fp_count, teacher_ready)


def fgd_distill_loss(teacher_feat, student_feat, gt_bboxes_3d, distill_params, train_cfg,
                     spatial_adaptation=None, heatmaps=None, teacher_heatmaps=None,
                     student_heatmaps=None, index=0, epoch=0, channel_adaptation=None, teacher_ready=None):
    """Drop-in for the body of ``BEVDetDistill.fgd_distill_loss`` after the adaptation layers
    (:1006-1293): returns the same loss dict. ``train_cfg`` = pts_bbox_head.train_cfg
    (grid_size, point_cloud_range, voxel_size); ``spatial_adaptation`` = the
    ``spatial_wise_adaptations[index]`` Conv2d(1,1,3,padding=1); ``heatmaps`` /
    ``teacher_heatmaps`` (raw logits) / ``student_heatmaps`` (already sigmoid) are [B,K,h,w]
    tensors or per-task lists, needed only when fp_as_foreground is active.
    ``channel_adaptation`` = ``channel_wise_adaptations[index]`` (:1004), applied to ``student_feat``
    here as the reference does; a 1x1 conv is fused with the loss (tcgen05 forward, bias gradient
    from the loss backward), any other module is simply called.
    ``teacher_ready`` (optional torch.cuda.Event): recorded by the stream that produces
    ``teacher_feat``; the current stream waits for it only right before the first kernel that reads
    the teacher, so the masks and the student's adaptation conv overlap the frozen teacher's forward
    (the reference runs teacher and student serially on one stream)."""
    adapt_w = adapt_b = None
    if channel_adaptation is not None:
        conv = channel_adaptation
        if (isinstance(conv, torch.nn.Conv2d) and conv.kernel_size == (1, 1) and conv.stride == (1, 1)
                and conv.padding == (0, 0) and conv.groups == 1 and conv.dilation == (1, 1)
                and student_feat.is_cuda and student_feat.dtype == torch.float32 and conv.in_channels % 32 == 0
                and conv.out_channels in (128, 256, 384, 512)):        # the shapes csrc/adapt_gemm.cu covers
            adapt_w, adapt_b = conv.weight, conv.bias
        else:
            student_feat = channel_adaptation(student_feat)
    B, _, H, W = student_feat.shape
    C = adapt_w.shape[0] if adapt_w is not None else student_feat.shape[1]
    cfg, fp_mode = make_config(B, C, H, W, distill_params, index, epoch=epoch)
    if distill_params.get("foreground_mask", "gt") != "gt":
        raise NotImplementedError("foreground_mask=%r" % distill_params.get("foreground_mask"))
    fg, fg_scale, _, fg_count = foreground_scale_mask(
        H, W, gt_bboxes_3d, train_cfg["grid_size"], train_cfg["point_cloud_range"],
        train_cfg["voxel_size"], student_feat.device,
        transpose_mask=distill_params.get("transpose_mask", False), return_counts=True)
    fp = fp_count = None
    if cfg.use_fp:
        scale_mode = distill_params.get("fp_scale_mode", "average")
        g = heatmap_class_max(heatmaps)
        t = heatmap_class_max(teacher_heatmaps, apply_clip_sigmoid=True)
        s = heatmap_class_max(student_heatmaps) if student_heatmaps is not None else None
        fp, fp_scale, cntf, fp_count = add_fp_as_fg(fp_mode, fg, g, t, s, distill_params["output_threshold"],
                                                    distill_params.get("groundtruth_threshold"), return_counts=True,
                                                    scale_mode=scale_mode)
        if scale_mode == "dfs":
            # the loss kernels weight FP cells by fp * (1 / fp_count): hand them scale * fp_count instead of the
            # {0,1} mask (same non-zero pattern; the 'average' factor cancels)
            fp = fp_scale * cntf.view(-1, 1, 1, 1)
    cw = spatial_adaptation.weight if spatial_adaptation is not None else None
    cb = spatial_adaptation.bias if spatial_adaptation is not None else None
    losses = fgd_loss_terms(student_feat, teacher_feat, cfg, fg, fg_scale, fg_count, fp, fp_count, cw, cb,
                            adapt_w, adapt_b, teacher_ready=teacher_ready)
    out = {"kd_fg_feat_loss": losses[0], "kd_bg_feat_loss": losses[1]}
    if cfg.channel_mask:
        out["kd_channel_loss"] = losses[3]
    if cfg.spatial_mask:
        out["kd_spatial_loss"] = losses[4]
    if cfg.use_fp:
        out["kd_fp_bg_feat_loss"] = losses[2]
    # affinity branch of fgd_distill_loss (:1294-1321)
    mode = distill_params.get("affinity_mode", "none")
    if isinstance(mode, (list, tuple)):
        mode = mode[index] if len(mode) > 1 else mode[0]
    if mode != "none":
        from . import affinity as _aff
        if mode not in ("foreground", "foreground+fp", "attention"):
            raise NotImplementedError("affinity_mode=%r" % mode)
        if mode == "foreground+fp":
            assert fp_mode != "none"                                # :1297
        aw = distill_params["affinity_weights"]
        aw = aw[index] if len(aw) > 1 else aw[0]
        adapted = student_feat
        if adapt_w is not None:
            from .adaptation import conv1x1
            adapted = conv1x1(student_feat, adapt_w, adapt_b)
        if mode == "attention":
            # :1302-1308 - cells whose (detached) spatial attention passes a threshold, or is above the k-th largest
            # of its sample. Non-default mode: the attention map is rebuilt here with torch ops (:1084-1108).
            with torch.no_grad():
                n_cell = float(H * W)
                att = torch.softmax(teacher_feat.abs().mean(1).view(B, -1) / cfg.spatial_t, 1) * n_cell
                if cfg.spatial_att == 1:
                    s_att = torch.softmax(adapted.abs().mean(1).view(B, -1) / cfg.spatial_t, 1) * n_cell
                    att = (att + s_att * cfg.spatial_student_ratio) / (1 + cfg.spatial_student_ratio)
                if "affinity_attention_threshold" in distill_params:
                    sel = (att / n_cell) > distill_params["affinity_attention_threshold"]
                else:
                    kth = torch.topk(att, k=int(distill_params["affinity_attention_topk"]), dim=1)[0][:, -1:]
                    sel = att > kth
                sel = sel.view(B, 1, H, W).float()
            mask_a, mask_b = sel, None
        else:
            mask_a, mask_b = fg, (fp if (mode == "foreground+fp" and cfg.use_fp) else None)
        out.update(_aff.affinity_distill_loss(
            teacher_feat, adapted, mask_a, mask_b, weight=aw,
            criterion=distill_params.get("affinity_criterion", dict(type="SmoothL1Loss")),
            split=int(distill_params.get("affinity_split", 1))))
    return out
