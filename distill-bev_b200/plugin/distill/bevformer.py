"""BEVFormer-student distillation losses on the B200 kernels.

Mirrors ``BEVFormerDistill`` (mmdet3d/models/detectors/bevformer_distill.py):
  * foreground_scale_mask   :391-482  — BEV cell CENTRE (+ vs*osf/2) and a FLOAT out_size_factor
    (grid/W, e.g. 512/200 = 2.56), unlike the BEVDet detector's corner + integer `//`
  * add_fp_as_fg_bbox       :573-647  — false-positive mask rasterised from the teacher's decoded
    boxes with score > output_threshold, minus the ground-truth boxes
  * fgd_distill_loss        :650-813  — fg / bg terms carry the spatial attention only; channel
    attention multiplies only the FP term; no channel loss
  * hs_distill_loss         :376-385, query_distill_loss :364-374
Same distill_params keys and loss-dict keys. Masks are rasterised on the device by the kernels of
csrc/distill_loss.cu (the reference: numpy + numba on the host per sample); the loss terms are the
same 3-read fused forward / 1-pass backward as the BEVDet path (plugin/distill/fgd.py).
"""
import torch

from ... import _lib
from . import fgd


def foreground_scale_mask(student_H, student_W, gt_bboxes_3d, train_cfg, device, transpose_mask=False,
                          return_counts=False):
    """-> foreground_mask, fg_scale_mask, bg_scale_mask [B,1,H,W] (:391-482)."""
    return fgd.foreground_scale_mask(student_H, student_W, gt_bboxes_3d, train_cfg["grid_size"],
                                     train_cfg["point_cloud_range"], train_cfg["voxel_size"], device,
                                     transpose_mask=transpose_mask, cell_center=True,
                                     float_out_size_factor=True, return_counts=return_counts)


def _pack_predictions(teacher_preds, thres, device):
    """(bboxes, scores, labels) per sample -> PackedBoxes; boxes with score <= thres are moved out
    of the grid instead of being compacted away (no device->host read of the kept count)."""
    boxes, counts = [], []
    for bboxes, scores, _ in teacher_preds:
        t = getattr(bboxes, "tensor", bboxes)
        t = torch.as_tensor(t, dtype=torch.float32).to(device).reshape(-1, t.shape[-1] if t.numel() else 9)
        s = torch.as_tensor(scores, dtype=torch.float32).to(device).reshape(-1)
        if t.shape[0]:
            t = t.clone()
            t[:, 0] = torch.where(s > thres, t[:, 0], torch.full_like(t[:, 0], 1e9))
        boxes.append(t)
        counts.append(t.shape[0])
    dim = max([b.shape[1] for b in boxes if b.shape[0] > 0] + [7])
    boxes = [b if b.shape[0] > 0 else torch.zeros((0, dim), device=device) for b in boxes]
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    allb = torch.cat(boxes, 0) if sum(counts) else torch.zeros((1, dim), device=device)
    return fgd.PackedBoxes(allb.contiguous(), _lib.h2d_async(torch.tensor(offs, dtype=torch.int32), device),
                           max(counts + [0]))


def add_fp_as_fg_bbox(student_H, student_W, mode, fg_mask, teacher_preds, gt_bboxes_3d, distill_params,
                      train_cfg, return_counts=False):
    """-> fp_masks, fp_scale_masks [B,1,H,W], fp count [B] (:573-647). As in the reference the two
    rasters of this function are NOT transposed to [H(y), W(x)] (reshape(1,1,H,W) of the x-major
    point list, :632), whatever distill_params['transpose_mask'] says."""
    if distill_params.get("fp_scale_mode", "average") != "average":
        raise NotImplementedError("fp_scale_mode=%r" % distill_params.get("fp_scale_mode"))
    thres = distill_params["output_threshold"]
    device = fg_mask.device
    pred = _pack_predictions(teacher_preds, float(thres), device)
    pred_mask = foreground_scale_mask(student_H, student_W, pred, train_cfg, device, transpose_mask=True)[0]
    gt_mask = foreground_scale_mask(student_H, student_W, gt_bboxes_3d, train_cfg, device,
                                    transpose_mask=True)[0]
    fp = ((pred_mask != 0) & (gt_mask == 0)).float()
    cnt = fp.sum(dim=(1, 2, 3))
    scale = torch.where(cnt > 0, 1.0 / cnt.clamp(min=1.0), torch.zeros_like(cnt)).view(-1, 1, 1, 1) * fp
    if return_counts:
        return fp, scale, cnt, cnt.to(torch.int32)
    return fp, scale, cnt


def fgd_distill_loss(teacher_feat, student_feat, gt_bboxes_3d, teacher_preds, distill_params, train_cfg,
                     spatial_adaptation=None, index=0, epoch=0, channel_adaptation=None, no_bg=False):
    """Drop-in for the body of ``BEVFormerDistill.fgd_distill_loss`` (:650-813) after the teacher
    adaptation: returns the same loss dict (kd_fg_feat_loss, kd_bg_feat_loss unless ``no_bg``,
    kd_spatial_loss, kd_fp_bg_feat_loss). ``teacher_preds`` = list of (bboxes, scores, labels)."""
    adapt_w = adapt_b = None
    atype = distill_params.get("adaptation_type", ["1x1conv"])
    atype = atype[index] if len(atype) > 1 else atype[0]
    if atype == "interpolate_1x1conv":
        student_feat = torch.nn.functional.interpolate(student_feat, teacher_feat.shape[-2:], mode="bilinear",
                                                       align_corners=True)       # :681-683
    if channel_adaptation is not None:
        conv = channel_adaptation
        if (isinstance(conv, torch.nn.Conv2d) and conv.kernel_size == (1, 1) and conv.stride == (1, 1)
                and conv.padding == (0, 0) and conv.groups == 1 and student_feat.is_cuda
                and student_feat.dtype == torch.float32 and conv.in_channels % 32 == 0
                and conv.out_channels in (128, 256, 384, 512)):
            adapt_w, adapt_b = conv.weight, conv.bias
        else:
            student_feat = channel_adaptation(student_feat)
    B, _, H, W = student_feat.shape
    C = adapt_w.shape[0] if adapt_w is not None else student_feat.shape[1]
    if distill_params.get("foreground_mask", "gt") != "gt":
        raise NotImplementedError("foreground_mask=%r" % distill_params.get("foreground_mask"))   # :697-698
    p = dict(distill_params)
    p["channel_mask"] = False                  # fg / bg: spatial attention only (:763-765)
    p.setdefault("channel_loss_weights", [0.0])
    cfg, fp_mode = fgd.make_config(B, C, H, W, p, index, epoch=epoch)
    fg, fg_scale, _, fg_count = foreground_scale_mask(
        H, W, gt_bboxes_3d, train_cfg, student_feat.device,
        transpose_mask=distill_params.get("transpose_mask", False), return_counts=True)
    fp = fp_count = None
    if cfg.use_fp:
        fp, _, _, fp_count = add_fp_as_fg_bbox(H, W, fp_mode, fg, teacher_preds, gt_bboxes_3d,
                                               distill_params, train_cfg, return_counts=True)
    cw = spatial_adaptation.weight if spatial_adaptation is not None else None
    cb = spatial_adaptation.bias if spatial_adaptation is not None else None
    losses = fgd.fgd_loss_terms(student_feat, teacher_feat, cfg, fg, fg_scale, fg_count, fp, fp_count,
                                cw, cb, adapt_w, adapt_b)
    out = {"kd_fg_feat_loss": losses[0]}
    if not no_bg:
        out["kd_bg_feat_loss"] = losses[1]
    if cfg.spatial_mask:
        out["kd_spatial_loss"] = losses[4]
    if cfg.use_fp:
        out["kd_fp_bg_feat_loss"] = losses[2]
    return out


def hs_distill_loss(teacher_feat, student_feat, distill_params):
    """:376-385 — summed squared error of the decoder states [B, C, Q] / B (mmdet MSELoss 'none')."""
    w = distill_params["hs_feat_loss_weights"]
    B = student_feat.shape[0]
    return {"hs_feat_loss": ((student_feat - teacher_feat) ** 2).sum() * w / B}


def query_distill_loss(teacher_feat, teacher_query, teacher_hs, student_feat, student_query, student_hs,
                       distill_params):
    """:364-374 — similarity-map losses between BEV features and object queries / decoder states."""
    crit = dict(distill_params["query_criterion"])
    fn = {"MSELoss": torch.nn.functional.mse_loss, "L1Loss": torch.nn.functional.l1_loss,
          "SmoothL1Loss": torch.nn.functional.smooth_l1_loss}[crit.get("type", "MSELoss")]
    red, lw = crit.get("reduction", "mean"), crit.get("loss_weight", 1.0)
    tf = teacher_feat.reshape(teacher_feat.shape[0], teacher_feat.shape[1], -1).permute(0, 2, 1)
    sf = student_feat.reshape(student_feat.shape[0], student_feat.shape[1], -1).permute(0, 2, 1)
    tq = (tf @ teacher_query[:, teacher_query.shape[1] // 2:].T).sum(dim=-1)
    sq = (sf @ student_query[:, student_query.shape[1] // 2:].T).sum(dim=-1)
    th = torch.einsum("bij,bjkl->bikl", tf, teacher_hs.permute(1, 3, 0, 2)).sum(dim=-1)
    sh = torch.einsum("bij,bjkl->bikl", sf, student_hs.permute(1, 3, 0, 2)).sum(dim=-1)
    loss = lw * fn(tq, sq, reduction=red) + lw * fn(th, sh, reduction=red)
    return {"query_loss": loss * distill_params["query_loss_weight"]}
