"""Glue of the distillation step — ``BEVDetDistill.distill_loss`` (:1365-1409) and the feature-position
loop of ``forward_distill`` (:1456-1507): which student / teacher maps are paired, the stride
asserts, the `multi_scale_epoch` gate and the loss-key suffixing. Arithmetic lives in fgd.py /
affinity.py (CUDA kernels); this module is Python control flow only, like the reference's.
"""
import torch

from . import affinity as _aff
from . import fgd as _fgd


def distill_loss(distill_type, teacher_feat, student_feat, teacher_preds, student_preds, heatmaps,
                 gt_bboxes_3d, distill_params, train_cfg, index, epoch=0, spatial_adaptation=None,
                 channel_adaptation=None, affinity_mask=None):
    """:1365-1409 for the distill types on the §8 path ('fgd', 'affinity'); the ablation-only types
    ('all', 'foreground_background', 's2m2_*', 'gauss_focal_heatmap', 'non_local', 'linfengzhang')
    raise NotImplementedError like an unknown type does in the reference."""
    assert isinstance(teacher_feat, torch.Tensor) and isinstance(student_feat, torch.Tensor)
    if distill_type == "fgd":
        t_hm = [p[0]["heatmap"] for p in teacher_preds] if teacher_preds is not None else None
        s_hm = [p[0]["heatmap"] for p in student_preds] if student_preds is not None else None
        return _fgd.fgd_distill_loss(teacher_feat, student_feat, gt_bboxes_3d, distill_params, train_cfg,
                                     spatial_adaptation=spatial_adaptation, heatmaps=heatmaps,
                                     teacher_heatmaps=t_hm, student_heatmaps=s_hm, index=index,
                                     epoch=epoch, channel_adaptation=channel_adaptation)
    if distill_type == "affinity":
        if channel_adaptation is not None:
            student_feat = channel_adaptation(student_feat)
        if affinity_mask is None:     # tensor branch (:716-734): every cell takes part
            affinity_mask = torch.ones_like(student_feat[:, :1])
        w = distill_params["affinity_weights"]
        return _aff.affinity_distill_loss(teacher_feat, student_feat, affinity_mask,
                                          weight=w[index] if len(w) > 1 else w[0],
                                          criterion=distill_params["affinity_criterion"],
                                          split=int(distill_params.get("affinity_split", 1)))
    raise NotImplementedError(distill_type)


def forward_distill_positions(distill_type, distill_params, train_cfg, img_feats, lss_feat, bev_backbone_feats,
                              teacher_neck_feat, teacher_backbone_feats, canvas_feat, teacher_preds,
                              student_preds, heatmaps, gt_bboxes_3d, channel_wise_adaptations,
                              teacher_adaptations, spatial_wise_adaptations, epoch=0):
    """The position loop of forward_distill (:1456-1507). Teacher features are taken as given (the
    frozen teacher runs under no_grad in the caller). Returns the suffixed loss dict."""
    sp, tp = distill_params["student_feat_pos"], distill_params["teacher_feat_pos"]
    assert len(set(sp)) == len(sp) and len(set(tp)) == len(tp) and len(sp) == len(tp)
    if not isinstance(teacher_neck_feat, (list, tuple)):
        teacher_neck_feat = [teacher_neck_feat]
    assert all(not f.requires_grad for f in teacher_neck_feat)
    out = dict()
    for index, (s_pos, t_pos) in enumerate(zip(sp, tp)):
        if s_pos == "head":
            student_feat = img_feats[0]
        elif s_pos == "lss":
            student_feat = lss_feat
        elif s_pos.startswith("backbone"):
            if epoch < distill_params["multi_scale_epoch"]:
                continue
            layer = int(s_pos[-1])
            assert layer in range(3)
            student_feat = bev_backbone_feats[layer]
        else:
            raise NotImplementedError(s_pos)
        if t_pos == "head":
            teacher_feat = teacher_neck_feat[0]
        elif t_pos.startswith("backbone"):
            layer = int(t_pos[-1])
            assert layer in range(3)
            teacher_feat = teacher_backbone_feats[layer]
        elif t_pos == "canvas":
            teacher_feat = canvas_feat
        else:
            raise NotImplementedError(t_pos)
        cadapt, tadapt = channel_wise_adaptations[index], teacher_adaptations[index]
        assert teacher_feat.shape[0] == student_feat.shape[0]
        cs = getattr(cadapt, "stride", (1, 1))
        ts = getattr(tadapt, "stride", (1, 1))
        assert student_feat.shape[2] / cs[0] == teacher_feat.shape[2] / ts[0]
        assert student_feat.shape[3] / cs[1] == teacher_feat.shape[3] / ts[1]
        teacher_feat = tadapt(teacher_feat)                      # fgd_distill_loss :1003
        losses = distill_loss(distill_type, teacher_feat, student_feat, teacher_preds, student_preds,
                              heatmaps, gt_bboxes_3d, distill_params, train_cfg, index, epoch=epoch,
                              spatial_adaptation=spatial_wise_adaptations[index],
                              channel_adaptation=cadapt)
        for key, val in losses.items():
            out["%s_%s_%s" % (key, s_pos, t_pos)] = val
    return out
