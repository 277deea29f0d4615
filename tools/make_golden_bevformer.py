"""Golden vectors for the BEVFormer-student distillation variants, produced by EXECUTING the
UNMODIFIED method bodies of ``BEVFormerDistill`` (mmdet3d/models/detectors/bevformer_distill.py):
foreground_scale_mask (:391-482, cell centre + float out_size_factor), add_fp_as_fg_bbox
(:573-647), fgd_distill_loss (:650-813), hs_distill_loss (:376-385), query_distill_loss (:364-374),
cut out with ``ast`` and run with the reference's own box_np_ops / LiDARPoints (tools/ref_import.py).

    python tools/make_golden_bevformer.py   ->  tests/golden/bevformer_small.npz
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ref_import  # noqa: E402
from make_golden import _Boxes, _fgd_self  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


class _PredBoxes(object):
    """LiDARInstance3DBoxes stand-in for the teacher predictions: `.tensor` and boolean indexing."""

    def __init__(self, t):
        self.tensor = t

    def __getitem__(self, idx):
        return _PredBoxes(self.tensor[idx])


def main():
    methods, AttrDict = ref_import.load_fgd_methods(
        names=("foreground_scale_mask", "add_fp_as_fg_bbox", "fgd_distill_loss", "hs_distill_loss",
               "query_distill_loss"),
        cls_name="BEVFormerDistill", relpath="mmdet3d/models/detectors/bevformer_distill.py")
    B, C, H = 3, 8, 50                      # 50 cells over a 128-voxel grid: out_size_factor 2.56
    grid, pc_range, voxel = [128, 128, 40], [-12.8, -12.8, -5.0, 12.8, 12.8, 3.0], [0.2, 0.2, 0.2]
    rng = np.random.RandomState(33)

    def rand_boxes(m):
        bx = np.zeros((m, 9), dtype=np.float32)
        bx[:, 0:2] = rng.uniform(-11, 11, (m, 2))
        bx[:, 2] = rng.uniform(-2, 0, m)
        bx[:, 3:5] = rng.uniform(1.0, 5.0, (m, 2))
        bx[:, 5] = rng.uniform(1, 3, m)
        bx[:, 6] = rng.uniform(-3.14, 3.14, m)
        return bx

    gt = [rand_boxes(m) for m in (5, 0, 9)]
    pred = [rand_boxes(m) for m in (12, 7, 10)]
    for b in range(B):                      # some predictions coincide with ground truth
        k = min(len(gt[b]), 3)
        pred[b][:k] = gt[b][:k]
    scores = [rng.uniform(0, 0.4, len(p)).astype(np.float32) for p in pred]
    g = torch.Generator().manual_seed(8)
    teacher = torch.relu(torch.randn(B, C, H, H, generator=g))
    student = torch.relu(torch.randn(B, C, H, H, generator=g))
    base = dict(spatial_t=0.5, spatial_student_ratio=1.0, channel_t=0.5,
                fg_feat_loss_weights=[6e-3], bg_feat_loss_weights=[4e-2], spatial_loss_weights=[2.5e-3],
                spatial_attentions=["teacher_student"], adaptation_type=["1x1conv"],
                feat_criterion=dict(type="MSELoss", reduction="none"),
                spatial_criterion=dict(type="L1Loss", reduction="none"),
                transpose_mask=False, foreground_mask="gt", background_mask="logical_not",
                scale_mask="combine_gt", spatial_mask=True, affinity_mode=["none"],
                output_threshold=0.1, groundtruth_threshold=None, fp_as_foreground=["teacher"],
                fp_weight=6e-2, fp_epoch=0, fp_scale_mode="average", context_length=0, context_weight=0,
                hs_feat_loss_weights=0.5, query_loss_weight=0.25,
                query_criterion=dict(type="MSELoss", reduction="mean"))
    variants = dict(recipe=dict(), nofp=dict(fp_as_foreground=["none"], spatial_attentions=["teacher"],
                                              scale_mask="separate_gt"))
    out = dict(teacher=teacher.numpy(), student=student.numpy(), grid=np.array(grid),
               pc_range=np.array(pc_range, np.float32), voxel=np.array(voxel, np.float32),
               n_gt=np.array([len(b) for b in gt]), gt=np.concatenate(gt),
               n_pred=np.array([len(b) for b in pred]), pred=np.concatenate(pred),
               scores=np.concatenate(scores))
    for name, over in variants.items():
        params = dict(base)
        params.update(over)
        me = _fgd_self(methods, AttrDict, params, C, grid, pc_range, voxel)
        me.no_bg = False
        st = student.clone().requires_grad_(True)
        gtb = [_Boxes(torch.from_numpy(b)) for b in gt]
        fg, fgs, bgs = me.foreground_scale_mask(H, H, gtb, 0, 0)
        teacher_preds = [(_PredBoxes(torch.from_numpy(p)), torch.from_numpy(s), None)
                         for p, s in zip(pred, scores)]
        losses = me.fgd_distill_loss(teacher.clone(), st, gtb, None, None, None, teacher_preds, None, 0)
        sum(losses.values()).backward()
        conv = me.spatial_wise_adaptations[0]
        out.update({name + "_fg": fg.numpy(), name + "_fg_scale": fgs.numpy(), name + "_bg_scale": bgs.numpy(),
                    name + "_params": json.dumps(params), name + "_loss_keys": json.dumps(sorted(losses)),
                    name + "_loss_vals": np.array([float(losses[k]) for k in sorted(losses)], np.float64),
                    name + "_grad_student": st.grad.numpy(),
                    name + "_conv_w": conv.weight.detach().numpy().reshape(3, 3),
                    name + "_conv_b": conv.bias.detach().numpy()})
        if params["fp_as_foreground"][0] != "none":
            fp, fps, cnt = me.add_fp_as_fg_bbox(H, H, "teacher", fg, teacher_preds, gtb)
            out.update({name + "_fp": fp.numpy(), name + "_fp_scale": fps.numpy(), name + "_fp_count": cnt.numpy()})
        print("bevformer fgd", name, {k: float(v) for k, v in losses.items()})
    # decoder-state and query-similarity losses
    me = _fgd_self(methods, AttrDict, dict(base), C, grid, pc_range, voxel)
    L, Q, D = 3, 20, C
    t_hs, s_hs = torch.randn(L, B, Q, D, generator=g), torch.randn(L, B, Q, D, generator=g)
    hs = me.hs_distill_loss(t_hs[-1].permute(0, 2, 1), s_hs[-1].permute(0, 2, 1).clone())
    t_q, s_q = torch.randn(Q, 2 * D, generator=g), torch.randn(Q, 2 * D, generator=g)
    ql = me.query_distill_loss(teacher, t_q, t_hs, student, s_q, s_hs)
    out.update(t_hs=t_hs.numpy(), s_hs=s_hs.numpy(), t_query=t_q.numpy(), s_query=s_q.numpy(),
               hs_loss=np.float64(float(hs["hs_feat_loss"])), query_loss=np.float64(float(ql["query_loss"])))
    print("hs", float(hs["hs_feat_loss"]), "query", float(ql["query_loss"]))
    path = os.path.join(GOLDEN, "bevformer_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KB")


if __name__ == "__main__":
    main()
