"""Golden vectors for fp_scale_mode='dfs' (bevdet_distill.py:926-966) from the UNMODIFIED reference methods
(ast-extracted by tools/ref_import.py, run here on CPU): add_fp_as_fg on small maps with a few FP blobs (the
reference's flood fill counts re-queued cells, so blob shapes matter), one fgd_distill_loss with the mode on, and
one fgd_distill_loss with affinity_mode='attention' (top-k of the spatial attention, :1302-1308).

    python tools/make_golden_fp_dfs.py      ->  tests/golden/fp_dfs.npz
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402
import ref_import  # noqa: E402


def main():
    methods, AttrDict = ref_import.load_fgd_methods()
    B, C, H = 2, 8, 32
    grid, pc_range, voxel = [256, 256, 40], [-12.8, -12.8, -5.0, 12.8, 12.8, 3.0], [0.1, 0.1, 0.2]
    g = torch.Generator().manual_seed(5)
    rng = np.random.RandomState(9)
    boxes = []
    for b in range(B):
        n = 3
        bx = np.zeros((n, 9), np.float32)
        bx[:, 0:2] = rng.uniform(-9, 9, (n, 2))
        bx[:, 2] = -1.0
        bx[:, 3:5] = rng.uniform(1.5, 4.0, (n, 2))
        bx[:, 5] = 1.5
        bx[:, 6] = rng.uniform(-3.1, 3.1, n)
        boxes.append(bx)
    # teacher heat-map logits: a handful of small blobs (plus, L, 2x3, diagonal pair, single cells) above threshold
    t_logit = torch.full((B, 2, H, H), -6.0)
    blobs = [[(3, 3), (3, 4), (3, 5), (2, 4), (4, 4)], [(10, 20), (11, 20), (12, 20), (12, 21)],
             [(20, 5), (20, 6), (20, 7), (21, 5), (21, 6), (21, 7)], [(27, 27)], [(28, 28)], [(15, 15), (15, 16)],
             [(6, 25), (7, 25), (7, 26), (8, 26), (8, 27)], [(0, 0), (0, 1), (1, 0), (1, 1)], [(31, 30), (31, 31)]]
    for i, blob in enumerate(blobs):
        for (y, x) in blob:
            t_logit[i % B, i % 2, y, x] = 3.0
    gt_hm = torch.zeros(B, 2, H, H)
    s_prob = torch.rand(B, 2, H, H, generator=g) * 0.05
    teacher = torch.randn(B, C, H, H, generator=g)
    student = torch.randn(B, C, H, H, generator=g)
    canvas = torch.zeros(B, 1, H, H)
    params = dict(spatial_t=0.5, spatial_student_ratio=1.0, channel_t=0.5,
                  fg_feat_loss_weights=[6e-3], bg_feat_loss_weights=[4e-2], channel_loss_weights=[0.25],
                  spatial_loss_weights=[2.5e-3], spatial_attentions=["teacher_student"],
                  feat_criterion=dict(type="MSELoss", reduction="none"),
                  spatial_criterion=dict(type="L1Loss", reduction="none"),
                  channel_criterion=dict(type="L1Loss", reduction="none"),
                  transpose_mask=False, foreground_mask="gt", background_mask="logical_not",
                  scale_mask="combine_gt", spatial_mask=True, channel_mask=False,
                  student_feat_pos=["head"], teacher_feat_pos=["head"], affinity_mode=["none"],
                  non_empty_weight=0, output_threshold=0.1, groundtruth_threshold=None,
                  fp_as_foreground=["teacher"], fp_weight=6e-2, fp_epoch=0, fp_scale_mode="dfs",
                  context_length=0, context_weight=0)
    me = mg._fgd_self(methods, AttrDict, params, C, grid, pc_range, voxel)
    bx = [mg._Boxes(torch.from_numpy(b)) for b in boxes]
    fg, fgs, bgs = me.foreground_scale_mask(H, H, bx, 0, 0)
    fp, fps, cnt = me.add_fp_as_fg("teacher", fg, [gt_hm.clone()], [[dict(heatmap=t_logit.clone())]],
                                   [[dict(heatmap=s_prob.clone())]])
    st = student.clone().requires_grad_(True)
    losses = me.fgd_distill_loss(teacher.clone(), st, bx, None, canvas, [gt_hm.clone()],
                                 [[dict(heatmap=t_logit.clone())]], [[dict(heatmap=s_prob.clone())]], 0)
    sum(losses.values()).backward()
    conv = me.spatial_wise_adaptations[0]
    out = dict(params=json.dumps(params), grid=np.array(grid), pc_range=np.array(pc_range, np.float32),
               voxel=np.array(voxel, np.float32), n_boxes=np.array([b.shape[0] for b in boxes]),
               boxes=np.concatenate(boxes), gt_hm=gt_hm.numpy(), teacher_logit=t_logit.numpy(),
               student_prob=s_prob.numpy(), teacher=teacher.numpy(), student=student.numpy(), fg=fg.numpy(),
               fp=fp.numpy(), fp_scale=fps.numpy(), fp_count=cnt.numpy(), loss_keys=json.dumps(sorted(losses)),
               loss_vals=np.array([float(losses[k]) for k in sorted(losses)], np.float64),
               grad_student=st.grad.numpy(), conv_w=conv.weight.detach().numpy().reshape(3, 3),
               conv_b=conv.bias.detach().numpy())
    # affinity_mode 'attention' (:1302-1308): cells above the k-th largest spatial attention of their sample
    p2 = dict(params, fp_as_foreground=["none"], fp_scale_mode="average", affinity_mode=["attention"],
              affinity_attention_topk=40, affinity_weights=[0.5], affinity_split=1,
              affinity_criterion=dict(type="SmoothL1Loss"))
    me2 = mg._fgd_self(methods, AttrDict, p2, C, grid, pc_range, voxel)
    st2 = student.clone().requires_grad_(True)
    losses2 = me2.fgd_distill_loss(teacher.clone(), st2, bx, None, canvas, [gt_hm.clone()],
                                   [[dict(heatmap=t_logit.clone())]], [[dict(heatmap=s_prob.clone())]], 0)
    sum(losses2.values()).backward()
    out.update(att_params=json.dumps(p2), att_loss_keys=json.dumps(sorted(losses2)),
               att_loss_vals=np.array([float(losses2[k]) for k in sorted(losses2)], np.float64),
               att_grad_student=st2.grad.numpy())
    print("attention affinity", {k: float(v) for k, v in losses2.items()})
    np.savez_compressed(os.path.join(mg.GOLDEN, "fp_dfs.npz"), **out)
    print("fp cells", cnt.numpy(), "distinct scales", np.unique(fps.numpy()).tolist(), {k: float(v) for k, v in losses.items()})


if __name__ == "__main__":
    main()
