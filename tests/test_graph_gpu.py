"""GPU: a CUDA-graph captured step (distill_bev_b200.CapturedStep) replays to the same bits as the
eager calls, also after the static inputs were refilled with a new batch."""
import numpy as np
import pytest
import torch

import distill_bev_b200 as dbev
from distill_bev_b200 import synthetic
from distill_bev_b200.plugin.distill import fgd

pytestmark = pytest.mark.gpu


def _params():
    return dict(spatial_t=0.5, spatial_student_ratio=1.0, channel_t=0.5, fg_feat_loss_weights=[6e-3],
                bg_feat_loss_weights=[4e-2], channel_loss_weights=[0.25], spatial_loss_weights=[2.5e-3],
                spatial_attentions=["teacher_student"], transpose_mask=False, foreground_mask="gt",
                background_mask="logical_not", scale_mask="combine_gt", spatial_mask=True,
                channel_mask=False, output_threshold=0.1, groundtruth_threshold=None,
                fp_as_foreground=["teacher"], fp_weight=6e-2, fp_epoch=0, fp_scale_mode="average")


def test_captured_step_matches_eager(cuda):
    B, nf, ncam, Cs, Ct, H = 2, 2, 6, 64, 128, 128
    tc = dict(grid_size=[1024, 1024, 40], point_cloud_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0],
              voxel_size=[0.1, 0.1, 0.2])
    torch.manual_seed(0)
    vt = dbev.ViewTransformerLiftSplatShoot(grid_config=synthetic.NUSC_GRID, numC_input=32, numC_Trans=64).to(cuda)
    enc = dbev.DynamicPillarFeatureNet(in_channels=5, feat_channels=(64,), voxel_size=[0.2, 0.2, 8.0],
                                       point_cloud_range=tc["point_cloud_range"]).to(cuda).eval()
    scat = dbev.PointPillarsScatter(64, [512, 512], channels_last=True)
    adapt = dbev.Conv1x1Adaptation(Cs, Ct).to(cuda)
    spatial = torch.nn.Conv2d(1, 1, 3, padding=1).to(cuda)
    g = torch.Generator().manual_seed(1)
    depth = torch.randn(nf * ncam, 59, 16, 44, generator=g).softmax(1).to(cuda).requires_grad_(True)
    feat = torch.randn(nf * ncam, 64, 16, 44, generator=g).to(cuda).requires_grad_(True)
    bev_grad = torch.rand(nf, 64, 128, 128, generator=g).to(cuda)
    student = torch.relu(torch.randn(B, Cs, H, H, generator=g)).to(cuda).requires_grad_(True)
    teacher = torch.relu(torch.randn(B, Ct, H, H, generator=g)).to(cuda)
    t_logit = (torch.randn(B, 10, H, H, generator=g) * 1.5 - 3.0).to(cuda)

    def batch(seed):
        calib = [torch.from_numpy(a) for a in synthetic.make_calibration(nf, ncam, seed=seed)]
        pts = [torch.from_numpy(c) for c in synthetic.make_lidar(B, 20000, seed=seed)]
        boxes = [torch.from_numpy(b).float() for b, _ in synthetic.make_gt_boxes(B, seed=seed)]
        hm = torch.rand(B, 10, H, H, generator=torch.Generator().manual_seed(seed)) ** 12
        return calib, pts, boxes, hm

    def compute(calib, pts, boxes, hm):
        geom = vt.get_geometry(*calib)
        plan = vt.make_plan(geom, nf)
        bev = dbev.lift_splat(depth, feat, plan)
        bev.backward(bev_grad)
        with torch.no_grad():
            canvas = dbev.pillar_canvas(pts, enc, scat)
        losses = fgd.fgd_distill_loss(teacher, student, boxes, _params(), tc, channel_adaptation=adapt,
                                      spatial_adaptation=spatial, heatmaps=hm, teacher_heatmaps=t_logit, epoch=1)
        sum(losses.values()).backward()
        out = [torch.stack([losses[k] for k in sorted(losses)]), bev.detach(), canvas, depth.grad, feat.grad,
               student.grad, adapt.weight.grad, adapt.bias.grad, spatial.weight.grad]
        for p in (depth, feat, student):
            p.grad = None
        adapt.zero_grad(set_to_none=True)
        spatial.zero_grad(set_to_none=True)
        return out

    calib, pts, boxes, hm = batch(11)
    cap, cap_sample = 256, 128   # static box capacity: rows in the batch / boxes in one sample
    d_calib = [t.to(cuda) for t in calib]
    d_pts = [t.to(cuda) for t in pts]
    d_hm = hm.to(cuda)
    d_boxes = torch.zeros(cap, boxes[0].shape[1], device=cuda)
    d_offs = torch.zeros(B + 1, dtype=torch.int32, device=cuda)

    def fill(calib, pts, boxes, hm):
        for d, h in zip(d_calib, calib):
            d.copy_(h)
        for d, h in zip(d_pts, pts):
            d.copy_(h)
        d_hm.copy_(hm)
        allb = torch.cat(boxes, 0)
        assert allb.shape[0] <= cap
        d_boxes[:allb.shape[0]].copy_(allb)
        d_offs.copy_(torch.tensor(np.concatenate([[0], np.cumsum([b.shape[0] for b in boxes])]), dtype=torch.int32))

    fill(calib, pts, boxes, hm)
    packed = fgd.PackedBoxes(d_boxes, d_offs, cap_sample)
    step = dbev.CapturedStep(lambda: compute(d_calib, d_pts, packed, d_hm), warmup=2)
    for seed in (11, 12, 13):
        calib, pts, boxes, hm = batch(seed)
        fill(calib, pts, boxes, hm)
        got = [t.clone() for t in step.replay()]
        ref = compute([t.to(cuda) for t in calib], [t.to(cuda) for t in pts], boxes, hm.to(cuda))
        torch.cuda.synchronize()
        for i, (a, b) in enumerate(zip(got, ref)):
            assert torch.equal(a, b), (seed, i, float((a - b).abs().max()))
    assert float(got[0].abs().sum()) > 0 and float(got[2].abs().sum()) > 0
