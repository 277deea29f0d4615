"""GPU parity: CenterHead targets kernel (csrc/center_targets.cu) vs tests/golden/center_targets.npz —
outputs of the UNMODIFIED CenterHead.get_targets (tools/make_golden_targets.py). ind / mask bit-exact,
heat maps and anno boxes to 1e-6 (exp / log / sin / cos of different math libraries)."""
import os

import numpy as np
import pytest
import torch

import distill_bev_b200 as dbev

pytestmark = pytest.mark.gpu

TASKS = [dict(num_class=1, class_names=["car"]), dict(num_class=2, class_names=["truck", "construction_vehicle"]),
         dict(num_class=2, class_names=["bus", "trailer"]), dict(num_class=1, class_names=["barrier"]),
         dict(num_class=2, class_names=["motorcycle", "bicycle"]),
         dict(num_class=2, class_names=["pedestrian", "traffic_cone"])]
CFG = dict(grid_size=[256, 256, 40], point_cloud_range=[-25.6, -25.6, -5.0, 25.6, 25.6, 3.0],
           voxel_size=[0.2, 0.2, 8.0], out_size_factor=4, dense_reg=1, gaussian_overlap=0.1, max_objs=40, min_radius=2)


def test_targets_match_reference(golden_dir, cuda):
    g = np.load(os.path.join(golden_dir, "center_targets.npz"))
    boxes, labels, o = [], [], 0
    for m in g["n"]:
        boxes.append(torch.from_numpy(g["boxes"][o:o + m].copy()))
        labels.append(torch.from_numpy(g["labels"][o:o + m].copy()))
        o += m
    gen = dbev.CenterHeadTargets(TASKS, CFG, norm_bbox=True)
    hm, ab, ind, mk = gen.get_targets(boxes, labels, device=cuda)
    for t in range(len(TASKS)):
        assert np.array_equal(mk[t].cpu().numpy(), g["mask%d" % t]), t
        assert np.array_equal(ind[t].cpu().numpy(), g["ind%d" % t]), t
        np.testing.assert_allclose(hm[t].cpu().numpy(), g["hm%d" % t], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(ab[t].cpu().numpy(), g["anno%d" % t], rtol=1e-5, atol=1e-6)
    assert hm[0].dtype == torch.float32 and ind[0].dtype == torch.int64 and mk[0].dtype == torch.uint8
    # the concatenated class heat map feeds add_fp_as_fg directly
    assert gen.last_heatmap.shape == (3, 10, 64, 64)


def test_nuscenes_size_runs(cuda):
    """configs[1] size: 8 samples x up to 60 boxes on a 128 x 128 map (grid 1024, out_size_factor 8)."""
    from distill_bev_b200 import synthetic
    cfg = dict(CFG, grid_size=[1024, 1024, 40], point_cloud_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0],
               voxel_size=[0.1, 0.1, 0.2], out_size_factor=8, max_objs=500)
    data = synthetic.make_gt_boxes(8, seed=2)
    gen = dbev.CenterHeadTargets(TASKS, cfg)
    hm, ab, ind, mk = gen.get_targets([torch.from_numpy(b) for b, _ in data], [torch.from_numpy(l) for _, l in data],
                                      device=cuda)
    assert gen.last_heatmap.shape == (8, 10, 128, 128)
    assert float(gen.last_heatmap.max()) == 1.0 and int(sum(m.sum() for m in mk)) > 0
