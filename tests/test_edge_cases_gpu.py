"""GPU: empty / degenerate inputs of the round-1 additions (the reference's own tests do not exist; these
are the edge cases SURVEY.md §8c lists: empty and ragged inputs, empty samples inside a batch)."""
import numpy as np
import pytest
import torch

import distill_bev_b200 as dbev
from distill_bev_b200.plugin.ops import spconv as sp

pytestmark = pytest.mark.gpu


def test_spconv_empty_and_single_voxel(cuda):
    empty = torch.zeros((0, 4), dtype=torch.int32, device=cuda)
    rb = sp.build_rulebook(empty, 2, [8, 16, 16], 3, 1, 1, 1, True)
    assert rb.n_out == 0
    rb = sp.build_rulebook(empty, 2, [8, 16, 16], 3, 2, 1, 1, False)
    assert rb.n_out == 0 and rb.out_shape == [4, 8, 8]
    w = torch.randn(3, 3, 3, 32, 32, device=cuda)
    y = sp.conv_table(torch.zeros((0, 32), device=cuda), w, rb.nbr, 0)
    assert y.shape == (0, 32)
    one = torch.tensor([[1, 3, 5, 7]], dtype=torch.int32, device=cuda)      # only sample 1 is populated
    rb = sp.build_rulebook(one, 2, [8, 16, 16], 3, 1, 1, 1, True)
    nbr = rb.nbr.cpu().numpy()
    assert nbr[13, 0] == 0 and (np.delete(nbr[:, 0], 13) == -1).all()
    x = torch.randn(1, 32, device=cuda)
    for impl in ("fma", "tc"):
        y = sp.conv_table(x, w, rb.nbr, 1, impl=impl)
        torch.testing.assert_close(y, x @ w[1, 1, 1], rtol=1e-4, atol=1e-4)
    rb2 = sp.build_rulebook(one, 2, [8, 16, 16], 3, 2, 1, 1, False)          # odd coords: 8 outputs
    assert rb2.n_out == 8 and (rb2.out_indices[:, 0] == 1).all()
    d = sp.dense_from_sparse(x, one, [8, 16, 16], 2)
    assert d.shape == (2, 32 * 8, 16, 16) and float(d[0].abs().sum()) == 0.0
    assert torch.equal(d[1, 3::8, 5, 7], x[0])


def test_dynamic_voxel_encoder_all_out_of_range_sample(cuda):
    enc = dbev.DynamicVoxelEncoder([-1, -1, -1, 1, 1, 1], [0.5, 0.5, 0.5])
    inside = torch.rand(50, 5, device=cuda) * 1.6 - 0.8
    outside = torch.rand(40, 5, device=cuda) + 5.0
    v, c, _ = enc([outside, inside])
    assert (c[:, 0] == 1).all() and v.shape[0] == c.shape[0] > 0
    v2, c2, _ = enc([outside])
    assert v2.shape[0] == 0 and c2.shape == (0, 4)


def test_affinity_no_selected_cells(cuda):
    t = torch.randn(2, 16, 8, 8, device=cuda)
    s = torch.randn(2, 16, 8, 8, device=cuda, requires_grad=True)
    out = dbev.affinity.affinity_distill_loss(t, s, torch.zeros(2, 1, 8, 8, device=cuda))
    assert float(out["kd_affinity_loss"]) == 0.0
    out["kd_affinity_loss"].backward()
    assert float(s.grad.abs().sum()) == 0.0


def test_center_targets_without_boxes(cuda):
    tasks = [dict(num_class=1, class_names=["car"]), dict(num_class=2, class_names=["bus", "trailer"])]
    cfg = dict(grid_size=[128, 128, 40], point_cloud_range=[-12.8, -12.8, -5.0, 12.8, 12.8, 3.0],
               voxel_size=[0.2, 0.2, 8.0], out_size_factor=4, dense_reg=1, gaussian_overlap=0.1, max_objs=10,
               min_radius=2)
    gen = dbev.CenterHeadTargets(tasks, cfg)
    hm, ab, ind, mk = gen.get_targets([torch.zeros((0, 9)), torch.zeros((0, 9))],
                                      [torch.zeros((0,), dtype=torch.int64)] * 2, device=cuda)
    assert float(gen.last_heatmap.abs().sum()) == 0.0 and int(mk[0].sum()) == 0 and hm[1].shape == (2, 2, 32, 32)


def test_hard_simple_vfe_empty(cuda):
    out = dbev.HardSimpleVFE(4)(torch.zeros((0, 10, 5), device=cuda), torch.zeros((0,), dtype=torch.int32, device=cuda))
    assert out.shape == (0, 4)


def test_shift_feature_far_motion_gives_zeros(cuda):
    n, v = 1, 6
    rots = torch.eye(3, device=cuda).expand(n, v, 3, 3).contiguous()
    t0 = torch.zeros(n, v, 3, device=cuda)
    t1 = t0.clone()
    t1[..., 0] = 500.0                               # adjacent frame 500 m away: every tap is out of bounds
    x = torch.randn(n, 8, 32, 32, device=cuda)
    out = dbev.shift_feature(x, [t0, t1], [rots, rots], [0.8, 0.8, 20.0], [-12.4, -12.4, 0.0])
    assert float(out.abs().sum()) == 0.0
