"""GPU parity: voxelization / dynamic scatter CUDA path (through the C-ABI shims) vs the C
oracle (pinned to the reference CPU build) and the committed reference fixtures.

Integer outputs (coors, voxel order, counts, point->voxel map) are bit-exact. Copied point
features and scatter-max are bit-exact; scatter-mean/sum are fp32 sums compared at rtol 1e-5
(the reference's own GPU path uses unordered float atomics, scatter_points.py:59-60).
"""
import os

import numpy as np
import pytest
import torch

import distill_bev_b200  # noqa: F401
from distill_bev_b200 import synthetic
from distill_bev_b200.plugin.ops import voxel as V
from oracle import voxel_oracle as vo

pytestmark = pytest.mark.gpu

PILLAR = ([0.2, 0.2, 8.0], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0])       # dynamic_centerpoint teacher
SPARSE = ([0.064, 0.064, 0.2], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0])   # lidarformer teacher (grid 1600x1600x40)
VOXEL01 = ([0.1, 0.1, 0.2], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0])


def _cloud(n, seed, nfeat=5, frac_out=0.05):
    pts = synthetic.make_lidar(1, n, seed=seed, num_features=nfeat)[0]
    k = int(n * frac_out)
    if k:
        pts[:k, 0] += 70.0
        pts[k:2 * k, 2] -= 9.0
    return pts


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_grid_size_matches_oracle(cuda):
    for vs, pcr in (PILLAR, SPARSE, VOXEL01, ([0.075, 0.075, 0.2], [-54.0, -54.0, -5.0, 54.0, 54.0, 3.0])):
        assert V.grid_size(vs, pcr) == vo.grid_size(vs, pcr)


def test_golden_fixture(cuda, golden_dir):
    g = np.load(os.path.join(golden_dir, "voxel_small.npz"))
    vs, pcr = g["voxel_size"].tolist(), g["coors_range"].tolist()
    pts = _t(g["points"], cuda)
    coors = V.voxelization(pts, vs, pcr, -1, -1)
    np.testing.assert_array_equal(coors.cpu().numpy(), g["dyn_coors"])
    for tag in "abc":
        mp, mv, m = [int(v) for v in g["hard_%s_cfg" % tag]]
        voxels, vc, num = V.voxelization(pts, vs, pcr, mp, mv)
        assert voxels.shape[0] == m
        np.testing.assert_array_equal(vc.cpu().numpy(), g["hard_%s_coors" % tag])
        np.testing.assert_array_equal(num.cpu().numpy(), g["hard_%s_num" % tag])
        np.testing.assert_array_equal(voxels.cpu().numpy(), g["hard_%s_voxels" % tag])
    feats = _t(g["sc_feats"], cuda)
    red, oc = V.dynamic_scatter(feats, coors, "max")
    np.testing.assert_array_equal(oc.cpu().numpy(), g["sc_out_coors"])
    np.testing.assert_array_equal(red.cpu().numpy(), g["sc_max"])
    r, oc2, cmap, cnt = V.voxel_layer.dynamic_point_to_voxel_forward(feats, coors, "sum")
    np.testing.assert_array_equal(cmap.cpu().numpy(), g["sc_map"])
    np.testing.assert_array_equal(cnt.cpu().numpy(), g["sc_count"])
    np.testing.assert_allclose(r.cpu().numpy(), g["sc_sum64"], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("n,cfg", [(30000, PILLAR), (240000, PILLAR), (30000, SPARSE), (240000, SPARSE),
                                   (1, PILLAR), (4097, VOXEL01)])
def test_dynamic_voxelize_vs_oracle(cuda, n, cfg):
    vs, pcr = cfg
    pts = _cloud(n, seed=n % 97)
    mod = V.Voxelization(vs, pcr, -1, -1)
    coors = mod(_t(pts, cuda))
    assert coors.dtype == torch.int32 and tuple(coors.shape) == (n, 3)
    np.testing.assert_array_equal(coors.cpu().numpy(), vo.dynamic_voxelize(pts, vs, pcr))


@pytest.mark.parametrize("n,cfg,mp,mv", [
    (30000, PILLAR, 20, 30000), (30000, PILLAR, 3, 2000), (240000, PILLAR, 20, 40000),
    (30000, SPARSE, 10, 90000), (240000, SPARSE, 10, 120000), (240000, SPARSE, 10, 5000),
    (50000, VOXEL01, 1, 100000), (7, PILLAR, 5, 3)])
def test_hard_voxelize_vs_oracle(cuda, n, cfg, mp, mv):
    vs, pcr = cfg
    pts = _cloud(n, seed=(n + mp) % 89, nfeat=5)
    pts[n // 2: n // 2 + n // 10, :3] = pts[: n // 10, :3]      # force multi-point voxels
    voxels, coors, num = V.voxelization(_t(pts, cuda), vs, pcr, mp, mv, True)
    ov, oc, ok = vo.hard_voxelize(pts, vs, pcr, mp, mv)
    assert voxels.shape[0] == ov.shape[0]
    np.testing.assert_array_equal(coors.cpu().numpy(), oc)       # voxel order = first appearance
    np.testing.assert_array_equal(num.cpu().numpy(), ok)
    np.testing.assert_array_equal(voxels.cpu().numpy(), ov)      # in-voxel order, zero padding
    # deterministic=False is allowed any order; ours returns the deterministic result
    v2, c2, n2 = V.voxelization(_t(pts, cuda), vs, pcr, mp, mv, False)
    assert torch.equal(v2, voxels) and torch.equal(c2, coors) and torch.equal(n2, num)


def test_voxelization_module_train_eval_switch(cuda):
    vs, pcr = PILLAR
    pts = _t(_cloud(20000, 3), cuda)
    mod = V.Voxelization(vs, pcr, 20, (500, 1500))
    mod.train()
    assert mod(pts)[0].shape[0] == 500
    mod.eval()
    assert mod(pts)[0].shape[0] == 1500
    assert "max_voxels=(500, 1500)" in repr(mod)
    assert mod.pcd_shape[0] == 1 and int(mod.grid_size[0]) == 512


def test_hard_voxelize_17_features_mvp(cuda):
    vs, pcr = SPARSE
    pts = _cloud(60000, 11, nfeat=17)
    voxels, coors, num = V.voxelization(_t(pts, cuda), vs, pcr, 10, 90000)
    ov, oc, ok = vo.hard_voxelize(pts, vs, pcr, 10, 90000)
    np.testing.assert_array_equal(coors.cpu().numpy(), oc)
    np.testing.assert_array_equal(voxels.cpu().numpy(), ov)


@pytest.mark.parametrize("reduce", ["max", "mean", "sum"])
@pytest.mark.parametrize("n,C", [(30000, 10), (240000, 64), (5, 3)])
def test_dynamic_scatter_vs_oracle(cuda, reduce, n, C):
    vs, pcr = PILLAR
    pts = _cloud(n, seed=n % 13 + 1)
    coors = vo.dynamic_voxelize(pts, vs, pcr)
    feats = np.random.RandomState(C).randn(n, C).astype(np.float32)
    ft = _t(feats, cuda).requires_grad_(True)
    ct = _t(coors, cuda)
    red, oc, cmap, cnt = V.voxel_layer.dynamic_point_to_voxel_forward(ft.detach(), ct, reduce)
    r0, o0, m0, c0 = vo.dynamic_scatter(feats, coors, reduce)
    np.testing.assert_array_equal(oc.cpu().numpy(), o0)          # lexicographic (z, y, x) order
    np.testing.assert_array_equal(cmap.cpu().numpy(), m0)
    np.testing.assert_array_equal(cnt.cpu().numpy(), c0)
    if reduce == "max":
        np.testing.assert_array_equal(red.cpu().numpy(), r0)
    else:
        np.testing.assert_allclose(red.cpu().numpy(), r0, rtol=1e-5, atol=1e-5)
    # autograd path of the module-level function
    out, _ = V.dynamic_scatter(ft, ct, reduce)
    gr = np.random.RandomState(1).randn(*out.shape).astype(np.float32)
    out.backward(_t(gr, cuda))
    g0 = vo.dynamic_scatter_backward(gr, feats, red.cpu().numpy(), m0, c0, reduce)
    if reduce == "mean":
        np.testing.assert_allclose(ft.grad.cpu().numpy(), g0, rtol=1e-6, atol=1e-7)
    else:
        np.testing.assert_array_equal(ft.grad.cpu().numpy(), g0)


def test_dynamic_scatter_module_batched_equals_per_sample_loop(cuda):
    """DynamicScatter.forward with [N, 4] coors: one launch == the reference's per-sample loop."""
    vs, pcr = PILLAR
    B = 4
    clouds = synthetic.make_lidar(B, 20000, seed=9)
    feats, coors = [], []
    for b, pts in enumerate(clouds):
        pts[:300, 0] += 80.0
        c = vo.dynamic_voxelize(pts, vs, pcr)
        coors.append(np.concatenate([np.full((c.shape[0], 1), b, np.int32), c], 1))
        feats.append(pts)
    feats, coors = np.concatenate(feats), np.concatenate(coors)
    for avg in (True, False):
        mod = V.DynamicScatter(vs, pcr, avg)
        out, oc = mod(_t(feats, cuda), _t(coors, cuda))
        exp_f, exp_c = [], []
        for b in range(B):
            sel = coors[:, 0] == b
            r, o, _, _ = vo.dynamic_scatter(feats[sel], coors[sel][:, 1:], "mean" if avg else "max")
            exp_f.append(r)
            exp_c.append(np.concatenate([np.full((o.shape[0], 1), b, np.int32), o], 1))
        np.testing.assert_array_equal(oc.cpu().numpy(), np.concatenate(exp_c))
        if avg:
            np.testing.assert_allclose(out.cpu().numpy(), np.concatenate(exp_f), rtol=1e-5, atol=1e-5)
        else:
            np.testing.assert_array_equal(out.cpu().numpy(), np.concatenate(exp_f))


def test_empty_and_all_invalid_inputs(cuda):
    vs, pcr = PILLAR
    empty = torch.zeros(0, 5, device=cuda)
    assert V.voxelization(empty, vs, pcr, -1, -1).shape == (0, 3)
    v, c, k = V.voxelization(empty, vs, pcr, 20, 100)
    assert v.shape == (0, 20, 5) and c.shape == (0, 3) and k.shape == (0,)
    far = torch.full((100, 5), 500.0, device=cuda)
    assert int((V.voxelization(far, vs, pcr, -1, -1) == -1).all())
    v, c, k = V.voxelization(far, vs, pcr, 20, 100)
    assert v.shape[0] == 0
    red, oc = V.dynamic_scatter(torch.rand(100, 4, device=cuda),
                                torch.full((100, 3), -1, dtype=torch.int32, device=cuda), "max")
    assert red.shape == (0, 4) and oc.shape == (0, 3)
    r = V.voxel_layer.dynamic_point_to_voxel_forward(torch.zeros(0, 4, device=cuda),
                                                     torch.zeros(0, 3, dtype=torch.int32, device=cuda), "mean")
    assert r[0].shape == (0, 4)
    with pytest.raises(RuntimeError, match="do not support reduce type"):
        V.dynamic_scatter(torch.rand(4, 4, device=cuda), torch.zeros(4, 3, dtype=torch.int32, device=cuda), "min")
