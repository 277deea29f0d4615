"""CPU: the C voxelization oracle against fixtures produced by the reference's own CPU
extension (tools/make_golden.py) and, when oracle/_ref is present, against that extension
directly on a nuScenes-sized cloud."""
import os

import numpy as np
import pytest

import distill_bev_b200  # noqa: F401
from distill_bev_b200 import synthetic
from oracle import voxel_oracle as vo


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "voxel_small.npz"))


def test_dynamic_voxelize_golden(g):
    coors = vo.dynamic_voxelize(g["points"], g["voxel_size"], g["coors_range"])
    np.testing.assert_array_equal(coors, g["dyn_coors"])
    assert (coors[:40] == -1).all()          # x == xmax is outside
    sel = coors[40:80]
    assert (sel[sel[:, 0] >= 0][:, 1] == 0).all() and (sel[:, 0] >= 0).any()   # y == ymin is cell 0


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_hard_voxelize_golden(g, tag):
    mp, mv, m = [int(v) for v in g["hard_%s_cfg" % tag]]
    voxels, coors, num = vo.hard_voxelize(g["points"], g["voxel_size"], g["coors_range"], mp, mv)
    assert voxels.shape[0] == m
    np.testing.assert_array_equal(coors, g["hard_%s_coors" % tag])
    np.testing.assert_array_equal(num, g["hard_%s_num" % tag])
    np.testing.assert_array_equal(voxels, g["hard_%s_voxels" % tag])


def test_dynamic_scatter_golden(g):
    red, oc, cmap, cnt = vo.dynamic_scatter(g["sc_feats"], g["dyn_coors"], "max")
    np.testing.assert_array_equal(oc, g["sc_out_coors"])
    np.testing.assert_array_equal(cmap, g["sc_map"])
    np.testing.assert_array_equal(cnt, g["sc_count"])
    np.testing.assert_array_equal(red, g["sc_max"])
    s, _, _, _ = vo.dynamic_scatter(g["sc_feats"], g["dyn_coors"], "sum")
    np.testing.assert_allclose(s, g["sc_sum64"], rtol=1e-5, atol=1e-5)
    mean, _, _, _ = vo.dynamic_scatter(g["sc_feats"], g["dyn_coors"], "mean")
    np.testing.assert_allclose(mean, g["sc_sum64"] / cnt[:, None], rtol=1e-5, atol=1e-5)


def test_against_reference_extension_when_present():
    ref = vo.load_reference_voxel_layer()
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    import torch
    pts = synthetic.make_lidar(1, 30000, seed=4)[0]
    pts[:500, 0] += 60.0  # push some points out of range
    for vs, pcr, mp, mv in [([0.2, 0.2, 8.0], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], 20, 30000),
                            ([0.064, 0.064, 0.2], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], 10, 9000)]:
        tp = torch.from_numpy(pts)
        tc = torch.zeros(pts.shape[0], 3, dtype=torch.int32)
        ref.dynamic_voxelize(tp, tc, vs, pcr, 3)
        np.testing.assert_array_equal(vo.dynamic_voxelize(pts, vs, pcr), tc.numpy())
        v = torch.zeros(mv, mp, 5)
        c = torch.zeros(mv, 3, dtype=torch.int32)
        k = torch.zeros(mv, dtype=torch.int32)
        m = ref.hard_voxelize(tp, v, c, k, vs, pcr, mp, mv, 3, True)
        ov, oc, ok = vo.hard_voxelize(pts, vs, pcr, mp, mv)
        assert ov.shape[0] == m
        np.testing.assert_array_equal(oc, c[:m].numpy())
        np.testing.assert_array_equal(ok, k[:m].numpy())
        np.testing.assert_array_equal(ov, v[:m].numpy())
