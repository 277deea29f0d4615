"""GPU parity: HardSimpleVFE and DynamicVoxelEncoder (plain / virtual) vs the reference fixtures
(tests/golden/sparse_small.npz). Voxel coordinates and order bit-exact; means rtol 1e-5."""
import os

import numpy as np
import pytest
import torch

import distill_bev_b200 as dbev

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "sparse_small.npz"))


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_hard_simple_vfe(g, cuda):
    vfe = dbev.HardSimpleVFE(num_features=5)
    out = vfe(_t(g["vfe_voxels"], cuda), _t(g["vfe_num"], cuda), None)
    np.testing.assert_allclose(out.cpu().numpy(), g["vfe_mean"], rtol=1e-5, atol=1e-6)
    out4 = dbev.HardSimpleVFE(num_features=4)(_t(g["vfe_voxels"], cuda), _t(g["vfe_num"], cuda), None)
    np.testing.assert_allclose(out4.cpu().numpy(), g["vfe_mean"][:, :4], rtol=1e-5, atol=1e-6)


def test_dynamic_voxel_encoder(g, cuda):
    enc = dbev.DynamicVoxelEncoder(g["dv_range"].tolist(), g["dv_voxel"].tolist(), virtual=False)
    v, c, shape = enc([_t(g["dv_pts0"], cuda), _t(g["dv_pts1"], cuda)])
    assert c.dtype == torch.int64
    assert np.array_equal(c.cpu().numpy(), g["dv_coors"])
    assert np.array_equal(shape, g["dv_shape"])
    np.testing.assert_allclose(v.cpu().numpy(), g["dv_voxels"], rtol=1e-5, atol=1e-5)


def test_dynamic_voxel_encoder_virtual(g, cuda):
    enc = dbev.DynamicVoxelEncoder(g["dv_range"].tolist(), g["dv_voxel"].tolist(), virtual=True)
    v, c, _ = enc([_t(g["dvv_pts0"], cuda), _t(g["dvv_pts1"], cuda)])
    assert np.array_equal(c.cpu().numpy(), g["dvv_coors"])
    assert v.shape[1] == 23
    np.testing.assert_allclose(v.cpu().numpy(), g["dvv_voxels"], rtol=1e-4, atol=1e-5)


def test_cpu_tensor_raises():
    enc = dbev.DynamicVoxelEncoder([-1, -1, -1, 1, 1, 1], [0.5, 0.5, 0.5])
    with pytest.raises(RuntimeError):
        enc([torch.zeros((4, 5))])
