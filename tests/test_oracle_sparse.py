"""CPU: oracle/spconv_oracle.py against tests/golden/sparse_small.npz — outputs of the reference's
own spconv extension (compiled unmodified, CPU branch) driven by its unmodified Python modules
(tools/make_golden_sparse.py). Rulebooks are bit-exact including order; features rtol 1e-5."""
import os

import numpy as np
import pytest

from oracle import spconv_oracle as so

ENC = {
    "lf": dict(in_channels=5, sparse_shape=[41, 48, 48], output_channels=128,
               encoder_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128)),
               encoder_paddings=((0, 0, 1), (0, 0, 1), (0, 0, [0, 1, 1]), (0, 0)),
               block_type="basicblock"),
    "sec": dict(in_channels=4, sparse_shape=[41, 32, 32], output_channels=128,
                encoder_channels=((16,), (32, 32, 32), (64, 64, 64), (64, 64, 64)),
                encoder_paddings=((1,), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1)),
                block_type="conv_module"),
}


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "sparse_small.npz"))


def _geom(g, name):
    v = g["rb_%s_geom" % name].tolist()
    return v[0:3], v[3:6], v[6:9], v[9:12], bool(v[12])


def test_rulebooks_bit_exact(g):
    for name in g["rb_names"]:
        shape, k, s, p, subm = _geom(g, name)
        out, pairs, num = so.get_indice_pairs(g["rb_%s_coors" % name], 2, shape, k, s, p, [1, 1, 1], subm)
        assert np.array_equal(out, g["rb_%s_outids" % name]), name
        assert np.array_equal(num, g["rb_%s_num" % name]), name
        assert np.array_equal(pairs, g["rb_%s_pairs" % name]), name


def test_indice_conv(g):
    for name in g["rb_names"]:
        y = so.indice_conv(g["rb_%s_feats" % name], g["rb_%s_w" % name], g["rb_%s_pairs" % name],
                           g["rb_%s_num" % name], len(g["rb_%s_outids" % name]))
        np.testing.assert_allclose(y, g["rb_%s_y" % name], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("tag", ["lf", "sec"])
def test_sparse_encoder(g, tag):
    cfg = ENC[tag]
    specs = so.fill_params(so.encoder_layer_specs(cfg["in_channels"], 16, cfg["output_channels"],
                                                  cfg["encoder_channels"], cfg["encoder_paddings"],
                                                  cfg["block_type"]), seed=11)
    y, _ = so.sparse_encoder(specs, g["enc_%s_feats" % tag], g["enc_%s_coors" % tag], 2,
                             cfg["sparse_shape"])
    ref = g["enc_%s_out" % tag]
    assert y.shape == ref.shape
    np.testing.assert_allclose(y, ref, rtol=1e-4, atol=1e-5 * np.abs(ref).max())


def test_hard_simple_vfe(g):
    np.testing.assert_allclose(so.hard_simple_vfe(g["vfe_voxels"], g["vfe_num"], 5), g["vfe_mean"],
                               rtol=1e-5, atol=1e-6)


def test_dynamic_voxel_encoder(g):
    v, c, shape = so.dynamic_voxel_encoder([g["dv_pts0"], g["dv_pts1"]], g["dv_range"], g["dv_voxel"])
    assert np.array_equal(c, g["dv_coors"]) and np.array_equal(shape, g["dv_shape"])
    np.testing.assert_allclose(v, g["dv_voxels"], rtol=1e-5, atol=1e-5)
    v, c, _ = so.dynamic_voxel_encoder([g["dvv_pts0"], g["dvv_pts1"]], g["dv_range"], g["dv_voxel"], True)
    assert np.array_equal(c, g["dvv_coors"])
    np.testing.assert_allclose(v, g["dvv_voxels"], rtol=1e-4, atol=1e-5)
