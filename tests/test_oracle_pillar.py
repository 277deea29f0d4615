"""CPU: pillar-path oracle vs the fixture produced by the unmodified reference classes."""
import os

import numpy as np

from oracle import pillar_oracle as po


def test_pillar_encode_and_scatter_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "pillar_small.npz"))
    vf, vc = po.pillar_encode(g["points"], g["coors"], g["weight"], g["bn_weight"], g["bn_bias"],
                              g["bn_mean"], g["bn_var"], g["bn_eps"], g["voxel_size"], g["coors_range"])
    np.testing.assert_array_equal(vc, g["voxel_coors"])
    np.testing.assert_allclose(vf, g["voxel_feats"], rtol=2e-5, atol=2e-5)
    canvas = po.pillar_scatter(vf, vc, 2, 64, 64)
    np.testing.assert_array_equal(np.stack(np.nonzero(canvas.sum(1))).astype(np.int32), g["canvas_nonzero"])
    assert abs(canvas.astype(np.float64).sum() - float(g["canvas_checksum"])) < 1e-3 * abs(float(g["canvas_checksum"]))
