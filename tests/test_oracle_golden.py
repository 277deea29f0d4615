"""CPU: the oracle (oracle/lss_oracle.py) against fixtures produced by the
reference's own Python code (tools/make_golden.py)."""
import json
import os

import numpy as np

from oracle import lss_oracle


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_frustum_matches_reference(golden_dir):
    g = _load(golden_dir, "lss_small.npz")
    grid = json.loads(str(g["grid"]))
    fr = lss_oracle.create_frustum(tuple(g["input_size"]), int(g["downsample"]), grid["dbound"])
    assert fr.shape == g["frustum"].shape
    np.testing.assert_array_equal(fr, g["frustum"])


def test_gen_dx_bx_matches_reference(golden_dir):
    g = _load(golden_dir, "lss_small.npz")
    grid = json.loads(str(g["grid"]))
    dx, bx, nx = lss_oracle.gen_dx_bx(grid["xbound"], grid["ybound"], grid["zbound"])
    np.testing.assert_array_equal(dx, g["dx"])
    np.testing.assert_array_equal(bx, g["bx"])
    np.testing.assert_array_equal(nx, g["nx"])


def test_geometry_matches_reference(golden_dir):
    g = _load(golden_dir, "lss_small.npz")
    geom = lss_oracle.get_geometry(g["frustum"], g["rots"], g["trans"], g["intrins"],
                                   g["post_rots"], g["post_trans"])
    # float32 3x3 inverses differ in the last bits between LAPACK and this restatement
    np.testing.assert_allclose(geom, g["geom"], rtol=2e-5, atol=2e-4)


def test_voxel_pooling_matches_reference_small(golden_dir):
    g = _load(golden_dir, "lss_small.npz")
    out = lss_oracle.voxel_pooling(g["geom"], g["x"], g["bx"], g["dx"], g["nx"])
    assert out.shape == g["out_cumsum"].shape
    # scatter-sum path of the reference: direct sums, tight
    np.testing.assert_allclose(out, g["out_accelerated"], rtol=1e-5, atol=1e-5)
    # cumsum path of the reference: fp32 prefix-sum error, north_star tolerance 1e-3
    np.testing.assert_allclose(out, g["out_cumsum"], rtol=1e-3, atol=1e-3)
    # empty cells agree exactly
    np.testing.assert_array_equal(out == 0, g["out_accelerated"] == 0)


def test_voxel_pooling_backward_matches_reference_small(golden_dir):
    g = _load(golden_dir, "lss_small.npz")
    C = g["x"].shape[-1]
    xg = lss_oracle.voxel_pooling_backward(g["geom"], g["out_weight"], C, g["bx"], g["dx"], g["nx"])
    np.testing.assert_array_equal(xg.reshape(g["x_grad"].shape), g["x_grad"])


def test_voxel_pooling_edge_cases(golden_dir):
    g = _load(golden_dir, "lss_edge.npz")
    out = lss_oracle.voxel_pooling(g["geom"], g["x"], g["bx"], g["dx"], g["nx"])
    np.testing.assert_allclose(out, g["out_cumsum"], rtol=1e-5, atol=1e-6)
    C = g["x"].shape[-1]
    xg = lss_oracle.voxel_pooling_backward(g["geom"], g["out_weight"], C, g["bx"], g["dx"], g["nx"])
    np.testing.assert_array_equal(xg.reshape(g["x_grad"].shape), g["x_grad"])
    # third sample lies entirely outside the grid
    assert np.all(out[2] == 0)
    idx, kept = lss_oracle.voxel_indices(g["geom"], g["bx"], g["dx"], g["nx"])
    assert kept.sum() == 10
    # trunc-toward-zero leak: x in (-5, -4) maps to cell 0 and is kept
    assert kept[1] and idx[1, 0] == 0


def test_bev_pool_matches_quickcumsum(golden_dir):
    g = _load(golden_dir, "quickcumsum.npz")
    B, D, H, W = int(g["B"]), int(g["D"]), int(g["H"]), int(g["W"])
    out = lss_oracle.bev_pool(g["feats"], g["coords"], B, D, H, W)
    np.testing.assert_allclose(out, g["dense"], rtol=1e-4, atol=1e-4)
    # interval restatement on the reference's own sorted order
    order = g["sort_index"]
    xs, cs = g["feats"][order], g["coords"][order]
    _, _, starts, lengths = lss_oracle.sorted_intervals(cs, B, D, H, W)
    assert starts.shape[0] == g["x_pooled"].shape[0]
    dense = lss_oracle.bev_pool_interval_forward(xs, cs, starts, lengths, B, D, H, W)
    np.testing.assert_allclose(dense.transpose(0, 4, 1, 2, 3), g["dense"], rtol=1e-4, atol=1e-4)
    # backward: gradient of sum(pooled * weight) w.r.t. the sorted rows
    og = np.zeros((B, D, H, W, xs.shape[1]), dtype=np.float32)
    gp = g["geom_pooled"]
    og[gp[:, 3], gp[:, 2], gp[:, 0], gp[:, 1]] = g["weight"]
    xg = lss_oracle.bev_pool_interval_backward(og, cs, starts, lengths, xs.shape[0])
    np.testing.assert_array_equal(xg, g["x_sorted_grad"])


def test_fullsize_report_pins_oracle(golden_dir):
    rep = json.load(open(os.path.join(golden_dir, "fullsize_report.json")))
    assert rep["idx_equal_on_kept"] is True
    assert rep["rel_oracle_vs_accelerated"] < 1e-5
    assert rep["rel_oracle_vs_cumsum"] < 1e-3


def test_lift_splat_matches_reference_small(golden_dir):
    g = _load(golden_dir, "lss_small.npz")
    B, N = g["rots"].shape[:2]
    out = lss_oracle.lift_splat(g["geom"], g["lift_depth"], g["lift_feat"], B, N, g["bx"], g["dx"], g["nx"])
    np.testing.assert_allclose(out, g["lift_out"], rtol=1e-5, atol=1e-5)
    dd, df = lss_oracle.lift_splat_backward(g["geom"], g["lift_depth"], g["lift_feat"], g["out_weight"],
                                            B, N, g["bx"], g["dx"], g["nx"])
    np.testing.assert_allclose(dd, g["lift_ddepth"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(df, g["lift_dfeat"], rtol=1e-4, atol=1e-5)
