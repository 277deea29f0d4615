"""CPU: the distillation-loss oracle (oracle/fgd_oracle.py) against fixtures produced by the
reference's own method bodies (tools/make_golden.py: foreground_scale_mask, add_fp_as_fg,
fgd_distill_loss, affinity_distill_loss executed unmodified)."""
import json
import os

import numpy as np
import pytest

from oracle import fgd_oracle as fo


def _sigmoid_clip(x, eps=1e-4):
    return np.clip(1.0 / (1.0 + np.exp(-x.astype(np.float64))), eps, 1 - eps).astype(np.float32)


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "fgd_small.npz"))


def _boxes(g):
    out, o = [], 0
    for n in g["n_boxes"]:
        out.append(g["boxes"][o:o + n])
        o += n
    return out


def oracle_params(p):
    return dict(spatial_t=p["spatial_t"], spatial_student_ratio=p["spatial_student_ratio"],
                channel_t=p["channel_t"], w_fg=p["fg_feat_loss_weights"][0],
                w_bg=p["bg_feat_loss_weights"][0], w_channel=p["channel_loss_weights"][0],
                w_spatial=p["spatial_loss_weights"][0], w_fp=p["fp_weight"],
                spatial_att=p["spatial_attentions"][0], spatial_mask=p["spatial_mask"],
                channel_mask=p["channel_mask"], scale_mask=p["scale_mask"],
                background_mask=p["background_mask"])


@pytest.mark.parametrize("name", ["recipe", "baseconfig", "separate"])
def test_masks_losses_and_grads(g, name):
    p = json.loads(str(g[name + "_params"]))
    H = g["teacher"].shape[2]
    fg, fgs, bgs = fo.foreground_scale_mask(H, H, _boxes(g), g["grid"], g["pc_range"], g["voxel"])
    np.testing.assert_array_equal(fg, g[name + "_fg"])
    np.testing.assert_array_equal(fgs, g[name + "_fg_scale"])
    np.testing.assert_allclose(bgs, g[name + "_bg_scale"], rtol=1e-7)
    assert fg[1].sum() == 0 and fg.sum() > 20          # sample 1 has no boxes
    kw = {}
    mode = p["fp_as_foreground"][0]
    if mode != "none":
        fp, fps, cnt = fo.add_fp_as_fg(mode, fg, g["gt_hm"], _sigmoid_clip(g["teacher_logit"]),
                                       g["student_prob"], p["output_threshold"], p["groundtruth_threshold"])
        np.testing.assert_array_equal(fp, g[name + "_fp"])
        np.testing.assert_allclose(fps, g[name + "_fp_scale"], rtol=1e-6)
        np.testing.assert_array_equal(cnt, g[name + "_fp_count"])
        assert cnt.sum() > 0
        kw = dict(fp=fp, fp_scale=fps, fp_count=cnt)
    res = fo.fgd_loss(g["teacher"], g["student"], fg, fgs, bgs, oracle_params(p),
                      conv_w=g[name + "_conv_w"], conv_b=float(g[name + "_conv_b"][0]),
                      want_grad=True, **kw)
    keys = json.loads(str(g[name + "_loss_keys"]))
    for k, v in zip(keys, g[name + "_loss_vals"]):
        assert abs(res[k] - v) <= 2e-5 * max(abs(v), 1e-3), (k, res[k], v)   # reference is fp32
    gs = g[name + "_grad_student"]
    np.testing.assert_allclose(res["grad_student"], gs, rtol=2e-4, atol=2e-6 * np.abs(gs).max())
    np.testing.assert_allclose(res["grad_conv_w"], g[name + "_grad_conv_w"], rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(res["grad_conv_b"], g[name + "_grad_conv_b"][0], rtol=2e-4, atol=1e-6)


def test_affinity(g):
    v = fo.affinity_loss([g["aff_t0"], g["aff_t1"]], [g["aff_s0"], g["aff_s1"]], 0.5)
    assert abs(v - float(g["aff_loss"])) < 1e-5 * abs(v)


# ---------------------------------------------------------------- fp_scale_mode 'dfs' (bevdet_distill.py:926-966)
@pytest.fixture(scope="module")
def gd(golden_dir):
    return np.load(os.path.join(golden_dir, "fp_dfs.npz"))      # tools/make_golden_fp_dfs.py


def test_fp_dfs_scale_pinned_to_reference(gd):
    """The literal FIFO walk and the layered-DP restatement both reproduce the reference's scale map bit for bit,
    including its repeated counting of re-queued cells (a 2x3 block scales by 1/9, not 1/6)."""
    p = json.loads(str(gd["params"]))
    fp, fps, cnt = fo.add_fp_as_fg("teacher", gd["fg"], gd["gt_hm"], _sigmoid_clip(gd["teacher_logit"]),
                                   gd["student_prob"], p["output_threshold"], p["groundtruth_threshold"],
                                   scale_mode="dfs")
    np.testing.assert_array_equal(fp, gd["fp"])
    np.testing.assert_array_equal(cnt, gd["fp_count"])
    np.testing.assert_array_equal(fps, gd["fp_scale"])
    np.testing.assert_array_equal(fo.fp_dfs_scale_literal(gd["fp"]), gd["fp_scale"])
    assert np.float32(1.0 / 9.0) in np.unique(gd["fp_scale"])


def test_fp_dfs_dp_equals_literal_on_random_blobs():
    rng = np.random.RandomState(4)
    for density in (0.15, 0.3, 0.45):
        fp = (rng.rand(3, 1, 12, 12) < density).astype(np.float32)
        np.testing.assert_array_equal(fo.fp_dfs_scale(fp), fo.fp_dfs_scale_literal(fp))
    # a 5 x 5 block: 1 / sum of binomial path counts; the literal walk needs 4421 pops for 25 cells
    fp = np.zeros((1, 1, 9, 9), np.float32)
    fp[0, 0, 2:7, 2:7] = 1
    lit, dp = fo.fp_dfs_scale_literal(fp), fo.fp_dfs_scale(fp)
    np.testing.assert_array_equal(dp, lit)
    assert dp.max() < 1.0 / 25


def test_fgd_loss_with_dfs_scale_pinned_to_reference(gd):
    p = json.loads(str(gd["params"]))
    H = gd["teacher"].shape[2]
    boxes, o = [], 0
    for n in gd["n_boxes"]:
        boxes.append(gd["boxes"][o:o + n])
        o += n
    fg, fgs, bgs = fo.foreground_scale_mask(H, H, boxes, gd["grid"], gd["pc_range"], gd["voxel"])
    np.testing.assert_array_equal(fg, gd["fg"])
    fp, fps, cnt = fo.add_fp_as_fg("teacher", fg, gd["gt_hm"], _sigmoid_clip(gd["teacher_logit"]), gd["student_prob"],
                                   p["output_threshold"], p["groundtruth_threshold"], scale_mode="dfs")
    res = fo.fgd_loss(gd["teacher"], gd["student"], fg, fgs, bgs, oracle_params(p), conv_w=gd["conv_w"],
                      conv_b=float(gd["conv_b"][0]), want_grad=True, fp=fp, fp_scale=fps, fp_count=cnt)
    for k, v in zip(json.loads(str(gd["loss_keys"])), gd["loss_vals"]):
        assert abs(res[k] - v) <= 2e-5 * max(abs(v), 1e-3), (k, res[k], v)
    gs = gd["grad_student"]
    np.testing.assert_allclose(res["grad_student"], gs, rtol=2e-4, atol=2e-6 * np.abs(gs).max())


def test_attention_affinity_pinned_to_reference(gd):
    """affinity_mode 'attention' (:1302-1308): rows = cells whose spatial attention is above the k-th largest of the
    sample; oracle restatement (attention :1084-1108 + affinity_loss) against the reference's own loss value."""
    p = json.loads(str(gd["att_params"]))
    t, s = gd["teacher"].astype(np.float64), gd["student"].astype(np.float64)
    B, C, H, W = t.shape

    def att(f):
        a = np.abs(f).mean(axis=1).reshape(B, -1) / p["spatial_t"]
        e = np.exp(a - a.max(axis=1, keepdims=True))
        return e / e.sum(axis=1, keepdims=True) * H * W
    r = p["spatial_student_ratio"]
    sa = (att(t) + att(s) * r) / (1 + r)
    kth = np.sort(sa, axis=1)[:, -p["affinity_attention_topk"]][:, None]
    sel = sa > kth
    assert (sel.sum(axis=1) == p["affinity_attention_topk"] - 1).all()
    t_rows = [t[b].reshape(C, -1).T[sel[b]] for b in range(B)]
    s_rows = [s[b].reshape(C, -1).T[sel[b]] for b in range(B)]
    got = fo.affinity_loss(t_rows, s_rows, p["affinity_weights"][0])
    want = dict(zip(json.loads(str(gd["att_loss_keys"])), gd["att_loss_vals"]))["kd_affinity_loss"]
    assert abs(got - want) <= 2e-5 * abs(want), (got, want)
