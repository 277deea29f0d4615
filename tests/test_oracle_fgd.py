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
