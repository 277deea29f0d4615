"""Student BEV encoder mirrors vs the fixture generated from the UNMODIFIED reference classes
(tools/make_golden_bev_encoder.py: ResNetForBEVDet resnet.py:12-62, BasicBlock res_block.py:10-99, FPN_LSS
lss_fpn.py:10-72). CPU part: state_dict keys, shapes and seeded initial values are identical (a reference
checkpoint loads with strict=True). GPU part: forward / loss / gradients on the fixture's input."""
import os

import numpy as np
import pytest
import torch

import distill_bev_b200 as dbev

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bev_encoder.npz")


def _checksum(t):
    t = t.detach().double().cpu()
    return np.array([float(t.sum()), float(t.abs().sum()),
                     float((t * torch.arange(1, t.numel() + 1, dtype=torch.float64).reshape(t.shape) % 7).sum())])


def _build():
    torch.manual_seed(0)
    backbone = dbev.ResNetForBEVDet(numC_input=128, num_channels=[128, 256, 512])
    neck = dbev.FPN_LSS(in_channels=640, out_channels=256)
    return backbone, neck


def test_state_dict_layout_and_seeded_init_match_the_reference_classes():
    ref = np.load(GOLDEN)
    backbone, neck = _build()
    keys, shapes, sd = [], [], {}
    for prefix, mod in (("backbone", backbone), ("neck", neck)):
        for k, v in mod.state_dict().items():
            keys.append(prefix + "." + k)
            shapes.append(",".join(map(str, v.shape)))
            sd[prefix + "." + k] = v
    assert keys == [str(k) for k in ref["keys"]]                # same names, same order
    assert shapes == [str(s) for s in ref["shapes"]]
    for k, v in sd.items():                                     # same seed -> same tensors (same construction order)
        want = ref["sum/" + k]
        got = _checksum(v.float()) if v.dtype != torch.long else np.array([float(v)])
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6, err_msg=k)


@pytest.mark.gpu
def test_forward_backward_match_the_reference_classes(cuda):
    ref = np.load(GOLDEN)
    backbone, neck = _build()
    backbone, neck = backbone.to(cuda).train(), neck.to(cuda).train()
    gen = torch.Generator().manual_seed(5)
    x = torch.relu(torch.randn(1, 128, 32, 32, generator=gen))
    g = torch.randn(1, 256, 32, 32, generator=gen) / (256 * 32 * 32) ** 0.5
    np.testing.assert_allclose(_checksum(x), ref["x_sum"], rtol=1e-6)
    np.testing.assert_allclose(_checksum(g), ref["g_sum"], rtol=1e-6, atol=1e-9)
    xin = x.to(cuda).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = neck(backbone(xin))
    loss = (y * g.to(cuda)).sum()
    loss.backward()
    want_y = torch.from_numpy(ref["y"]).to(cuda)
    # TF32 products through 20 layers vs the fp32 CPU reference: 5e-3 of the output range (cuDNN TF32 measures the same)
    assert float((y - want_y).abs().max()) <= 5e-3 * float(want_y.abs().max())
    assert abs(float(loss) - float(ref["loss"])) <= 1e-3 * float((want_y * g.to(cuda)).abs().sum())
    # the last layers see little ReLU-mask / BatchNorm amplification: the final conv's bias gradient is exact (it is
    # the sum of g), the last BatchNorm's gamma is a plain TF32 bar, its beta already carries flipped ReLU masks
    # (cuDNN TF32 itself is 4.5e-2 off fp32 there, tools/debug_encoder_grads.py)
    for name, mod, key, tol in (("neck.up2.4.bias", neck, "up2.4.bias", 1e-4), ("neck.up2.2.weight", neck, "up2.2.weight", 2e-2),
                                ("neck.up2.2.bias", neck, "up2.2.bias", 1.5e-1)):
        got = dict(mod.named_parameters())[key].grad.cpu().numpy()
        want = ref["grad/" + name]
        assert np.abs(got - want).max() <= tol * np.abs(want).max() + 1e-7, name
    # every parameter received a gradient of the reference's magnitude (early layers: TF32-noise-limited, see
    # test_bev_encoder_gpu.py::test_encoder_forward_backward_matches_torch_modules for the calibrated bar)
    for prefix, mod in (("backbone", backbone), ("neck", neck)):
        for k, p in mod.named_parameters():
            want = float(ref["gradmax/" + prefix + "." + k])
            assert p.grad is not None and 0.5 * want <= float(p.grad.abs().max()) <= 2.0 * want + 1e-12, (k, want)
    # BatchNorm running statistics after this one training forward, as torch updates them
    for prefix, mod in (("backbone", backbone), ("neck", neck)):
        for k, v in mod.state_dict().items():
            if k.endswith("running_mean") or k.endswith("running_var"):
                want = torch.from_numpy(ref["after/" + prefix + "." + k]).to(cuda)
                torch.testing.assert_close(v, want, rtol=5e-3, atol=5e-4, msg=lambda m, k=k: "%s: %s" % (k, m))
    xg = torch.from_numpy(ref["x_grad"]).to(cuda)
    cos = torch.nn.functional.cosine_similarity(xin.grad.flatten(), xg.flatten(), dim=0)
    assert float(cos) >= 0.98


ADAPT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "adaptation_3layer.npz")


def _adaptation_modules():
    from distill_bev_b200.plugin.distill import adaptation as A
    p = dict(adaptation_type=["upsample_3layer", "mlp"], teacher_adaptation_type="identity",
             student_adaptation_params=dict(kernel_size=1, stride=1, upsample_factor=4), student_channels=[128, 128],
             teacher_channels=[128, 128], spatial_mask=False)
    return A.build_adaptation_layers(p)[0]


def test_adaptation_state_dict_keys_match_the_reference_classes():
    """'upsample_3layer' = nn.Sequential(nn.Upsample, ThreeLayer) and 'mlp' = Mlp as BEVDetDistill.__init__ builds them
    (bevdet_distill.py:230-301): key-for-key the state_dict of the UNMODIFIED classes (tools/make_golden_bev_encoder.py)."""
    ref = np.load(ADAPT)
    layers = _adaptation_modules()
    assert list(layers[0].state_dict().keys()) == [str(k) for k in ref["keys"]]
    assert list(layers[1].state_dict().keys()) == [str(k) for k in ref["mlp_keys"]]
    for k, v in layers[0].state_dict().items():
        assert tuple(v.shape) == tuple(ref["sd/" + k].shape), k


@pytest.mark.gpu
def test_upsample_3layer_and_mlp_match_the_reference_classes(cuda):
    ref = np.load(ADAPT)
    layers = _adaptation_modules()
    sd = {str(k): torch.from_numpy(ref["sd/" + str(k)]) for k in ref["keys"]}
    for k in list(sd):                     # the fixture holds the state AFTER one training forward: reset the statistics
        if k.endswith("running_mean"):
            sd[k] = torch.zeros_like(sd[k])
        elif k.endswith("running_var"):
            sd[k] = torch.ones_like(sd[k])
        elif k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros_like(sd[k])
    up3 = layers[0]
    up3.load_state_dict(sd, strict=True)
    up3 = up3.to(cuda).train()
    x = torch.from_numpy(ref["x"]).to(cuda).requires_grad_(True)
    y = up3(x)
    want = torch.from_numpy(ref["y"]).to(cuda)
    assert float((y - want).abs().max()) <= 3e-3 * float(want.abs().max())          # three TF32 1x1 convs
    g = torch.randn(tuple(want.shape), generator=torch.Generator().manual_seed(10)).to(cuda)
    y.backward(g)
    xg = torch.from_numpy(ref["x_grad"]).to(cuda)
    assert float(torch.nn.functional.cosine_similarity(x.grad.flatten(), xg.flatten(), dim=0)) >= 0.999
    # TF32 rounding flips ReLU masks and three small-batch BatchNorms (800 pixels per channel) amplify it: single entries
    # move by up to ~1e-1 of the max (cuDNN TF32 does the same, test_bev_encoder_gpu.py), the direction is preserved
    assert float((x.grad - xg).abs().max()) <= 1.5e-1 * float(xg.abs().max())
    grads = dict(up3.named_parameters())
    for k, p in grads.items():
        w = torch.from_numpy(ref["grad/" + k]).to(cuda)
        if k.endswith(".bias") and ".conv" in k:
            # a conv bias in front of a BatchNorm has a mathematically ZERO gradient (BN subtracts the mean): both the
            # reference's value and ours are rounding noise - bound it against the weight gradient's scale instead
            scale = float(grads[k.replace(".bias", ".weight")].grad.abs().max())
            assert float(p.grad.abs().max()) <= 1e-3 * scale and float(w.abs().max()) <= 1e-3 * scale, k
            continue
        assert float((p.grad - w).abs().max()) <= 1.5e-1 * float(w.abs().max()) + 1e-6, k
        if w.numel() > 1000:
            assert float(torch.nn.functional.cosine_similarity(p.grad.flatten(), w.flatten(), dim=0)) >= 0.999, k
    for k, v in up3.state_dict().items():  # running statistics after the forward, like torch
        if k.endswith("running_mean") or k.endswith("running_var"):
            torch.testing.assert_close(v, torch.from_numpy(ref["sd/" + k]).to(cuda), rtol=5e-3, atol=5e-4)
    mlp = layers[1]
    mlp.load_state_dict({str(k): torch.from_numpy(ref["mlp_sd/" + str(k)]) for k in ref["mlp_keys"]}, strict=True)
    mlp = mlp.to(cuda)
    ym = mlp(torch.from_numpy(ref["mlp_x"]).to(cuda))
    wm = torch.from_numpy(ref["mlp_y"]).to(cuda)
    assert float((ym - wm).abs().max()) <= 3e-3 * float(wm.abs().max())


GOLDEN_BN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bev_encoder_bottleneck.npz")


def _build_bottleneck():
    torch.manual_seed(3)
    return dbev.ResNetForBEVDet(numC_input=128, num_layer=[2], num_channels=[512], stride=[2], block_type="BottleNeck")


def test_bottleneck_state_dict_layout_and_seeded_init_match_the_reference_class():
    """block_type='BottleNeck' against the fixture of the UNMODIFIED reference Bottleneck (res_block.py:102-311,
    tools/make_golden_bev_encoder.py --bottleneck): same keys, shapes and seeded tensors."""
    ref = np.load(GOLDEN_BN)
    net = _build_bottleneck()
    sd = net.state_dict()
    assert list(sd.keys()) == [str(k) for k in ref["keys"]]
    assert [",".join(map(str, v.shape)) for v in sd.values()] == [str(s) for s in ref["shapes"]]
    for k, v in sd.items():
        got = _checksum(v.float()) if v.dtype != torch.long else np.array([float(v)])
        np.testing.assert_allclose(got, ref["sum/" + k], rtol=1e-6, atol=1e-6, err_msg=k)


@pytest.mark.gpu
def test_bottleneck_forward_backward_match_the_reference_class(cuda):
    ref = np.load(GOLDEN_BN)
    net = _build_bottleneck().to(cuda).train()
    gen = torch.Generator().manual_seed(6)
    x = torch.relu(torch.randn(2, 128, 16, 16, generator=gen))
    g = torch.randn(2, 512, 8, 8, generator=gen) / (2 * 512 * 8 * 8) ** 0.5
    np.testing.assert_allclose(_checksum(x), ref["x_sum"], rtol=1e-6)
    np.testing.assert_allclose(_checksum(g), ref["g_sum"], rtol=1e-6, atol=1e-9)
    xin = x.to(cuda).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = net(xin)[0]
    loss = (y * g.to(cuda)).sum()
    loss.backward()
    want_y = torch.from_numpy(ref["y"]).to(cuda)
    assert float((y - want_y).abs().max()) <= 5e-3 * float(want_y.abs().max())       # TF32 products, six conv + BN layers
    assert abs(float(loss) - float(ref["loss"])) <= 1e-3 * float((want_y * g.to(cuda)).abs().sum())
    want_gx = torch.from_numpy(ref["x_grad"]).to(cuda)
    cos = float((xin.grad * want_gx).sum() / (xin.grad.norm() * want_gx.norm()))
    # TF32 rounding flips ReLU masks; with only 128 pixels per channel the batch statistics amplify that (the bar against
    # cuDNN TF32 is in test_bev_encoder_gpu.py::test_bottleneck_backbone_matches_torch_modules)
    assert cos >= 0.995 and float((xin.grad - want_gx).abs().max()) <= 1.5e-1 * float(want_gx.abs().max())
    for k, p in net.named_parameters():
        want = float(ref["gradmax/" + k])
        assert p.grad is not None and 0.5 * want <= float(p.grad.abs().max()) <= 2.0 * want + 1e-12, (k, want)
    last_beta = dict(net.named_parameters())["layers.0.1.bn3.bias"].grad.cpu().numpy()
    want = ref["grad/layers.0.1.bn3.bias"]
    # a beta gradient is the sum of g over the pixels whose output passed the ReLU: 128 pixels per channel here, so one
    # TF32-flipped mask bit moves it by ~1e-1 of the largest entry
    assert np.abs(last_beta - want).max() <= 1.5e-1 * np.abs(want).max() + 1e-7
    for k, v in net.state_dict().items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            np.testing.assert_allclose(v.cpu().numpy(), ref["after/" + k], rtol=2e-3, atol=2e-4, err_msg=k)
