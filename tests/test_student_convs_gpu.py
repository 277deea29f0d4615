"""GPU parity: trainable convs of the student BEV encoder (plugin/student_convs.py) - forward on the tcgen05 conv
kernels (TF32 multiply, tolerance 2e-3 of the output range vs fp32 cuDNN), backward = aten convolution_backward on the
same tensors (identical to nn.Conv2d's own gradients: 1e-5)."""
import pytest
import torch
import torch.nn as nn

import distill_bev_b200 as dbev

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_reference():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("cin,cout,k,stride,pad,hw,bias", [(640, 512, 3, 1, 1, (32, 32), False), (512, 256, 3, 1, 1, (24, 40), False),
                                                           (128, 128, 3, 2, 1, (64, 64), False), (256, 512, 3, 2, 1, (32, 32), True),
                                                           (256, 256, 1, 1, 0, (40, 40), True), (128, 384, 3, 1, 1, (16, 16), False)])
def test_conv2d_tc_forward_backward(cuda, cin, cout, k, stride, pad, hw, bias):
    torch.manual_seed(cin + cout)
    ref = nn.Conv2d(cin, cout, k, stride, pad, bias=bias).to(cuda)
    ours = dbev.convert_convs(nn.Sequential(nn.Conv2d(cin, cout, k, stride, pad, bias=bias))).to(cuda)
    ours[0].load_state_dict(ref.state_dict())
    assert isinstance(ours[0], dbev.Conv2dTC)
    x = torch.randn(2, cin, *hw, device=cuda).contiguous(memory_format=torch.channels_last)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = ref(xa), ours(xb)
    assert yb.shape == ya.shape
    err = (ya - yb).abs().max().item()
    assert err <= 2e-3 * ya.abs().max().item(), err
    g = torch.randn_like(ya)
    ya.backward(g)
    yb.backward(g)
    for a, b in [(xa.grad, xb.grad), (ref.weight.grad, ours[0].weight.grad)] + ([(ref.bias.grad, ours[0].bias.grad)] if bias else []):
        torch.testing.assert_close(b, a, rtol=1e-4, atol=1e-5 * float(a.abs().max()))


def test_unsupported_shapes_fall_back_to_cudnn(cuda):
    conv = dbev.convert_convs(nn.Sequential(nn.Conv2d(3, 48, 7, 2, 3))).to(cuda)   # C_in = 3: not a tcgen05 shape
    x = torch.randn(1, 3, 32, 32, device=cuda)
    want = torch.nn.functional.conv2d(x, conv[0].weight, conv[0].bias, 2, 3)
    torch.testing.assert_close(conv(x), want)
    assert not dbev.conv2d_tc_supported(conv[0].weight, (2, 2), (3, 3))
