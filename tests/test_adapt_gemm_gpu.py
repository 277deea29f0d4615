"""GPU: tcgen05 TF32 1x1 adaptation conv vs torch fp32 (fp64-accumulated) reference.
TF32 keeps 10 mantissa bits: tolerance 2e-3 of the output scale (north_star loss bound 1e-3
applies to the loss, which averages these errors out)."""
import numpy as np
import pytest
import torch

import distill_bev_b200  # noqa: F401
from distill_bev_b200.plugin.distill.adaptation import Conv1x1Adaptation, conv1x1

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,Cin,Cout,H", [(2, 256, 384, 128), (1, 512, 256, 64), (1, 256, 128, 128),
                                          (1, 64, 128, 16), (2, 32, 512, 32), (1, 256, 384, 50)])
def test_conv1x1_matches_fp32_reference(cuda, B, Cin, Cout, H):
    g = torch.Generator().manual_seed(B * 100 + Cin + Cout)
    x = torch.randn(B, Cin, H, H, generator=g)
    w = torch.randn(Cout, Cin, 1, 1, generator=g) / Cin ** 0.5
    b = torch.randn(Cout, generator=g)
    ref = torch.nn.functional.conv2d(x.double(), w.double(), b.double()).float()
    y = conv1x1(x.to(cuda), w.to(cuda), b.to(cuda))
    torch.cuda.synchronize()
    err = (y.cpu() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * scale, (err, scale)
    # channels_last input (no transpose kernel) gives the same bits
    y_cl = conv1x1(x.to(cuda).contiguous(memory_format=torch.channels_last), w.to(cuda), b.to(cuda))
    assert y_cl.is_contiguous() and torch.equal(y_cl, y)
    # exactness of the plumbing: a 0/1 weight matrix selects channels exactly in TF32
    sel = torch.zeros(Cout, Cin, 1, 1)
    for n in range(Cout):
        sel[n, (n * 7) % Cin] = 1.0
    xq = (x * 64).round() / 64                      # values exactly representable in TF32
    y2 = conv1x1(xq.to(cuda), sel.to(cuda), None).cpu()
    assert torch.equal(y2, xq[:, [(n * 7) % Cin for n in range(Cout)]])


def test_conv1x1_module_grads(cuda):
    torch.manual_seed(0)
    m = Conv1x1Adaptation(256, 384).to(cuda)
    ref = torch.nn.Conv2d(256, 384, 1).to(cuda)
    ref.load_state_dict(m.state_dict())
    x = torch.randn(2, 256, 32, 32, device=cuda, requires_grad=True)
    x2 = x.detach().clone().requires_grad_(True)
    gy = torch.randn(2, 384, 32, 32, device=cuda)
    m(x).backward(gy)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        ref(x2).backward(gy)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    torch.testing.assert_close(x.grad, x2.grad, rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(m.weight.grad, ref.weight.grad, rtol=2e-2, atol=5e-2)
    torch.testing.assert_close(m.bias.grad, ref.bias.grad, rtol=1e-4, atol=1e-3)
