"""CPU: a converted conv on a CPU tensor is plain nn.Conv2d (no tcgen05 path, no oracle)."""
import torch
import torch.nn as nn

import distill_bev_b200 as dbev


def test_cpu_goes_through_torch():
    conv = dbev.convert_convs(nn.Sequential(nn.Conv2d(64, 64, 3, 1, 1)))
    x = torch.randn(1, 64, 16, 16)
    torch.testing.assert_close(conv(x), torch.nn.functional.conv2d(x, conv[0].weight, conv[0].bias, 1, 1))
