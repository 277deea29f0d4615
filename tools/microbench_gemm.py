"""GPU microbench: tcgen05 1x1 adaptation conv vs cuDNN (torch) at the configs[1] head shape."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import distill_bev_b200  # noqa: E402,F401
from distill_bev_b200.plugin.distill.adaptation import conv1x1  # noqa: E402


def t(fn, flush, it=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(it):
        flush.zero_()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    dev = torch.device("cuda:0")
    B, Cin, Cout, H = 8, 256, 384, 128
    x = torch.relu(torch.randn(B, Cin, H, H, device=dev))
    xcl = x.contiguous(memory_format=torch.channels_last)
    conv = torch.nn.Conv2d(Cin, Cout, 1).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flops = 2.0 * B * H * H * Cin * Cout
    res = {"shape": [B, Cin, Cout, H], "gflop": flops / 1e9}
    with torch.no_grad():
        res["cudnn_nchw_ms"] = t(lambda: conv(x), flush)
        res["cudnn_cl_ms"] = t(lambda: conv(xcl), flush)
        res["ours_nchw_ms"] = t(lambda: conv1x1(x, conv.weight, conv.bias), flush)
        res["ours_cl_ms"] = t(lambda: conv1x1(xcl, conv.weight, conv.bias), flush)
    res["ours_cl_tflops"] = flops / res["ours_cl_ms"] / 1e9
    res["ours_cl_GBps"] = (x.numel() + B * Cout * H * H) * 4 / res["ours_cl_ms"] / 1e6
    print(json.dumps(res))


if __name__ == "__main__":
    main()
