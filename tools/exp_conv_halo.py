"""Halo-reuse 3x3 conv (csrc/conv2d_tc.cu, DBEV_CONV_HALO modes) vs torch fp32 conv2d: error + time per layer shape."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import distill_bev_b200 as dbev  # noqa: E402
from distill_bev_b200.plugin import dense_teacher as dt  # noqa: E402


def timed(fn, iters=8):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    res = {"mode": os.environ.get("DBEV_CONV_HALO", "default")}
    torch.manual_seed(0)
    for (n, h, w, ci, co) in [(2, 40, 24, 64, 64), (1, 64, 64, 128, 128), (1, 37, 19, 64, 256), (1, 32, 8, 32, 64)]:
        x = torch.randn(n, ci, h, w, device=dev)
        wt = torch.randn(co, ci, 3, 3, device=dev) * 0.05
        sc, sh = torch.rand(co, device=dev) + 0.5, torch.randn(co, device=dev)
        ref = torch.relu(torch.nn.functional.conv2d(x.double(), wt.double(), padding=1) * sc.double()[None, :, None, None]
                         + sh.double()[None, :, None, None])
        wp = wt.permute(0, 2, 3, 1).reshape(co, -1).contiguous()
        out = dt.conv_nhwc(x.permute(0, 2, 3, 1).contiguous(), wp, co, 3, 3, 1, 1, sc, sh, relu=True)
        torch.cuda.synchronize()
        err = float((out.permute(0, 3, 1, 2).double() - ref).abs().max() / ref.abs().max())
        res["err_%dx%dx%d_%d_%d" % (n, h, w, ci, co)] = err
    for (h, c) in [(256, 64), (128, 128), (64, 256)]:
        x = torch.randn(8, h, h, c, device=dev)
        wp = torch.randn(c, 9 * c, device=dev) * 0.05
        sc, sh = torch.rand(c, device=dev), torch.randn(c, device=dev)
        out = torch.empty(8, h, h, c, device=dev)
        ms = timed(lambda: dt.conv_nhwc(x, wp, c, 3, 3, 1, 1, sc, sh, relu=True, out=out))
        res["ms_%d_%d" % (h, c)] = round(ms, 4)
        res["tflops_%d_%d" % (h, c)] = round(2.0 * 8 * h * h * 9 * c * c / ms / 1e9, 1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
