"""Student BEV encoder (ResNetForBEVDet + FPN_LSS, B=8, 128x128x128 input) fwd+bwd: our tcgen05 training path vs
the same torch modules through cuDNN (TF32, channels_last), plus per-layer forward / input-gradient /
weight-gradient kernel times (CUDA events, median of 5). Writes gpurun_out/bev_encoder_bench.json."""
import json
import os
import sys
import time

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import distill_bev_b200 as dbev  # noqa: E402
from distill_bev_b200 import conv_train as ct  # noqa: E402
from test_bev_encoder_gpu import _OurEncoder, _RefEncoder, _load_ours_from_ref  # noqa: E402


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    B = int(os.environ.get("B", 8))
    res = {"batch": B}
    ref = _RefEncoder().to(dev).train().to(memory_format=torch.channels_last)
    ours = _OurEncoder().to(dev).train()
    _load_ours_from_ref(ours, ref)
    x = torch.relu(torch.randn(B, 128, 128, 128, device=dev)).contiguous(memory_format=torch.channels_last)
    g = torch.randn(B, 256, 128, 128, device=dev).contiguous(memory_format=torch.channels_last)

    def step(net):
        xin = x.detach().requires_grad_(True)
        y = net(xin)
        y.backward(g)
        net.zero_grad(set_to_none=True)

    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True
    res["cudnn_tf32_fwd_bwd_ms"] = timed(lambda: step(ref))
    res["ours_fwd_bwd_ms"] = timed(lambda: step(ours))
    with torch.no_grad():
        res["cudnn_tf32_fwd_ms"] = timed(lambda: ref(x))
        res["ours_fwd_ms"] = timed(lambda: ours(x))
    # CUDA graph of our step (launch overhead removed)
    try:
        cap = dbev.CapturedStep(lambda: step(ours), warmup=2, device=dev)
        res["ours_fwd_bwd_graph_ms"] = timed(cap.replay)
    except Exception as exc:  # noqa: BLE001
        res["ours_graph_error"] = repr(exc)[:300]
    layers = [("l1.conv1 s2", 128, 128, 128, 3, 2), ("l1 s1", 128, 128, 64, 3, 1), ("l2.conv1 s2", 128, 256, 64, 3, 2),
              ("l2 s1", 256, 256, 32, 3, 1), ("l3.conv1 s2", 256, 512, 32, 3, 2), ("l3 s1", 512, 512, 16, 3, 1),
              ("fpn 640->512", 640, 512, 64, 3, 1), ("fpn 512->512", 512, 512, 64, 3, 1), ("fpn 512->256 @128", 512, 256, 128, 3, 1),
              ("fpn 1x1 256->256 @128", 256, 256, 128, 1, 1)]
    rows = []
    for name, ci, co, hw, k, s in layers:
        pad = k // 2
        xin = torch.randn(B, hw, hw, ci, device=dev)
        w = torch.randn(co, ci, k, k, device=dev) * 0.05
        ho = (hw + 2 * pad - k) // s + 1
        dy = torch.randn(B, ho, ho, co, device=dev)
        wf, wb = ct.pack_weights(w, 0), ct.pack_weights(w, 1 if s == 1 else 2)
        flops = 2.0 * B * ho * ho * ci * co * k * k
        t_f = timed(lambda: ct.conv_forward(xin, wf, co, k, k, s, pad))
        t_d = timed(lambda: ct.conv_input_grad(dy, wb, ci, k, k, s, pad, (hw, hw)))
        t_w = timed(lambda: ct.conv_weight_grad(xin, dy, k, k, s, pad))
        xc = xin.permute(0, 3, 1, 2)
        dyc = dy.permute(0, 3, 1, 2)
        t_c = timed(lambda: torch.ops.aten.convolution_backward(dyc, xc, w.contiguous(memory_format=torch.channels_last), None,
                                                                [s, s], [pad, pad], [1, 1], False, [0, 0], 1, [True, True, False]))
        t_cf = timed(lambda: torch.nn.functional.conv2d(xc, w.contiguous(memory_format=torch.channels_last), None, s, pad))
        rows.append({"layer": name, "gflop": round(flops / 1e9, 1), "fwd_ms": round(t_f, 4), "dgrad_ms": round(t_d, 4),
                     "wgrad_ms": round(t_w, 4), "fwd_tflops": round(flops / t_f / 1e9, 1), "dgrad_tflops": round(flops / t_d / 1e9, 1),
                     "wgrad_tflops": round(flops / t_w / 1e9, 1), "cudnn_fwd_ms": round(t_cf, 4), "cudnn_bwd_ms": round(t_c, 4)})
    res["layers"] = rows
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bev_encoder_bench.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    t0 = time.time()
    main()
    sys.stderr.write("done in %.1f s\n" % (time.time() - t0))
