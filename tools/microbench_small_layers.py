"""Per-layer GPU time of the ResNetForBEVDet convs (B = 8 bench shapes) measured INSIDE a CUDA graph of 10 back-to-back
launches (eager timings of these 10-60 us kernels are dominated by Python launch overhead): forward, input gradient,
weight gradient (+ its reduce), with the TFLOP/s each reaches. Prints one JSON object."""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from distill_bev_b200 import conv_train as ct  # noqa: E402


def graph_time(fn, reps=10, iters=11):
    for _ in range(3):
        fn()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / reps)
    return statistics.median(ts)


def main():
    dev = torch.device("cuda:0")
    rows = []
    for name, n, h, cin, cout, stride in [("l1.conv1/downsample s2", 8, 128, 128, 128, 2), ("l1 s1", 8, 64, 128, 128, 1),
                                          ("l2.conv1/downsample s2", 8, 64, 128, 256, 2), ("l2 s1", 8, 32, 256, 256, 1),
                                          ("l3.conv1/downsample s2", 8, 32, 256, 512, 2), ("l3 s1", 8, 16, 512, 512, 1)]:
        ho = h // stride
        x = torch.randn(n, h, h, cin, device=dev)
        dy = torch.randn(n, ho, ho, cout, device=dev)
        w = torch.randn(cout, cin, 3, 3, device=dev) * 0.05
        wf, wb = ct.pack_weights_train(w, stride)
        gf = 2.0 * n * ho * ho * cin * cout * 9 / 1e9
        r = {"layer": name, "shape": "[%d,%d,%d,%d] -> %d ch" % (n, h, h, cin, cout), "GFLOP": round(gf, 2)}
        for key, fn in (("fwd", lambda: ct.conv_forward(x, wf, cout, 3, 3, stride, 1)),
                        ("dgrad", lambda: ct.conv_input_grad(dy, wb, cin, 3, 3, stride, 1, (h, h))),
                        ("wgrad", lambda: ct.conv_weight_grad(x, dy, 3, 3, stride, 1))):
            ms = graph_time(fn)
            r[key + "_us"] = round(ms * 1e3, 1)
            r[key + "_TFLOPs"] = round(gf / ms, 1)
        rows.append(r)
        print(r, flush=True)
    print(json.dumps({"rows": rows}))


if __name__ == "__main__":
    main()
