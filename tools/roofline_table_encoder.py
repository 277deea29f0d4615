"""Round 2 companion of tools/roofline_table.py: the memory-bound kernels of the student BEV encoder's training step
and the fused adaptation + loss kernel, each alone on the step's largest shape: CUDA-event time (median of 15, L2
flushed between launches), algorithmic bytes, GB/s and the fraction of the measured HBM peak (MEASURED_PEAKS.json).

    python tools/roofline_table_encoder.py > gpurun_out/roofline_table_encoder.json
"""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    import distill_bev_b200 as dbev
    from distill_bev_b200 import conv_train as ct
    from distill_bev_b200 import bev_encoder
    dev = torch.device("cuda:0")
    peak, how = bench.measured_peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []

    def timed(fn, iters=15):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(True), torch.cuda.Event(True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts)

    def add(name, shape, fn, nbytes, note=None):
        ms = timed(fn)
        r = {"op": name, "shape": shape, "ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1),
             "GBps": round(nbytes / ms / 1e6, 1), "hbm_frac": round(nbytes / ms / 1e6 / peak, 3)}
        if note:
            r["note"] = note
        rows.append(r)

    n, h, w, c = 8, 128, 128, 256                     # output of the FPN's 512 -> 256 conv: the largest BatchNorm of the step
    m = n * h * w * c * 4
    y = torch.randn(n, h, w, c, device=dev)
    dz = torch.randn(n, h, w, c, device=dev)
    gamma, beta = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev)
    ws = ct.stats_workspace(n * h * w, c, dev)
    fwd = ct.bn_batch_stats(y, gamma, beta, 1e-5, ws=ws)
    z, mask = ct.bn_act(y, fwd, None, True, want_mask=True)
    shape = "[8,128,128,256] NHWC fp32"
    add("BatchNorm batch statistics (channel_stats<0> + finalize)", shape, lambda: ct.bn_batch_stats(y, gamma, beta, 1e-5, ws=ws), m,
        "reads y once")
    add("BatchNorm apply + ReLU + mask (bn_act)", shape, lambda: ct.bn_act(y, fwd, None, True, want_mask=True), 2 * m + m // 16,
        "reads y, writes z and the 1-byte-per-quad ReLU mask")
    add("BatchNorm backward (channel_stats<1> + finalize + bn_bwd_apply)", shape,
        lambda: ct.bn_backward(dz, None, y, fwd, ws=ws, mask=mask), 5 * m + m // 8,
        "two passes: (dz, y, mask) read twice, dy written")
    add("BatchNorm backward reading z instead of the mask (round-2 first version)", shape,
        lambda: ct.bn_backward(dz, z, y, fwd, ws=ws), 7 * m, "(dz, y, z) read twice, dy written")
    add("conv bias gradient (channel_sums)", shape, lambda: ct.channel_sums(dz, ws=ws), m)
    x = torch.randn(8, 64, 64, 512, device=dev)
    big = torch.randn(8, 128, 128, 512, device=dev)
    add("bilinear x2 forward (align_corners)", "[8,64,64,512] -> [8,128,128,512]", lambda: ct.upsample_bilinear(x, 2),
        x.numel() * 4 + big.numel() * 4)
    add("bilinear x2 backward (gather)", "[8,128,128,512] -> [8,64,64,512]", lambda: ct.upsample_bilinear_backward(big, (64, 64)),
        x.numel() * 4 + big.numel() * 4, "L2-bound: every output-gradient element is read by up to four input pixels")
    x4 = torch.randn(8, 16, 16, 512, device=dev)
    big4 = torch.randn(8, 64, 64, 512, device=dev)
    add("bilinear x4 forward", "[8,16,16,512] -> [8,64,64,512]", lambda: ct.upsample_bilinear(x4, 4), x4.numel() * 4 + big4.numel() * 4)
    add("bilinear x4 backward", "[8,64,64,512] -> [8,16,16,512]", lambda: ct.upsample_bilinear_backward(big4, (16, 16)),
        x4.numel() * 4 + big4.numel() * 4)
    # weight packing: all conv layers of the encoder in one launch
    torch.manual_seed(0)
    net = torch.nn.ModuleList([dbev.ResNetForBEVDet(numC_input=2 * bench.C_TRANS, num_channels=bench.ENC_CHANNELS),
                               dbev.FPN_LSS(in_channels=bench.ENC_CHANNELS[-1] + bench.ENC_CHANNELS[0], out_channels=256)]).to(dev)
    n_w = sum(mod.weight.numel() for mod in net.modules() if isinstance(mod, torch.nn.Conv2d))
    bev_encoder.prepack(net)
    add("weight packing, all layers (pack_conv_weights_batch)", "%d conv weights = %.1f M floats" % (
        sum(1 for mod in net.modules() if isinstance(mod, torch.nn.Conv2d)), n_w / 1e6), lambda: bev_encoder.prepack(net), 3 * n_w * 4,
        "reads the weights once, writes the forward and the input-gradient matrices; 128-byte runs")
    # weight gradient of the largest layer: split-K partials + reduce
    xa = torch.randn(8, 128, 128, 512, device=dev)
    dya = torch.randn(8, 128, 128, 256, device=dev)
    fl = 2.0 * 8 * 128 * 128 * 512 * 256 * 9
    ms = timed(lambda: ct.conv_weight_grad(xa, dya, 3, 3, 1, 1))
    mp = os.path.join(bench.ROOT, "MEASURED_PEAKS.json")
    tf32 = float(json.load(open(mp))["bf16_tflops"]) / 2.0 if os.path.exists(mp) else 1125.0
    rows.append({"op": "weight gradient 512 -> 256 3x3 @128^2 (conv_wgrad_tc + wgrad_reduce)", "ms": round(ms, 4), "GFLOP": round(fl / 1e9, 1),
                 "TFLOPs": round(fl / ms / 1e9, 1), "tf32_frac": round(fl / ms / 1e9 / tf32, 3), "tf32_peak_TFLOPs": tf32})
    print(json.dumps({"hbm_peak_GBps": peak, "peak_source": how, "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
