"""SECOND + SECONDFPN teacher forward (B=8, canvas [8,64,512,512] -> [8,384,128,128]) on the tcgen05 conv
kernel vs the same torch modules through cuDNN (TF32 allowed, channels_last), CUDA events."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import distill_bev_b200 as dbev  # noqa: E402


def timed(fn, iters=8):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main(batch=8):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = dbev.SECOND(64, [64, 128, 256], [3, 5, 5], [2, 2, 2]).to(dev).eval()
    fpn = dbev.SECONDFPN([64, 128, 256], [128, 128, 128], [0.5, 1, 2]).to(dev).eval()
    x = torch.relu(torch.randn(batch, 64, 512, 512, device=dev)).contiguous(memory_format=torch.channels_last)
    flops = 0.0
    hw, cin = 512, 64
    for n, c in zip((3, 5, 5), (64, 128, 256)):
        hw //= 2
        flops += 2.0 * batch * hw * hw * 9 * cin * c + n * 2.0 * batch * hw * hw * 9 * c * c
        cin = c
    flops += 2.0 * batch * 128 * 128 * (4 * 64 * 128 + 128 * 128 + 256 * 128)
    res = {"batch": batch, "GFLOP": flops / 1e9}
    with torch.no_grad():
        t_ours = timed(lambda: fpn(net(x))[0])

        def ref():
            h, outs = x, []
            for b in net.blocks:
                h = b(h)
                outs.append(h)
            return torch.cat([d(o) for d, o in zip(fpn.deblocks, outs)], 1)
        net_cl = net.to(memory_format=torch.channels_last)
        fpn_cl = fpn.to(memory_format=torch.channels_last)
        t_cudnn = timed(ref)
        torch.backends.cudnn.allow_tf32 = False
        t_cudnn_fp32 = timed(ref)
        torch.backends.cudnn.allow_tf32 = True
    res.update(tcgen05_ms=t_ours, tcgen05_TFLOPs=flops / t_ours / 1e9, cudnn_tf32_ms=t_cudnn,
               cudnn_tf32_TFLOPs=flops / t_cudnn / 1e9, cudnn_fp32_ms=t_cudnn_fp32)
    print(json.dumps(res))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 8)
