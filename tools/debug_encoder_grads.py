"""Per-parameter gradient error of the BEV encoder vs the fp32 torch modules: ours (tcgen05 TF32) next to
cuDNN TF32 (the reference's own GPU arithmetic under torch defaults)."""
import copy
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_bev_encoder_gpu import _OurEncoder, _RefEncoder, _load_ours_from_ref  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    batch, hw = int(os.environ.get("B", 2)), int(os.environ.get("HW", 64))
    ref = _RefEncoder().to(dev).train()
    ours = _OurEncoder().to(dev).train()
    _load_ours_from_ref(ours, ref)
    tf = copy.deepcopy(ref)
    x = torch.relu(torch.randn(batch, 128, hw, hw, device=dev))
    g = None
    outs = {}
    for name, net, tf32 in (("fp32", ref, False), ("cudnn_tf32", tf, True), ("ours", ours, False)):
        torch.backends.cudnn.allow_tf32 = tf32
        xin = x.clone().contiguous(memory_format=torch.channels_last).requires_grad_(True)
        y = net(xin)
        if g is None:
            g = torch.randn_like(y) / y.numel() ** 0.5
        (y * g).sum().backward()
        grads = {}
        for k, p in net.named_parameters():
            k = k.replace("backbone.", "").replace("neck.", "")
            grads[k] = p.grad.detach().clone()
        grads["__x"] = xin.grad.detach().clone()
        grads["__y"] = y.detach().clone()
        outs[name] = grads
    torch.backends.cudnn.allow_tf32 = False

    def err(a, b):
        return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20)), float(F.cosine_similarity(a.flatten(), b.flatten(), dim=0))

    for k in outs["fp32"]:
        e1, c1 = err(outs["cudnn_tf32"][k], outs["fp32"][k])
        e2, c2 = err(outs["ours"][k], outs["fp32"][k])
        print("%-34s cudnn_tf32 err %.2e cos %.6f | ours err %.2e cos %.6f" % (k, e1, c1, e2, c2))


if __name__ == "__main__":
    main()
