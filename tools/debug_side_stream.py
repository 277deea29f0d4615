"""Which gradients differ between the single-stream and the side-stream backward (if any), over several repetitions."""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from distill_bev_b200 import bev_encoder, conv_train as ct  # noqa: E402
from test_bev_encoder_gpu import _OurEncoder  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(7)
    net = _OurEncoder().to(dev).train()
    x = torch.relu(torch.randn(2, 128, 32, 32, device=dev)).contiguous(memory_format=torch.channels_last)
    g = torch.randn(2, 256, 32, 32, device=dev).contiguous(memory_format=torch.channels_last)
    state = copy.deepcopy(net.state_dict())
    names = [k for k, _ in net.named_parameters()] + ["input"]

    def run(side, prepack):
        ct._side["test_delay_cycles"] = int(os.environ.get("DBG_DELAY_SIDE", "0"))
        net.load_state_dict(state)
        net.zero_grad(set_to_none=True)
        ct.set_side_stream(side)
        if prepack:
            bev_encoder.prepack(net)
        xin = x.clone().requires_grad_(True)
        net(xin).backward(g)
        ct.join_side_stream(dev)
        torch.cuda.synchronize()
        out = [p.grad.clone() for p in net.parameters()] + [xin.grad.clone()]
        ct.set_side_stream(False)
        for m in net.modules():
            if hasattr(m, "_dbev_prepacked"):
                del m._dbev_prepacked
        return out

    base = run(False, False)
    for rep in range(3):
        for side, pre in ((False, False), (False, True), (True, False), (True, True)):
            got = run(side, pre)
            bad = [(n, float((a - b).abs().max()), float(b.abs().max())) for n, a, b in zip(names, got, base) if not torch.equal(a, b)]
            print("rep", rep, "side", side, "prepack", pre, "mismatches", len(bad), bad[:4])
            if bad and os.environ.get("DETAIL"):
                n0 = bad[0][0]
                i0 = names.index(n0)
                a, b = got[i0], base[i0]
                d = (a - b).abs()
                print("  tensor", n0, tuple(a.shape), "frac differing", float((d > 0).float().mean()),
                      "got zeros frac", float((a == 0).float().mean()), "base zeros frac", float((b == 0).float().mean()))
                if a.dim() == 4:
                    co, ci = a.shape[0], a.shape[1]
                    blk = d.reshape(co // 128, 128, ci // 128, 128, a.shape[2], a.shape[3]).amax(dim=(1, 3))
                    print("  max diff per (co block, ci block, ky, kx):", blk.flatten().tolist())
                    print("  ratio got/base at differing entries (first 8):", (a[d > 0][:8] / b[d > 0][:8]).tolist())


if __name__ == "__main__":
    main()
