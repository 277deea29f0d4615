"""One forward of the frozen dense teacher (SECOND + SECONDFPN, B = 8) between cudaProfilerStart / Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum ...` (per-launch list) or `--set full`."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import distill_bev_b200 as dbev  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = dbev.SECOND(64, [64, 128, 256], [3, 5, 5], [2, 2, 2]).to(dev).eval()
    fpn = dbev.SECONDFPN([64, 128, 256], [128, 128, 128], [0.5, 1, 2]).to(dev).eval()
    x = torch.relu(torch.randn(8, 64, 512, 512, device=dev)).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        for _ in range(2):
            fpn(net(x))
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        fpn(net(x))
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
