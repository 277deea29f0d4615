"""Single launches of the training conv kernels on the step's largest layers, for `ncu --set full`:
  conv3x3_halo_kernel<256>   forward  512 -> 256 @ 128x128 (B=8)    and its input gradient (256 -> 512, 2 column blocks)
  conv_wgrad_tc_kernel       weight gradient of the same layer, and of 640 -> 512 @ 64x64
  conv3x3_halo_kernel<64>    512 -> 512 @ 16x16 (column blocks), conv2d_tc_kernel stride-2 dgrad class
Usage: ncu --set full --clock-control none --import-source on -k regex:'conv|wgrad' -o gpurun_out/r02_conv_train python tools/ncu_conv_train.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distill_bev_b200 import conv_train as ct  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    B = 8
    for ci, co, hw, s in ((512, 256, 128, 1), (640, 512, 64, 1), (512, 512, 16, 1), (256, 512, 32, 2)):
        x = torch.randn(B, hw, hw, ci, device=dev)
        w = torch.randn(co, ci, 3, 3, device=dev) * 0.05
        ho = hw // s
        dy = torch.randn(B, ho, ho, co, device=dev)
        wf, wb = ct.pack_weights_train(w, s)
        ct.conv_forward(x, wf, co, 3, 3, s, 1)
        ct.conv_input_grad(dy, wb, ci, 3, 3, s, 1, (hw, hw))
        ct.conv_weight_grad(x, dy, 3, 3, s, 1)
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
