import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from test_bev_encoder_gpu import _OurEncoder
dev = torch.device("cuda:0"); torch.manual_seed(0)
net = _OurEncoder().to(dev).train()
x = torch.relu(torch.randn(8, 128, 128, 128, device=dev)).contiguous(memory_format=torch.channels_last)
g = torch.randn(8, 256, 128, 128, device=dev).contiguous(memory_format=torch.channels_last)
for i in range(3):
    xin = x.detach().requires_grad_(True)
    net.zero_grad(set_to_none=True)
    if i == 2: torch.cuda.cudart().cudaProfilerStart()
    y = net(xin); y.backward(g)
    torch.cuda.synchronize()
    if i == 2: torch.cuda.cudart().cudaProfilerStop()
