import os, sys, collections
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, bench
import distill_bev_b200 as dbev
dev = torch.device("cuda:0")
hp = bench.HotPath(dev, 0)
def loss():
    losses = dbev.fgd.fgd_distill_loss(hp.teacher, hp.student, hp.boxes, bench.DISTILL_PARAMS, bench.TRAIN_CFG,
        channel_adaptation=hp.adapt, spatial_adaptation=hp.spatial, heatmaps=hp.d_gt_hm, teacher_heatmaps=hp.teacher_logit, epoch=1)
    return losses["kd_fg_feat_loss"] + losses["kd_bg_feat_loss"] + losses["kd_spatial_loss"] + losses["kd_fp_bg_feat_loss"]
for _ in range(3):
    loss().backward()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile
for name, fn in (("fwd", lambda: loss()), ("fwd+bwd", lambda: loss().backward())):
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(); torch.cuda.synchronize()
    print("==", name)
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            print("%8.1f us  %s" % (e.device_time_total, e.name[:110]))
