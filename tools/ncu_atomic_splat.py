"""Driver for an ncu capture of the sort-free lift+splat forward kernel at the bench shape."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
import distill_bev_b200 as dbev  # noqa: E402

hp = bench.HotPath(torch.device("cuda:0"), 0)
geom = hp.vt.get_geometry(*hp.d_calib)
cells = hp.vt.make_cells(geom, bench.BATCH * bench.FRAMES)
with torch.no_grad():
    for _ in range(3):
        dbev.lift_splat(hp.depth, hp.feat, cells)
torch.cuda.synchronize()
