"""Golden fixture for the hard-voxel PillarFeatureNet from the UNMODIFIED reference class
(mmdet3d/models/voxel_encoders/pillar_encoder.py:14-162 + utils.py PFNLayer), eval mode, legacy False and True, on
voxels produced by the reference's own CPU hard_voxelize (oracle/_ref/ref_voxel_layer.so). Writes
tests/golden/pillar_hard.npz. Build container only."""
import glob
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_import  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("_syn", os.path.join(ROOT, "distill-bev_b200", "synthetic.py"))
    syn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(syn)
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "ref_voxel_layer.so"))[0]
    spec = importlib.util.spec_from_file_location("ref_voxel_layer", so)
    vl = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vl)
    ref_import.pillar_modules()
    pe = sys.modules["mmdet3d.models.voxel_encoders.pillar_encoder"]
    vs, rng, T = [0.2, 0.2, 8.0], [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], 20
    vox, cos, nps = [], [], []
    for b in range(2):
        pts = torch.from_numpy(syn.make_lidar(1, 1200, seed=20 + b)[0])
        v = torch.zeros(5000, T, 5)
        c = torch.zeros(5000, 3, dtype=torch.int32)
        k = torch.zeros(5000, dtype=torch.int32)
        m = vl.hard_voxelize(pts, v, c, k, vs, rng, T, 5000, 3, True)
        vox.append(v[:m]), nps.append(k[:m])
        cos.append(torch.nn.functional.pad(c[:m], (1, 0), value=b))
    voxels, num_points, coors = torch.cat(vox), torch.cat(nps), torch.cat(cos)
    out = {"voxels": voxels.numpy(), "num_points": num_points.numpy(), "coors": coors.numpy(), "voxel_size": np.array(vs),
           "range": np.array(rng)}
    for legacy in (False, True):
        torch.manual_seed(3)
        net = pe.PillarFeatureNet(in_channels=5, feat_channels=[64], with_distance=False, voxel_size=tuple(vs),
                                  point_cloud_range=tuple(rng), norm_cfg=dict(type="BN1d", eps=1e-3, momentum=0.01),
                                  legacy=legacy).eval()
        bn = net.pfn_layers[0].norm
        with torch.no_grad():
            bn.running_mean.normal_(0, 0.3)
            bn.running_var.uniform_(0.5, 1.5)
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.normal_(0, 0.3)
            y = net(voxels.clone(), num_points, coors)
        tag = "legacy" if legacy else "new"
        out["out_" + tag] = y.numpy()
        if not legacy:
            out["keys"] = np.array(list(net.state_dict().keys()))
            for k2, v2 in net.state_dict().items():
                out["sd/" + k2] = v2.numpy()
    # MVP virtual points (virtual=True, pillar_encoder.py:108-113): the second-to-last feature is the virtual label,
    # -1 for a virtual point; the same voxels with ~40 % of the points marked virtual, same weights as the runs above
    vmask = torch.rand(voxels.shape[:2], generator=torch.Generator().manual_seed(11)) < 0.4
    feats_v = voxels.clone()
    feats_v[..., -2][vmask] = -1.0
    torch.manual_seed(3)
    net = pe.PillarFeatureNet(in_channels=5, feat_channels=[64], with_distance=False, voxel_size=tuple(vs),
                              point_cloud_range=tuple(rng), norm_cfg=dict(type="BN1d", eps=1e-3, momentum=0.01),
                              legacy=False, virtual=True).eval()
    net.load_state_dict({k2: torch.from_numpy(out["sd/" + k2]) for k2 in out["keys"]})
    with torch.no_grad():
        out["out_virtual"] = net(feats_v, num_points, coors).numpy()
    out["virtual_mask"] = np.packbits(vmask.numpy(), axis=1)
    path = os.path.join(ROOT, "tests", "golden", "pillar_hard.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, voxels.shape, "%.2f MB" % (os.path.getsize(path) / 1e6))


if __name__ == "__main__":
    main()
