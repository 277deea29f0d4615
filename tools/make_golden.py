"""Generate tests/golden/*.npz by EXECUTING the reference's own Python code.

Run in the build container only (needs /root/reference):

    python tools/make_golden.py

The reference ships no tests or golden vectors; these fixtures
are the pin for ``oracle/``: unmodified reference files are imported through
``tools/ref_import.py`` (import stubs only, no arithmetic changed) and run on
the CPU on small seeded inputs. The fixtures are small (tens of KB) and are
committed; the GPU box never sees /root/reference.

Fixtures
  lss_small.npz    ViewTransformerLiftSplatShoot: create_frustum, get_geometry,
                   voxel_pooling (cumsum path), voxel_pooling_accelerated,
                   autograd gradient of voxel_pooling w.r.t. x
  lss_edge.npz     hand-placed points on cell borders / outside the grid
                   (trunc-toward-zero leak, dropped points, empty batch item)
  quickcumsum.npz  QuickCumsum of mmdet3d/ops/bev_pool/bev_pool.py (pure torch)
  voxel_small.npz  reference voxel_layer CPU extension (oracle/_ref): dynamic_voxelize,
                   hard_voxelize x3 configs; dynamic scatter via torch.unique(dim=0)
  fgd_small.npz    BEVDetDistill.foreground_scale_mask / add_fp_as_fg / fgd_distill_loss /
                   affinity_distill_loss method bodies (ast-extracted, unmodified): masks,
                   losses, autograd gradients for three option sets
  pillar_small.npz reference DynamicPillarFeatureNet (eval) + PointPillarsScatter, unmodified
                   files; DynamicScatter (CUDA-only) replaced by its host sequence on torch CPU
Also prints (not stored) the full-size config-1 comparison oracle vs reference.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import ref_import  # noqa: E402
import distill_bev_b200  # noqa: E402,F401
from distill_bev_b200 import synthetic  # noqa: E402
from oracle import lss_oracle  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def build_vt(vtm, grid, input_size, downsample, numC_Trans):
    torch.manual_seed(0)
    return vtm.ViewTransformerLiftSplatShoot(
        grid_config=grid, data_config={"input_size": input_size}, numC_input=8,
        numC_Trans=numC_Trans, downsample=downsample, accelerate=False)


def lss_small(vtm):
    grid = dict(xbound=[-8.0, 8.0, 1.0], ybound=[-8.0, 8.0, 1.0], zbound=[-10.0, 10.0, 20.0],
                dbound=[1.0, 9.0, 1.0])
    input_size, down, C = (64, 96), 16, 8
    vt = build_vt(vtm, grid, input_size, down, C)
    B, N = 2, 3
    rots, trans, intrins, post_rots, post_trans = synthetic.make_calibration(
        B, N, seed=3, input_size=input_size, src_size=(225, 400))
    # shrink the scene so that many points land inside the +-8 m grid
    intrins[:, :, 0, 0] = intrins[:, :, 1, 1] = 300.0
    intrins[:, :, 0, 2], intrins[:, :, 1, 2] = 200.0, 112.0
    t = [torch.from_numpy(a) for a in (rots, trans, intrins, post_rots, post_trans)]
    geom = vt.get_geometry(*t)
    D, fH, fW = vt.frustum.shape[:3]
    g = torch.Generator().manual_seed(1)
    x = torch.rand(B, N, D, fH, fW, C, generator=g)
    x.requires_grad_(True)
    out = vt.voxel_pooling(geom, x)
    w = torch.rand(out.shape, generator=g)
    (out * w).sum().backward()
    out_acc = vt.voxel_pooling_accelerated(geom, x.detach())
    # lift + splat exactly as ViewTransformerLiftSplatShoot.forward does it (:250-262)
    depth = torch.rand(B * N, D, fH, fW, generator=g).softmax(dim=1).requires_grad_(True)
    feat = torch.randn(B * N, C, fH, fW, generator=g).requires_grad_(True)
    volume = depth.unsqueeze(1) * feat.unsqueeze(2)
    volume = volume.view(B, N, C, D, fH, fW).permute(0, 1, 3, 4, 5, 2)
    lift_out = vt.voxel_pooling_accelerated(geom, volume)
    (lift_out * w).sum().backward()
    np.savez_compressed(
        os.path.join(GOLDEN, "lss_small.npz"),
        lift_depth=depth.detach().numpy(), lift_feat=feat.detach().numpy(),
        lift_out=lift_out.detach().numpy(), lift_ddepth=depth.grad.numpy(),
        lift_dfeat=feat.grad.numpy(),
        grid=json.dumps(grid), input_size=np.array(input_size), downsample=down,
        rots=rots, trans=trans, intrins=intrins, post_rots=post_rots, post_trans=post_trans,
        frustum=vt.frustum.detach().numpy(), geom=geom.detach().numpy(),
        dx=vt.dx.numpy(), bx=vt.bx.numpy(), nx=vt.nx.numpy(),
        x=x.detach().numpy(), out_cumsum=out.detach().numpy(), out_accelerated=out_acc.numpy(),
        out_weight=w.numpy(), x_grad=x.grad.numpy())
    print("lss_small: out", tuple(out.shape), "nonzero cells",
          int((out.detach().abs().sum(1) > 0).sum()), "max|cumsum-accel|",
          float((out.detach() - out_acc).abs().max()))


def lss_edge(vtm):
    grid = dict(xbound=[-4.0, 4.0, 1.0], ybound=[-2.0, 2.0, 0.5], zbound=[-10.0, 10.0, 20.0],
                dbound=[1.0, 3.0, 1.0])
    vt = build_vt(vtm, grid, (32, 32), 16, 4)
    # geometry given directly: B=3 (last sample entirely outside), N=1, D=2, H=2, W=2
    pts = np.array([
        [-4.0, -2.0, 0.0], [-4.5, -2.0, 0.0], [-4.999, -2.25, -29.0], [-5.0, 0.0, 0.0],   # low edge leak
        [3.999, 1.999, 9.9], [4.0, 0.0, 0.0], [0.0, 2.0, 0.0], [0.0, 0.0, 10.0],          # high edge
        [0.5, 0.25, 0.0], [0.5, 0.25, 1.0], [0.5, 0.25, -5.0], [0.999, 0.499, 3.0],       # many-to-one
        [-0.5, -0.25, 0.0], [-1.0, -0.5, 0.0], [1e9, 0.0, 0.0], [0.0, -1e9, 0.0],         # far away
        [100.0, 100.0, 0.0], [-100.0, 0.0, 0.0], [0.0, 50.0, 0.0], [9.0, 9.0, 9.0],       # sample 2: all out
        [8.0, 0.0, 0.0], [0.0, 3.0, 0.0], [0.0, 0.0, 31.0], [-6.0, -6.0, 0.0],
    ], dtype=np.float32)
    geom = torch.from_numpy(pts).view(3, 1, 2, 2, 2, 3)
    g = torch.Generator().manual_seed(2)
    x = torch.rand(3, 1, 2, 2, 2, 4, generator=g)
    x.requires_grad_(True)
    out = vt.voxel_pooling(geom, x)
    w = torch.rand(out.shape, generator=g)
    (out * w).sum().backward()
    np.savez_compressed(
        os.path.join(GOLDEN, "lss_edge.npz"), grid=json.dumps(grid), geom=geom.numpy(),
        dx=vt.dx.numpy(), bx=vt.bx.numpy(), nx=vt.nx.numpy(), x=x.detach().numpy(),
        out_cumsum=out.detach().numpy(), out_weight=w.numpy(), x_grad=x.grad.numpy())
    print("lss_edge: kept rows with grad", int((x.grad.abs().sum(-1) > 0).sum()), "of 24")


def quickcumsum(bp):
    g = torch.Generator().manual_seed(5)
    B, D, H, W, C, n = 2, 2, 6, 5, 8, 400
    coords = torch.stack([torch.randint(0, H, (n,), generator=g), torch.randint(0, W, (n,), generator=g),
                          torch.randint(0, D, (n,), generator=g), torch.randint(0, B, (n,), generator=g)], 1)
    feats = torch.rand(n, C, generator=g)
    ranks = coords[:, 0] * (W * D * B) + coords[:, 1] * (D * B) + coords[:, 2] * B + coords[:, 3]
    idx = ranks.argsort()
    xs, cs, rs = feats[idx], coords[idx], ranks[idx]
    xs.requires_grad_(True)
    xo, go = bp.QuickCumsum.apply(xs, cs, rs)
    wgt = torch.rand(xo.shape, generator=g)
    (xo * wgt).sum().backward()
    # dense [B, C, D, H, W] as bev_pool() would return after permute(0,4,1,2,3)
    dense = torch.zeros(B, D, H, W, C)
    dense[go[:, 3], go[:, 2], go[:, 0], go[:, 1]] = xo.detach()
    np.savez_compressed(
        os.path.join(GOLDEN, "quickcumsum.npz"), B=B, D=D, H=H, W=W, feats=feats.numpy(),
        coords=coords.numpy(), sort_index=idx.numpy(), x_pooled=xo.detach().numpy(),
        geom_pooled=go.numpy(), dense=dense.permute(0, 4, 1, 2, 3).contiguous().numpy(),
        weight=wgt.numpy(), x_sorted_grad=xs.grad.numpy())
    print("quickcumsum: intervals", xo.shape[0], "of", n, "rows")


def fullsize_check(vtm):
    """Config 1 (B=1, 6 cams, D=59, 16x44, C=64 -> 128x128): oracle vs reference, not stored."""
    grid = synthetic.NUSC_GRID
    vt = build_vt(vtm, grid, (256, 704), 16, 64)
    calib = synthetic.make_calibration(1, 6, seed=0)
    geom = vt.get_geometry(*[torch.from_numpy(a) for a in calib])
    D, fH, fW = vt.frustum.shape[:3]
    x = torch.from_numpy(synthetic.make_frustum_feats(6 * D * fH * fW, 64, seed=0)).view(1, 6, D, fH, fW, 64)
    t0 = time.perf_counter()
    ref = vt.voxel_pooling(geom, x).numpy()
    t_ref = time.perf_counter() - t0
    ref_acc = vt.voxel_pooling_accelerated(geom, x).numpy()
    ours = lss_oracle.voxel_pooling(geom.numpy(), x.numpy(), vt.bx.numpy(), vt.dx.numpy(), vt.nx.numpy())
    og = lss_oracle.get_geometry(vt.frustum.numpy(), *calib)
    idx_ref = ((geom - (vt.bx - vt.dx / 2.)) / vt.dx).long().view(-1, 3).numpy()
    idx_or, kept = lss_oracle.voxel_indices(geom.numpy(), vt.bx.numpy(), vt.dx.numpy(), vt.nx.numpy())
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))
    rep = dict(kept=int(kept.sum()), nprime=int(kept.size),
               nonempty_cells=int((np.abs(ref_acc).sum(1) > 0).sum()),
               idx_equal_on_kept=bool((idx_ref[kept] == idx_or[kept]).all()),
               rel_oracle_vs_accelerated=rel(ours, ref_acc), rel_oracle_vs_cumsum=rel(ours, ref),
               rel_cumsum_vs_accelerated=rel(ref, ref_acc),
               geom_max_abs_diff=float(np.abs(og - geom.numpy()).max()),
               ref_voxel_pooling_cpu_s=t_ref, threads=torch.get_num_threads())
    print("fullsize:", json.dumps(rep))
    with open(os.path.join(GOLDEN, "fullsize_report.json"), "w") as f:
        json.dump(rep, f, indent=1)


def voxel_small():
    """Reference CPU extension (oracle/_ref, compiled unmodified) on a small cloud with
    out-of-range points; dynamic scatter through torch.unique(dim=0) = at::unique_dim."""
    from oracle import voxel_oracle
    ref = voxel_oracle.load_reference_voxel_layer()
    assert ref is not None, "run python oracle/build_oracle.py first"
    rng = np.random.RandomState(7)
    n = 3000
    pts = np.zeros((n, 5), dtype=np.float32)
    pts[:, 0] = rng.uniform(-12.0, 12.0, n)
    pts[:, 1] = rng.uniform(-9.0, 9.0, n)
    pts[:, 2] = rng.uniform(-4.0, 2.0, n)
    pts[:, 3:] = rng.random_sample((n, 2))
    pts[:40, 0] = 10.0   # exactly on the upper x border -> out
    pts[40:80, 1] = -8.0  # exactly on the lower y border -> in
    pts[80:200, :3] = pts[200:320, :3]  # duplicates -> multi-point voxels
    vs, pcr = [0.5, 0.5, 1.0], [-10.0, -8.0, -3.0, 10.0, 8.0, 1.0]
    tp = torch.from_numpy(pts)
    coors = torch.zeros(n, 3, dtype=torch.int32)
    ref.dynamic_voxelize(tp, coors, vs, pcr, 3)
    out = dict(points=pts, voxel_size=np.array(vs, np.float32), coors_range=np.array(pcr, np.float32),
               dyn_coors=coors.numpy())
    for tag, (mp, mv) in dict(a=(5, 4000), b=(2, 300), c=(1, 100000 // 50)).items():
        v = torch.zeros(mv, mp, 5)
        c = torch.zeros(mv, 3, dtype=torch.int32)
        k = torch.zeros(mv, dtype=torch.int32)
        m = ref.hard_voxelize(tp, v, c, k, vs, pcr, mp, mv, 3, True)
        out["hard_%s_cfg" % tag] = np.array([mp, mv, m])
        out["hard_%s_voxels" % tag] = v[:m].numpy()
        out["hard_%s_coors" % tag] = c[:m].numpy()
        out["hard_%s_num" % tag] = k[:m].numpy()
    # dynamic scatter: scatter_points_cuda.cu:202-214 host sequence, executed with torch CPU
    clean = coors.masked_fill(coors.lt(0).any(-1, True), -1)
    oc, inv, cnt = torch.unique(clean, dim=0, sorted=True, return_inverse=True, return_counts=True)
    if oc[0, 0] < 0:
        oc, cnt, inv = oc[1:], cnt[1:], inv - 1
    feats = torch.from_numpy(rng.randn(n, 7).astype(np.float32))
    valid = inv >= 0
    m = oc.shape[0]
    ix = inv[valid][:, None].expand(-1, 7)
    mx = torch.full((m, 7), -float("inf")).scatter_reduce(0, ix, feats[valid], "amax")
    sm = torch.zeros(m, 7, dtype=torch.float64).index_add_(0, inv[valid], feats[valid].double())
    out.update(sc_feats=feats.numpy(), sc_out_coors=oc.numpy(), sc_map=inv.to(torch.int32).numpy(),
               sc_count=cnt.to(torch.int32).numpy(), sc_max=mx.numpy(), sc_sum64=sm.numpy())
    np.savez_compressed(os.path.join(GOLDEN, "voxel_small.npz"), **out)
    print("voxel_small: invalid points", int((coors[:, 0] < 0).sum()), "unique voxels", m)


def _fgd_self(methods, AttrDict, params, C, grid, pc_range, voxel):
    """Fake `self` carrying exactly what the reference methods read."""
    class Head(object):
        train_cfg = dict(grid_size=grid, point_cloud_range=pc_range, voxel_size=voxel)

    class Fake(object):
        pass
    for name, fn in methods.items():
        setattr(Fake, name, fn)
    me = Fake()
    me.distill_params = AttrDict(params)
    me.pts_bbox_head = Head()
    me._epoch = 5
    me.count = 0
    me.teacher_adaptations = [torch.nn.Identity()]
    me.channel_wise_adaptations = [torch.nn.Identity()]
    torch.manual_seed(11)
    me.spatial_wise_adaptations = [torch.nn.Conv2d(1, 1, kernel_size=3, stride=1, padding=1)]
    return me


class _Boxes(object):   # LiDARInstance3DBoxes stand-in: the methods only read .tensor
    def __init__(self, t):
        self.tensor = t


def fgd_golden():
    methods, AttrDict = ref_import.load_fgd_methods()
    B, C, H = 3, 12, 32
    grid, pc_range, voxel = [256, 256, 40], [-12.8, -12.8, -5.0, 12.8, 12.8, 3.0], [0.1, 0.1, 0.2]
    rng = np.random.RandomState(21)
    boxes = []
    for b, m in enumerate([6, 0, 11]):
        bx = np.zeros((m, 9), dtype=np.float32)
        bx[:, 0:2] = rng.uniform(-11, 11, (m, 2))
        bx[:, 2] = rng.uniform(-2, 0, m)
        bx[:, 3:5] = rng.uniform(0.6, 4.5, (m, 2))
        bx[:, 5] = rng.uniform(1, 3, m)
        bx[:, 6] = rng.uniform(-3.14, 3.14, m)
        boxes.append(bx)
    g = torch.Generator().manual_seed(4)
    teacher = torch.relu(torch.randn(B, C, H, H, generator=g))
    student = torch.relu(torch.randn(B, C, H, H, generator=g))
    canvas = torch.rand(B, 1, H * 4, H * 4, generator=g)
    # heatmaps: 2 tasks x (1, 2 classes); gt sparse gaussians, teacher logits, student probs
    gt_hm = [torch.rand(B, k, H, H, generator=g) ** 8 for k in (1, 2)]
    t_logit = [torch.randn(B, k, H, H, generator=g) * 1.5 - 2.0 for k in (1, 2)]
    s_prob = [torch.rand(B, k, H, H, generator=g) * 0.3 for k in (1, 2)]
    base = dict(spatial_t=0.5, spatial_student_ratio=1.0, channel_t=0.5,
                fg_feat_loss_weights=[6e-3], bg_feat_loss_weights=[4e-2], channel_loss_weights=[0.25],
                spatial_loss_weights=[2.5e-3], spatial_attentions=["teacher_student"],
                feat_criterion=dict(type="MSELoss", reduction="none"),
                spatial_criterion=dict(type="L1Loss", reduction="none"),
                channel_criterion=dict(type="L1Loss", reduction="none"),
                transpose_mask=False, foreground_mask="gt", background_mask="logical_not",
                scale_mask="combine_gt", spatial_mask=True, channel_mask=False,
                student_feat_pos=["head"], teacher_feat_pos=["head"], affinity_mode=["none"],
                non_empty_weight=0, output_threshold=0.1, groundtruth_threshold=None,
                fp_as_foreground=["teacher"], fp_weight=6e-2, fp_epoch=0, fp_scale_mode="average",
                context_length=0, context_weight=0)
    variants = dict(
        recipe=dict(),                                                 # scripts/teacher_to_bevdepth4d recipe
        baseconfig=dict(spatial_attentions=["teacher"], channel_mask=True,
                        fp_as_foreground=["none"], fg_feat_loss_weights=[1.5e-3]),  # shipped .py config
        separate=dict(scale_mask="separate_gt", channel_mask=True, fp_as_foreground=["student"]),
    )
    out = dict(teacher=teacher.numpy(), student=student.numpy(), grid=np.array(grid),
               pc_range=np.array(pc_range, np.float32), voxel=np.array(voxel, np.float32),
               n_boxes=np.array([b.shape[0] for b in boxes]),
               boxes=np.concatenate(boxes), gt_hm=torch.cat(gt_hm, 1).numpy(),
               teacher_logit=torch.cat(t_logit, 1).numpy(), student_prob=torch.cat(s_prob, 1).numpy())
    for name, over in variants.items():
        params = dict(base)
        params.update(over)
        me = _fgd_self(methods, AttrDict, params, C, grid, pc_range, voxel)
        st = student.clone().requires_grad_(True)
        fg, fgs, bgs = me.foreground_scale_mask(H, H, [_Boxes(torch.from_numpy(b)) for b in boxes], 0, 0)
        teacher_preds = [[dict(heatmap=t.clone())] for t in t_logit]   # clip_sigmoid is in-place
        student_preds = [[dict(heatmap=p.clone())] for p in s_prob]
        losses = me.fgd_distill_loss(teacher.clone(), st, [_Boxes(torch.from_numpy(b)) for b in boxes],
                                     None, canvas, [h.clone() for h in gt_hm], teacher_preds,
                                     student_preds, 0)
        total = sum(losses.values())
        total.backward()
        conv = me.spatial_wise_adaptations[0]
        out.update({name + "_fg": fg.numpy(), name + "_fg_scale": fgs.numpy(), name + "_bg_scale": bgs.numpy(),
                    name + "_params": json.dumps(params), name + "_loss_keys": json.dumps(sorted(losses)),
                    name + "_loss_vals": np.array([float(losses[k]) for k in sorted(losses)], np.float64),
                    name + "_grad_student": st.grad.numpy(),
                    name + "_conv_w": conv.weight.detach().numpy().reshape(3, 3),
                    name + "_conv_b": conv.bias.detach().numpy(),
                    name + "_grad_conv_w": conv.weight.grad.numpy().reshape(3, 3),
                    name + "_grad_conv_b": conv.bias.grad.numpy()})
        if params["fp_as_foreground"][0] != "none":
            fp, fps, cnt = me.add_fp_as_fg(
                params["fp_as_foreground"][0], fg, [h.clone() for h in gt_hm],
                [[dict(heatmap=t.clone())] for t in t_logit], [[dict(heatmap=p.clone())] for p in s_prob])
            out.update({name + "_fp": fp.numpy(), name + "_fp_scale": fps.numpy(), name + "_fp_count": cnt.numpy()})
        print("fgd", name, {k: float(v) for k, v in losses.items()})
    # affinity (list branch) on foreground rows of sample 0
    me = _fgd_self(methods, AttrDict, dict(base, affinity_weights=[0.5], affinity_split=1,
                                           affinity_criterion=dict(type="SmoothL1Loss")), C, grid, pc_range, voxel)
    K = 37
    tf = [torch.randn(K, C, generator=g), torch.randn(5, C, generator=g)]
    sf = [torch.randn(K, C, generator=g), torch.randn(5, C, generator=g)]
    aff = me.affinity_distill_loss(tf, sf, 0)
    out.update(aff_t0=tf[0].numpy(), aff_t1=tf[1].numpy(), aff_s0=sf[0].numpy(), aff_s1=sf[1].numpy(),
               aff_loss=np.float64(float(aff["kd_affinity_loss"])))
    np.savez_compressed(os.path.join(GOLDEN, "fgd_small.npz"), **out)


def pillar_golden():
    """Unmodified reference DynamicPillarFeatureNet (eval) + PointPillarsScatter on a small batch."""
    from oracle import voxel_oracle
    DPFN, PPS = ref_import.pillar_modules()
    vs, pcr = [0.5, 0.5, 8.0], [-16.0, -16.0, -5.0, 16.0, 16.0, 3.0]
    torch.manual_seed(5)
    enc = DPFN(in_channels=5, feat_channels=(64,), voxel_size=vs, point_cloud_range=pcr,
               norm_cfg=dict(type="BN1d", eps=1e-3, momentum=0.01))
    bn = enc.pfn_layers[0][1]
    with torch.no_grad():
        bn.running_mean.copy_(torch.randn(64) * 0.3)
        bn.running_var.copy_(torch.rand(64) + 0.5)
        bn.weight.copy_(torch.rand(64) + 0.5)
        bn.bias.copy_(torch.randn(64) * 0.2)
    enc.eval()
    rng = np.random.RandomState(17)
    pts, coors = [], []
    for b, n in enumerate([1500, 900]):
        p = np.zeros((n, 5), dtype=np.float32)
        p[:, :2] = rng.normal(0, 7.0, (n, 2))
        p[:, 2] = rng.uniform(-4.5, 2.5, n)
        p[:, 3] = rng.uniform(0, 255, n)
        p[:, 4] = rng.randint(0, 10, n) * 0.05
        p[:30, 0] += 40.0                      # out of range -> coors -1
        c = voxel_oracle.dynamic_voxelize(p, vs, pcr)
        pts.append(p)
        coors.append(np.concatenate([np.full((n, 1), b, np.int32), c], 1))
    pts, coors = np.concatenate(pts), np.concatenate(coors)
    with torch.no_grad():
        vf, vc = enc(torch.from_numpy(pts), torch.from_numpy(coors))
        scat = PPS(64, [64, 64])
        canvas = scat(vf, vc, 2)
    np.savez_compressed(
        os.path.join(GOLDEN, "pillar_small.npz"), points=pts, coors=coors,
        voxel_size=np.array(vs, np.float32), coors_range=np.array(pcr, np.float32),
        weight=enc.pfn_layers[0][0].weight.detach().numpy(), bn_weight=bn.weight.detach().numpy(),
        bn_bias=bn.bias.detach().numpy(), bn_mean=bn.running_mean.numpy(), bn_var=bn.running_var.numpy(),
        bn_eps=np.float32(bn.eps), voxel_feats=vf.numpy(), voxel_coors=vc.numpy(),
        canvas_nonzero=np.stack(np.nonzero(canvas.numpy().sum(1))).astype(np.int32),
        canvas_checksum=np.float64(canvas.double().sum().item()))
    print("pillar_small: pillars", vf.shape[0], "canvas", tuple(canvas.shape))


if __name__ == "__main__":
    os.makedirs(GOLDEN, exist_ok=True)
    vtm = ref_import.view_transformer_mine()
    bp = ref_import.bev_pool_py()
    lss_small(vtm)
    lss_edge(vtm)
    quickcumsum(bp)
    fullsize_check(vtm)
    voxel_small()
    fgd_golden()
    pillar_golden()
