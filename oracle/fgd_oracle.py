"""CPU oracle for the DistillBEV feature-distillation loss (numpy, float64 sums).

TEST INFRASTRUCTURE ONLY — never imported by ``distill-bev_b200/``.

Restates (paths relative to the reference checkout qcraftai/distill-bev @ 3e8f6a4):
  foreground_scale_mask  mmdet3d/models/detectors/bevdet_distill.py:755-843
                         (BEVFormer variant, cell centres + float osf:
                          mmdet3d/models/detectors/bevformer_distill.py:391-482)
    points_in_rbbox      mmdet3d/core/bbox/box_np_ops.py:426-446 with
                         center_to_corner_box3d :206-235, rotation_3d_in_axis :175-203,
                         _points_in_convex_polygon_3d_jit :719-753 (sign >= 0 -> outside)
  add_fp_as_fg           mmdet3d/models/detectors/bevdet_distill.py:846-970
                         (modes of :893-903; fp_scale_mode 'average' :923-925)
  fgd_loss               mmdet3d/models/detectors/bevdet_distill.py:1084-1293
                         (attention :1084-1108, masks :1110-1168, losses :1252-1293)
  affinity_loss          mmdet3d/models/detectors/bevdet_distill.py:703-752 (list branch)
mmdet's MSELoss / L1Loss / SmoothL1Loss(reduction='none') are third-party
(mmdet==2.24.0): elementwise (p-t)^2, |p-t|, smooth_l1(beta=1) times
loss_weight (=1); parity unpinned for that dependency.

Parity pin: tools/make_golden.py executes the UNMODIFIED reference method
bodies (cut out of the class with ast, tools/ref_import.py) on seeded inputs
and stores inputs, masks, losses and autograd gradients in
tests/golden/fgd_*.npz; tests/test_oracle_fgd.py checks this file against them.

Geometry note: the reference evaluates the point-in-box plane test in float32;
this restatement uses the equivalent rotated-rectangle test in float64, so a
BEV cell whose centre lies within float32 rounding of a box edge may differ
(measure-zero; the fixtures contain no such cell).
"""
import numpy as np

F32 = np.float32


def cell_coords(n, voxel, osf, pc_min, center=False):
    """x_i = i * voxel * osf + pc_min evaluated in float32, left to right (:766-767).
    center=True adds voxel*osf/2 (bevformer_distill.py:399-400)."""
    i = np.arange(n, dtype=F32)
    x = ((i * F32(voxel)).astype(F32) * F32(osf)).astype(F32)
    x = (x + F32(pc_min)).astype(F32)
    if center:
        x = (x + (F32(voxel) * F32(osf) / F32(2))).astype(F32)
    return x


def points_in_rbbox_xy(px, py, boxes):
    """[N] x/y (float32) vs boxes [M, >=7] -> bool [N, M]. z is ignored: the caller
    flattens boxes to z in [0, 1] and tests z = 0.5 (:783-786)."""
    b = np.asarray(boxes, dtype=np.float64)
    cx, cy, w, l, yaw = b[:, 0], b[:, 1], b[:, 3], b[:, 4], b[:, 6]
    c, s = np.cos(yaw), np.sin(yaw)
    dx = px.astype(np.float64)[:, None] - cx[None]
    dy = py.astype(np.float64)[:, None] - cy[None]
    lx = dx * c[None] - dy * s[None]
    ly = dx * s[None] + dy * c[None]
    return (np.abs(lx) < w[None] / 2) & (np.abs(ly) < l[None] / 2)


def foreground_scale_mask(H, W, boxes_list, grid_size, pc_range, voxel_size,
                          transpose_mask=False, center=False, float_osf=False):
    """-> foreground_mask, fg_scale_mask, bg_scale_mask, each [B, 1, H, W] float32."""
    assert H == W
    osf = (grid_size[0] / W) if float_osf else (grid_size[0] // W)
    xs = cell_coords(W, voxel_size[0], osf, pc_range[0], center)
    ys = cell_coords(H, voxel_size[1], osf, pc_range[1], center)
    gx, gy = np.meshgrid(xs, ys, indexing="ij")          # [W, H], point index p = i*H + j
    px, py = gx.reshape(-1), gy.reshape(-1)
    num = (F32(voxel_size[0]) * F32(voxel_size[1])).astype(F32)
    num = ((num * F32(osf)).astype(F32) * F32(osf)).astype(F32)
    fgs, fss, bss = [], [], []
    for boxes in boxes_list:
        boxes = np.asarray(boxes, dtype=F32)
        n = H * W
        if boxes.shape[0] == 0:
            inside = np.zeros((n, 0), dtype=bool)
        else:
            inside = points_in_rbbox_xy(px, py, boxes)
        fg = inside.any(axis=1)
        first = np.argmax(inside, axis=1) if boxes.shape[0] else np.zeros(n, dtype=np.int64)
        fg_scale = np.zeros(n, dtype=np.float64)
        if fg.any():
            den = (boxes[first[fg], 3] * boxes[first[fg], 4]).astype(F32)
            fg_scale[fg] = np.sqrt((num / den).astype(F32)).astype(F32)
        bg_scale = np.full(n, 1.0 / (n - fg.sum()), dtype=np.float64)

        def lay(a):
            a = a.reshape(W, H)
            return (a if transpose_mask else a.T).reshape(1, 1, H, W)
        fgs.append(lay(fg.astype(np.float64)))
        fss.append(lay(fg_scale))
        bss.append(lay(bg_scale))
    return (np.concatenate(fgs).astype(F32), np.concatenate(fss).astype(F32),
            np.concatenate(bss).astype(F32))


def _to_res(a, target):
    """[B,1,S,S] -> [B,1,target,target]: max-pool when larger, repeat when smaller (:876-891)."""
    S = a.shape[2]
    if S > target:
        k = S // target
        return a.reshape(a.shape[0], 1, target, k, target, k).max(axis=(3, 5))
    if S < target:
        k = target // S
        return np.repeat(np.repeat(a, k, axis=2), k, axis=3)
    return a


def fp_dfs_scale_literal(fp, max_pops=2_000_000):
    """fp_scale_mode 'dfs' exactly as written (bevdet_distill.py:926-966): FIFO flood fill from every unvisited
    FP cell in row-major order, neighbours pushed in the order y+1, y-1, x+1, x-1, `visited` set when a cell is
    POPPED - so a cell reachable from several already-queued neighbours is queued (and later counted) more than
    once, and the component's scale is 1 / len(count) with those repeats included. fp: [B,1,H,W] -> scale."""
    B, _, H, W = fp.shape
    scale = np.zeros_like(fp, dtype=F32)
    for b in range(B):
        m = fp[b, 0] > 0
        visited = np.zeros((H, W), bool)
        for y0, x0 in zip(*np.nonzero(m)):
            if visited[y0, x0]:
                continue
            queue, count, head = [(int(y0), int(x0))], [], 0
            while head < len(queue):
                y, x = queue[head]
                head += 1
                if head > max_pops:
                    raise RuntimeError("fp_dfs_scale_literal: component too large for the literal walk")
                visited[y, x] = True
                count.append((y, x))
                if y + 1 < H and not visited[y + 1, x] and m[y + 1, x]:
                    queue.append((y + 1, x))
                if y - 1 >= 0 and not visited[y - 1, x] and m[y - 1, x]:
                    queue.append((y - 1, x))
                if x + 1 < W and not visited[y, x + 1] and m[y, x + 1]:
                    queue.append((y, x + 1))
                if x - 1 >= 0 and not visited[y, x - 1] and m[y, x - 1]:
                    queue.append((y, x - 1))
            val = F32(1.0 / len(count))
            for y, x in count:
                scale[b, 0, y, x] = val
    return scale


def fp_dfs_scale(fp):
    """Same result without the repeated work. The grid graph is bipartite, so neighbours differ by exactly one
    BFS layer and the FIFO pops a whole layer (repeats included) before the next one: a cell is queued once per
    pop of each previous-layer neighbour, pops(c) = sum of pops(n) over those neighbours, pops(seed) = 1, and
    len(count) = sum of pops over the component (exact in Python ints; the literal walk is exponential in the
    component's diameter)."""
    B, _, H, W = fp.shape
    scale = np.zeros_like(fp, dtype=F32)
    for b in range(B):
        m = fp[b, 0] > 0
        layer = -np.ones((H, W), np.int64)
        for y0, x0 in zip(*np.nonzero(m)):
            if layer[y0, x0] >= 0:
                continue
            pops = {(int(y0), int(x0)): 1}
            layer[y0, x0] = 0
            order, head = [(int(y0), int(x0))], 0
            while head < len(order):
                y, x = order[head]
                head += 1
                for yy, xx in ((y + 1, x), (y - 1, x), (y, x + 1), (y, x - 1)):
                    if 0 <= yy < H and 0 <= xx < W and m[yy, xx]:
                        if layer[yy, xx] < 0:
                            layer[yy, xx] = layer[y, x] + 1
                            pops[(yy, xx)] = 0
                            order.append((yy, xx))
                        if layer[yy, xx] == layer[y, x] + 1:
                            pops[(yy, xx)] += pops[(y, x)]
            total = sum(pops.values())
            val = F32(1.0 / total)
            for y, x in order:
                scale[b, 0, y, x] = val
    return scale


def add_fp_as_fg(mode, fg_mask, gt_hm, teacher_hm, student_hm, thres, gt_thres=None, scale_mode="average"):
    """gt_hm / teacher_hm / student_hm: [B, K, h, w] class heatmaps (teacher already
    through clip_sigmoid). -> fp_mask, fp_scale_mask [B,1,H,W] float32, fp_count [B]."""
    if gt_thres is None:
        gt_thres = thres
    g = gt_hm.max(axis=1, keepdims=True)
    t = teacher_hm.max(axis=1, keepdims=True)
    s = student_hm.max(axis=1, keepdims=True)
    T = t.shape[2]
    s, g = _to_res(s, T), _to_res(g, T)
    if mode == "teacher":
        fp = (g < gt_thres) & (t > thres)
    elif mode == "student":
        fp = (g < gt_thres) & (s > thres)
    elif mode == "teacher_selected_student":
        fp = (g < gt_thres) & (s > thres) & (t < gt_thres)
    elif mode == "teacher+teacher_selected_student":
        fp = ((g < gt_thres) & (t > thres)) | ((g < gt_thres) & (s > thres) & (t < gt_thres))
    else:
        raise NotImplementedError(mode)
    fp = _to_res(fp.astype(F32), fg_mask.shape[2]) > 0
    fp = (fp & (fg_mask == 0)).astype(F32)
    cnt = fp.sum(axis=(1, 2, 3))
    if scale_mode == "dfs":
        return fp, fp_dfs_scale(fp), cnt.astype(F32)
    if scale_mode != "average":
        raise NotImplementedError(scale_mode)
    scale = np.zeros_like(fp)
    for b in range(fp.shape[0]):
        if cnt[b] > 0:
            scale[b][fp[b] > 0] = F32(1.0) / F32(cnt[b])
    return fp, scale, cnt.astype(F32)


def _softmax(x, axis):
    x = x - x.max(axis=axis, keepdims=True)
    e = np.exp(x)
    return e / e.sum(axis=axis, keepdims=True)


def conv3x3(x, w, b):
    """nn.Conv2d(1, 1, 3, padding=1) on [B,1,H,W] (cross-correlation)."""
    B, _, H, W = x.shape
    xp = np.pad(x, ((0, 0), (0, 0), (1, 1), (1, 1)))
    out = np.full_like(x, b, dtype=np.float64)
    for dy in range(3):
        for dx in range(3):
            out += w[dy, dx] * xp[:, :, dy:dy + H, dx:dx + W]
    return out


def conv3x3_transpose(g, w):
    """gradient of conv3x3 w.r.t. its input."""
    B, _, H, W = g.shape
    gp = np.pad(g, ((0, 0), (0, 0), (1, 1), (1, 1)))
    out = np.zeros_like(g, dtype=np.float64)
    for dy in range(3):
        for dx in range(3):
            out += w[dy, dx] * gp[:, :, 2 - dy:2 - dy + H, 2 - dx:2 - dx + W]
    return out


def fgd_loss(teacher, student, fg, fg_scale, bg_scale, params, conv_w=None, conv_b=0.0,
             fp=None, fp_scale=None, fp_count=None, want_grad=False):
    """Losses of fgd_distill_loss for already-adapted features (:1084-1293).

    teacher, student [B,C,H,W]; fg/fg_scale/bg_scale [B,1,H,W] (foreground_scale_mask);
    params: spatial_t, spatial_student_ratio, channel_t, w_fg, w_bg, w_channel, w_spatial,
    w_fp, spatial_att ('teacher'|'teacher_student'), spatial_mask, channel_mask,
    scale_mask ('combine_gt'|'separate_gt'|'bg_only'|None), background_mask.
    Returns dict of float64 scalars (+ 'grad_student', 'grad_conv_w', 'grad_conv_b' of the
    SUM of all losses when want_grad).
    """
    t = np.asarray(teacher, dtype=np.float64)
    s = np.asarray(student, dtype=np.float64)
    B, C, H, W = t.shape
    HW = H * W
    S_T, C_T, r = params["spatial_t"], params["channel_t"], params["spatial_student_ratio"]
    t_att = _softmax(np.abs(t).mean(axis=1).reshape(B, -1) / S_T, 1).reshape(B, 1, H, W) * HW
    s_att = _softmax(np.abs(s).mean(axis=1).reshape(B, -1) / S_T, 1).reshape(B, 1, H, W) * HW
    c_att = _softmax(np.abs(t).mean(axis=(2, 3)) / C_T, 1).reshape(B, C, 1, 1) * C
    if params["spatial_att"] == "teacher":
        sum_att = t_att
    elif params["spatial_att"] == "teacher_student":
        sum_att = (t_att + s_att * r) / (1 + r)
    else:
        raise NotImplementedError
    fgm = np.asarray(fg, dtype=np.float64)
    fg_sc = np.asarray(fg_scale, dtype=np.float64)
    bg_sc = np.asarray(bg_scale, dtype=np.float64).copy()
    bgm = (fgm == 0).astype(np.float64) if params.get("background_mask", "logical_not") == "logical_not" \
        else 1.0 - fgm
    use_fp = fp is not None
    if use_fp:
        fpm = np.asarray(fp, dtype=np.float64)
        bgm = np.where(fpm != 0, 0.0, bgm)
        bg_pts = HW - fgm.sum(axis=(1, 2, 3))
        for b in range(B):
            bg_sc[b] = 1.0 / (bg_pts[b] - fp_count[b]) if bg_pts[b] > fp_count[b] else 0.0
    sm = params.get("scale_mask", "combine_gt")
    if sm == "combine_gt":
        sc = np.maximum(fg_sc, bg_sc)
        fg_w, bg_w = fgm * sc, bgm * sc
    elif sm == "separate_gt":
        fg_w, bg_w = fgm * fg_sc, bgm * bg_sc
    elif sm == "bg_only":
        fg_w, bg_w = fgm * bg_sc, bgm * bg_sc
    elif not sm:
        fg_w, bg_w = fgm, bgm
    else:
        raise NotImplementedError(sm)
    if params["spatial_mask"]:
        fg_w, bg_w = fg_w * sum_att, bg_w * sum_att
    if params["channel_mask"]:
        fg_w, bg_w = fg_w * c_att, bg_w * c_att
    d2 = (s - t) ** 2
    out = {}
    out["kd_fg_feat_loss"] = (d2 * fg_w).sum() * params["w_fg"] / B
    out["kd_bg_feat_loss"] = (d2 * bg_w).sum() * params["w_bg"] / B
    wtot = fg_w * params["w_fg"] / B + bg_w * params["w_bg"] / B
    grad = None
    if params["channel_mask"]:
        mt, ms = t.mean(axis=(2, 3)), s.mean(axis=(2, 3))
        out["kd_channel_loss"] = np.abs(mt - ms).sum() * params["w_channel"] / B
    if params["spatial_mask"]:
        tp, sp = t.mean(axis=1, keepdims=True), s.mean(axis=1, keepdims=True)
        o = conv3x3(sp, np.asarray(conv_w, dtype=np.float64), float(conv_b))
        out["kd_spatial_loss"] = np.abs(tp - o).sum() * params["w_spatial"] / B
    if use_fp:
        fp_w = fpm * np.asarray(fp_scale, dtype=np.float64) * sum_att * c_att
        out["kd_fp_bg_feat_loss"] = (d2 * fp_w).sum() * params["w_fp"] / B
        wtot = wtot + fp_w * params["w_fp"] / B
    if want_grad:
        grad = 2.0 * (s - t) * wtot
        if params["channel_mask"]:
            grad = grad - (np.sign(mt - ms) * params["w_channel"] / (B * HW))[:, :, None, None]
        if params["spatial_mask"]:
            go = -np.sign(tp - o) * params["w_spatial"] / B
            grad = grad + conv3x3_transpose(go, np.asarray(conv_w, dtype=np.float64)) / C
            spp = np.pad(sp, ((0, 0), (0, 0), (1, 1), (1, 1)))
            gw = np.zeros((3, 3))
            for dy in range(3):
                for dx in range(3):
                    gw[dy, dx] = (go * spp[:, :, dy:dy + H, dx:dx + W]).sum()
            out["grad_conv_w"], out["grad_conv_b"] = gw, go.sum()
        out["grad_student"] = grad
    out["_t_att"], out["_s_att"], out["_c_att"] = t_att, s_att, c_att
    return out


def affinity_loss(t_rows, s_rows, weight, perm=None, split=1, beta=1.0):
    """List branch of affinity_distill_loss (:737-748): per sample [K, C] rows ->
    mean SmoothL1 between K x K gram matrices (mmdet SmoothL1Loss default reduction 'mean')."""
    total = 0.0
    for i, (tf, sf) in enumerate(zip(t_rows, s_rows)):
        tf, sf = np.asarray(tf, dtype=np.float64), np.asarray(sf, dtype=np.float64)
        K = tf.shape[0]
        p = np.arange(K) if perm is None else np.asarray(perm[i])
        loss = 0.0
        for j in range(split):
            idx = p[j::split]
            ta, sa = tf[idx] @ tf[idx].T, sf[idx] @ sf[idx].T
            d = np.abs(ta - sa)
            l = np.where(d < beta, 0.5 * d * d / beta, d - 0.5 * beta)
            loss += (l.mean() if l.size else 0.0) * weight
        total += loss / split
    return total
