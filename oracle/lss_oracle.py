"""CPU oracle for the LSS view transform / bev_pool path (numpy).

TEST INFRASTRUCTURE ONLY. Nothing under ``distill-bev_b200/`` may import this
module; it is the checker used by ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.

It restates, function by function, the reference algorithm (paths relative to
the reference checkout qcraftai/distill-bev @ 3e8f6a4):

  gen_dx_bx          mmdet3d/models/necks/view_transformer_mine.py:14-18
  create_frustum     mmdet3d/models/necks/view_transformer_mine.py:98-109
  get_geometry       mmdet3d/models/necks/view_transformer_mine.py:111-139
  voxel_indices      mmdet3d/models/necks/view_transformer_mine.py:150-161
  voxel_pooling      mmdet3d/models/necks/view_transformer_mine.py:141-181
                     (= voxel_pooling_accelerated :184-240, same result)
  bev_pool           mmdet3d/ops/bev_pool/bev_pool.py:83-97 with
                     bev_pool_kernel mmdet3d/ops/bev_pool/src/bev_pool_cuda.cu:20-42
  bev_pool_backward  mmdet3d/ops/bev_pool/src/bev_pool_cuda.cu:61-84 and
                     QuickCumsum.backward view_transformer_mine.py:48-56

Parity pin: the reference ships no tests or golden vectors, so
this oracle is pinned against outputs of the reference's own Python code
executed in the build container: ``tools/make_golden.py`` runs the unmodified
``view_transformer_mine.py`` / ``bev_pool.py`` (import stubs only) on seeded
inputs and commits them under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this file against them.

Numerics: the reference's shipped cumsum path is itself lossy (fp32 prefix
sums; up to 2.5e-3 relative vs. a direct sum). The oracle
defines the result as the DIRECT per-cell sum accumulated in float64 and
rounded to float32; integer outputs (indices, kept mask, counts) are exact.
"""
import numpy as np


def gen_dx_bx(xbound, ybound, zbound):
    """view_transformer_mine.py:14-18 — float32 triples, as torch.Tensor() stores them."""
    rows = [xbound, ybound, zbound]
    dx = np.array([r[2] for r in rows], dtype=np.float32)
    bx = np.array([r[0] + r[2] / 2.0 for r in rows], dtype=np.float32)
    nx = np.array([(r[1] - r[0]) / r[2] for r in rows], dtype=np.float32)
    return dx, bx, nx


def create_frustum(input_size, downsample, dbound):
    """view_transformer_mine.py:98-109 -> [D, fH, fW, 3] float32 (x_px, y_px, depth)."""
    ogfH, ogfW = input_size
    fH, fW = ogfH // downsample, ogfW // downsample
    # torch.arange(*dbound, dtype=float): start + i*step computed in float64 then cast
    n = int(np.ceil((dbound[1] - dbound[0]) / dbound[2]))
    ds = (dbound[0] + np.arange(n, dtype=np.float64) * dbound[2]).astype(np.float32)
    D = ds.shape[0]
    xs = _torch_linspace(0, ogfW - 1, fW)
    ys = _torch_linspace(0, ogfH - 1, fH)
    fr = np.empty((D, fH, fW, 3), dtype=np.float32)
    fr[..., 0] = xs[None, None, :]
    fr[..., 1] = ys[None, :, None]
    fr[..., 2] = ds[:, None, None]
    return fr


def _torch_linspace(start, end, steps):
    """torch.linspace(float32) on CPU: symmetric evaluation from both ends, step in fp32."""
    if steps == 1:
        return np.array([start], dtype=np.float32)
    start32, end32 = np.float32(start), np.float32(end)
    step = np.float32((end32 - start32) / np.float32(steps - 1))
    out = np.empty(steps, dtype=np.float32)
    half = steps // 2
    idx = np.arange(steps, dtype=np.float32)
    out[:half] = start32 + step * idx[:half]
    out[half:] = end32 - step * (np.float32(steps - 1) - idx[half:])
    return out


def get_geometry(frustum, rots, trans, intrins, post_rots, post_trans):
    """view_transformer_mine.py:111-139 -> [B, N, D, fH, fW, 3] float32.

    float32 throughout like the reference; the 3x3 inverses go through float64
    and are rounded (torch.inverse uses LAPACK getrf/getri in float32, so the
    last bits of this restatement can differ: compare with a tolerance).
    """
    B, N, _ = trans.shape
    f32 = np.float32
    points = frustum[None, None].astype(f32) - post_trans.reshape(B, N, 1, 1, 1, 3).astype(f32)
    inv_post = np.linalg.inv(post_rots.astype(np.float64)).astype(f32)
    points = np.einsum("bnij,bndhwj->bndhwi", inv_post, points).astype(f32)
    points = np.concatenate((points[..., :2] * points[..., 2:3], points[..., 2:3]), axis=-1)
    combine = np.matmul(rots.astype(f32), np.linalg.inv(intrins.astype(np.float64)).astype(f32))
    points = np.einsum("bnij,bndhwj->bndhwi", combine.astype(f32), points).astype(f32)
    points = points + trans.reshape(B, N, 1, 1, 1, 3).astype(f32)
    return points.astype(f32)


def voxel_indices(geom, bx, dx, nx):
    """view_transformer_mine.py:150-161.

    geom [..., 3] float32 -> (idx int64 [n, 3], kept bool [n]). float32
    subtract, float32 divide, truncation toward zero (``.long()``), then the
    in-bounds test against the FLOAT nx.
    """
    g = np.asarray(geom, dtype=np.float32).reshape(-1, 3)
    bx = np.asarray(bx, dtype=np.float32)
    dx = np.asarray(dx, dtype=np.float32)
    nx = np.asarray(nx, dtype=np.float32)
    off = (bx - dx / np.float32(2.0)).astype(np.float32)
    q = ((g - off).astype(np.float32) / dx).astype(np.float32)
    with np.errstate(invalid="ignore"):
        idx = np.trunc(q)
    bad = ~np.isfinite(idx)
    idx = np.where(bad, -1, idx).astype(np.int64)
    kept = np.ones(g.shape[0], dtype=bool)
    for a in range(3):
        kept &= (idx[:, a] >= 0) & (idx[:, a].astype(np.float32) < nx[a])
    kept &= ~bad.any(axis=1)
    return idx, kept


def voxel_pooling(geom, x, bx, dx, nx):
    """view_transformer_mine.py:141-181 -> [B, C*nz, ny, nx] float32.

    geom [B, N, D, H, W, 3], x [B, N, D, H, W, C]. Direct per-cell sum in
    float64 (see module docstring), channel index iz*C + c after
    ``cat(final.unbind(dim=2), 1)``.
    """
    B = x.shape[0]
    C = x.shape[-1]
    nprime = int(np.prod(x.shape[:-1]))
    xf = np.asarray(x, dtype=np.float32).reshape(nprime, C)
    idx, kept = voxel_indices(geom, bx, dx, nx)
    n_i = np.asarray(nx, dtype=np.float32).astype(np.int64)  # nx.to(torch.long)
    batch_ix = np.repeat(np.arange(B, dtype=np.int64), nprime // B)
    # the float test can admit an index the integer canvas does not have
    kept = kept & (idx[:, 0] < n_i[0]) & (idx[:, 1] < n_i[1]) & (idx[:, 2] < n_i[2])
    final = np.zeros((B, n_i[2], n_i[1], n_i[0], C), dtype=np.float64)
    k = np.nonzero(kept)[0]
    np.add.at(final, (batch_ix[k], idx[k, 2], idx[k, 1], idx[k, 0]), xf[k].astype(np.float64))
    # [B, nz, ny, nx, C] -> [B, nz, C, ny, nx] -> [B, nz*C, ny, nx]
    out = final.transpose(0, 1, 4, 2, 3).reshape(B, n_i[2] * C, n_i[1], n_i[0])
    return out.astype(np.float32)


def lift(depth, feat, B, N):
    """view_transformer_mine.py:333-335 (= bevdet_distill_more.py:413-416): outer product of
    the depth distribution [BN, D, fH, fW] and image features [BN, C, fH, fW], permuted to
    channels-last -> [B, N, D, fH, fW, C] float32."""
    d = np.asarray(depth, dtype=np.float32)
    f = np.asarray(feat, dtype=np.float32)
    vol = d[:, None] * f[:, :, None]                       # [BN, C, D, fH, fW]
    BN, C, D, fH, fW = vol.shape
    return vol.reshape(B, N, C, D, fH, fW).transpose(0, 1, 3, 4, 5, 2)


def lift_splat(geom, depth, feat, B, N, bx, dx, nx):
    """lift followed by voxel_pooling."""
    return voxel_pooling(geom, lift(depth, feat, B, N), bx, dx, nx)


def lift_splat_backward(geom, depth, feat, out_grad, B, N, bx, dx, nx):
    """Gradients of lift_splat w.r.t. depth and feat (float64 accumulation)."""
    d = np.asarray(depth, dtype=np.float64)
    f = np.asarray(feat, dtype=np.float64)
    BN, C, fH, fW = f.shape
    D = d.shape[1]
    gx = voxel_pooling_backward(geom, out_grad, C, bx, dx, nx).astype(np.float64)  # [Nprime, C]
    gx = gx.reshape(BN, D, fH, fW, C)
    d_depth = np.einsum("bdhwc,bchw->bdhw", gx, f)
    d_feat = np.einsum("bdhwc,bdhw->bchw", gx, d)
    return d_depth.astype(np.float32), d_feat.astype(np.float32)


def voxel_pooling_backward(geom, out_grad, C, bx, dx, nx):
    """Gradient of voxel_pooling w.r.t. x: every kept point receives its cell's
    gradient row, dropped points receive zero (x[kept] indexing :160 and
    QuickCumsum.backward :48-56). out_grad [B, C*nz, ny, nx] -> [Nprime, C]."""
    B = out_grad.shape[0]
    idx, kept = voxel_indices(geom, bx, dx, nx)
    n_i = np.asarray(nx, dtype=np.float32).astype(np.int64)
    kept = kept & (idx[:, 0] < n_i[0]) & (idx[:, 1] < n_i[1]) & (idx[:, 2] < n_i[2])
    nprime = idx.shape[0]
    batch_ix = np.repeat(np.arange(B, dtype=np.int64), nprime // B)
    g = np.asarray(out_grad, dtype=np.float32).reshape(B, n_i[2], C, n_i[1], n_i[0])
    g = g.transpose(0, 1, 3, 4, 2)  # [B, nz, ny, nx, C]
    xg = np.zeros((nprime, C), dtype=np.float32)
    k = np.nonzero(kept)[0]
    xg[k] = g[batch_ix[k], idx[k, 2], idx[k, 1], idx[k, 0]]
    return xg


def bev_pool(feats, coords, B, D, H, W):
    """bev_pool.py:83-97 + bev_pool_kernel: feats [n, C], coords [n, 4] =
    (c0 < H, c1 < W, c2 < D, b < B) -> [B, C, D, H, W] float32 (after the
    reference's permute(0, 4, 1, 2, 3))."""
    feats = np.asarray(feats, dtype=np.float32)
    coords = np.asarray(coords).astype(np.int64)
    C = feats.shape[1]
    out = np.zeros((B, D, H, W, C), dtype=np.float64)
    np.add.at(out, (coords[:, 3], coords[:, 2], coords[:, 0], coords[:, 1]),
              feats.astype(np.float64))
    return out.transpose(0, 4, 1, 2, 3).astype(np.float32)


def bev_pool_backward(out_grad, coords):
    """bev_pool_grad_kernel: x_grad[i, :] = out_grad[b, :, z, c0, c1]."""
    g = np.asarray(out_grad, dtype=np.float32)
    coords = np.asarray(coords).astype(np.int64)
    return g[coords[:, 3], :, coords[:, 2], coords[:, 0], coords[:, 1]].astype(np.float32)


def sorted_intervals(coords, B, D, H, W):
    """The host prelude of bev_pool.py:86-93,40-46: ranks, (stable) argsort,
    interval starts / lengths. Returns (order, ranks_sorted, starts, lengths)."""
    coords = np.asarray(coords).astype(np.int64)
    ranks = coords[:, 0] * (W * D * B) + coords[:, 1] * (D * B) + coords[:, 2] * B + coords[:, 3]
    order = np.argsort(ranks, kind="stable")
    rs = ranks[order]
    kept = np.ones(rs.shape[0], dtype=bool)
    kept[1:] = rs[1:] != rs[:-1]
    starts = np.nonzero(kept)[0].astype(np.int32)
    lengths = np.empty_like(starts)
    if starts.size:
        lengths[:-1] = starts[1:] - starts[:-1]
        lengths[-1] = rs.shape[0] - starts[-1]
    return order, rs, starts, lengths


def bev_pool_interval_forward(x_sorted, geom_sorted, starts, lengths, b, d, h, w):
    """bev_pool_kernel (bev_pool_cuda.cu:20-42) on pre-sorted rows -> [b,d,h,w,c]."""
    x_sorted = np.asarray(x_sorted, dtype=np.float32)
    c = x_sorted.shape[1]
    out = np.zeros((b, d, h, w, c), dtype=np.float32)
    if len(starts) == 0:
        return out
    sums = np.add.reduceat(x_sorted.astype(np.float64), np.asarray(starts, dtype=np.int64), axis=0)
    g = np.asarray(geom_sorted)[np.asarray(starts, dtype=np.int64)]
    out[g[:, 3], g[:, 2], g[:, 0], g[:, 1]] = sums.astype(np.float32)
    return out


def bev_pool_interval_backward(out_grad, geom_sorted, starts, lengths, n):
    """bev_pool_grad_kernel (bev_pool_cuda.cu:61-84) -> x_grad [n, c]."""
    og = np.asarray(out_grad, dtype=np.float32)
    c = og.shape[-1]
    xg = np.zeros((n, c), dtype=np.float32)
    g = np.asarray(geom_sorted)
    for s, l in zip(starts, lengths):
        xg[s:s + l] = og[g[s, 3], g[s, 2], g[s, 0], g[s, 1]]
    return xg
