"""CPU restatement of the sparse LiDAR teacher path (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (distill-bev_b200/) never does.

Follows, function by function (paths relative to the reference checkout):
  conv_output_size      mmdet3d/ops/spconv/ops.py:20-31
  valid_out_pos         mmdet3d/ops/spconv/include/spconv/geometry.h:24-84 (getValidOutPos)
  get_indice_pairs      geometry.h:141-199 (getIndicePairsConv), :259-311 (getIndicePairsSubM),
                        driver include/spconv/spconv_ops.h:28-141 (CPU branch)
  indice_conv           spconv_ops.h:261-361 (gather -> mm -> scatter-add per kernel offset)
  dense                 mmdet3d/ops/spconv/structure.py:53-64
  sparse_conv_layer     mmdet3d/ops/spconv/conv.py:126-229 (SparseConvolution.forward)
  sparse_encoder        mmdet3d/models/middle_encoders/sparse_encoder.py:97-128,130-204,
                        mmdet3d/ops/sparse_block.py:101-121 (SparseBasicBlock.forward), :124-186
  hard_simple_vfe       mmdet3d/models/voxel_encoders/voxel_encoder.py:29-45
  dyn_voxelization(_virtual)  mmdet3d/models/voxel_encoders/dynamic_voxel_encoder.py:8-17,19-68
                        (scatter_mean: mmdet3d/core/utils/scatter.py:38-60)

Pinned by tests/golden/sparse_small.npz, produced by tools/make_golden_sparse.py from the
reference's own extension (oracle/_ref/ref_sparse_conv_ext.so, compiled unmodified) driven by the
reference's unmodified Python files. BatchNorm1d (eval) and ReLU are torch.nn semantics; mmdet
2.24's BasicBlock constructor (third party, absent) is restated in the generator — parity
unpinned at that boundary.
"""
import numpy as np


def conv_output_size(in_shape, ksize, stride, padding, dilation):
    out = []
    for i in range(len(in_shape)):
        size = (in_shape[i] + 2 * padding[i] - dilation[i] * (ksize[i] - 1) - 1) // stride[i] + 1
        out.append(1 if ksize[i] == -1 else size)
    return out


def _cdiv(a, b):
    """C integer division (truncation toward zero), as in geometry.h."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q


def valid_out_pos(pos, ksize, stride, padding, dilation, out_shape):
    """list of (out_pos tuple, kernel offset) in the order getValidOutPos emits them."""
    nd = len(pos)
    lowers = [_cdiv(pos[i] - (ksize[i] - 1) * dilation[i] - 1 + stride[i] + padding[i], stride[i])
              for i in range(nd)]
    uppers = [_cdiv(pos[i] + padding[i], stride[i]) for i in range(nd)]
    csize = [_cdiv(uppers[i] - lowers[i], dilation[i]) + 1 for i in range(nd)]
    num = 1
    for c in csize:
        num *= c
    counter = [0] * nd
    res = []
    for _ in range(max(num, 0)):
        valid, m, offset = True, 1, 0
        out = [0] * nd
        for j in range(nd - 1, -1, -1):
            val = uppers[j] - counter[j] * dilation[j]
            out[j] = val
            if val < 0 or val > out_shape[j] - 1:
                valid = False
            offset += _cdiv(m * (pos[j] - val * stride[j] + padding[j]), dilation[j])
            m *= ksize[j]
        if valid:
            res.append((tuple(out), offset))
        counter[nd - 1] += 1
        for c in range(nd - 1, 0, -1):
            if counter[c] == csize[c]:
                counter[c - 1] += 1
                counter[c] = 0
    return res


def get_indice_pairs(indices, batch_size, spatial_shape, ksize, stride, padding, dilation,
                     subm=False):
    """-> (out_indices [M,4] int32, indice_pairs [K,2,N] int32 (-1 filled), indice_num [K])
    exactly as the reference CPU branch orders them (input-major, first appearance)."""
    indices = np.asarray(indices, dtype=np.int64)
    nd = indices.shape[1] - 1
    n = indices.shape[0]
    ksize, dilation = list(ksize), list(dilation)
    if subm:
        out_shape = list(spatial_shape)
        stride = [1] * nd
        padding = [k // 2 for k in ksize]
    else:
        out_shape = conv_output_size(spatial_shape, ksize, stride, padding, dilation)
    kvol = int(np.prod(ksize))
    pairs = np.full((kvol, 2, n), -1, dtype=np.int32)
    num = np.zeros(kvol, dtype=np.int32)
    grid = {}
    if subm:
        for j in range(n):
            grid[tuple(indices[j])] = j
        for j in range(n):
            for out, off in valid_out_pos(indices[j, 1:], ksize, stride, padding, dilation, out_shape):
                key = (indices[j, 0],) + out
                if key in grid:
                    pairs[off, 0, num[off]] = j
                    pairs[off, 1, num[off]] = grid[key]
                    num[off] += 1
        return indices.astype(np.int32), pairs, num
    out_inds = []
    for j in range(n):
        for out, off in valid_out_pos(indices[j, 1:], ksize, stride, padding, dilation, out_shape):
            key = (indices[j, 0],) + out
            if key not in grid:
                grid[key] = len(out_inds)
                out_inds.append(key)
            pairs[off, 0, num[off]] = j
            pairs[off, 1, num[off]] = grid[key]
            num[off] += 1
    out_inds = np.asarray(out_inds, dtype=np.int32).reshape(-1, nd + 1)
    return out_inds, pairs, num


def indice_conv(features, filters, pairs, num, n_out):
    """filters [*k, Cin, Cout]; fp64 accumulation of out[o] += in[i] @ W[k]."""
    features = np.asarray(features, dtype=np.float64)
    cin, cout = filters.shape[-2], filters.shape[-1]
    w = np.asarray(filters, dtype=np.float64).reshape(-1, cin, cout)
    out = np.zeros((n_out, cout))
    for k in range(w.shape[0]):
        h = int(num[k])
        if h <= 0:
            continue
        np.add.at(out, pairs[k, 1, :h], features[pairs[k, 0, :h]] @ w[k])
    return out


def dense(features, indices, spatial_shape, batch_size):
    """[B, C, *spatial] (channels first), structure.py:53-64."""
    c = features.shape[1]
    res = np.zeros([batch_size] + list(spatial_shape) + [c], dtype=features.dtype)
    idx = np.asarray(indices, dtype=np.int64)
    res[tuple(idx[:, i] for i in range(idx.shape[1]))] = features
    nd = len(spatial_shape)
    return np.ascontiguousarray(res.transpose([0, nd + 1] + list(range(1, nd + 1))))


def _bn_eval(x, bn):
    return (x - bn["mean"]) / np.sqrt(bn["var"] + bn["eps"]) * bn["weight"] + bn["bias"]


def _triple(v):
    return list(v) if isinstance(v, (list, tuple)) else [v] * 3


class SparseTensor(object):
    def __init__(self, features, indices, spatial_shape, batch_size):
        self.features, self.indices = features, indices
        self.spatial_shape, self.batch_size = list(spatial_shape), batch_size
        self.indice_dict = {}


def sparse_conv_layer(x, weight, ksize, stride, padding, subm, indice_key=None, bias=None):
    """SparseConvolution.forward (conv.py:126-229) for the non-transposed, non-1x1 case."""
    ksize, stride, padding = _triple(ksize), _triple(stride), _triple(padding)
    dilation = [1, 1, 1]
    out_shape = x.spatial_shape if subm else conv_output_size(x.spatial_shape, ksize, stride,
                                                              padding, dilation)
    datas = x.indice_dict.get(indice_key) if indice_key is not None else None
    if datas is not None:
        outids, pairs, num = datas
    else:
        outids, pairs, num = get_indice_pairs(x.indices, x.batch_size, x.spatial_shape, ksize,
                                              stride, padding, dilation, subm)
        x.indice_dict[indice_key] = (outids, pairs, num)
    feats = indice_conv(x.features, weight, pairs, num, outids.shape[0])
    if bias is not None:
        feats = feats + bias
    out = SparseTensor(feats, outids, out_shape, x.batch_size)
    out.indice_dict = x.indice_dict
    return out


def sparse_encoder(layers, voxel_features, coors, batch_size, sparse_shape):
    """Run a flattened SparseEncoder. `layers` is the list tools/make_golden_sparse.py /
    the plugin's export_layers() produce: dicts with
      kind 'conv'  : weight, ksize, stride, padding, subm, indice_key, bn (dict or None), relu
      kind 'block' : conv1 / conv2 (as above, without relu) — SparseBasicBlock
    Returns the dense [B, C*D, H, W] tensor (sparse_encoder.py:121-127) and the last sparse tensor."""
    x = SparseTensor(np.asarray(voxel_features, dtype=np.float64), np.asarray(coors, dtype=np.int32),
                     sparse_shape, batch_size)

    def conv_bn(x, l, relu):
        y = sparse_conv_layer(x, l["weight"], l["ksize"], l["stride"], l["padding"], l["subm"],
                              l.get("indice_key"))
        if l.get("bn") is not None:
            y.features = _bn_eval(y.features, l["bn"])
        if relu:
            y.features = np.maximum(y.features, 0.0)
        return y

    for l in layers:
        if l["kind"] == "conv":
            x = conv_bn(x, l, l.get("relu", True))
        else:
            identity = x.features
            y = conv_bn(x, l["conv1"], True)
            y = conv_bn(y, l["conv2"], False)
            y.features = np.maximum(y.features + identity, 0.0)
            x = y
    d = dense(x.features, x.indices, x.spatial_shape, batch_size)
    n, c, dd, h, w = d.shape
    return d.reshape(n, c * dd, h, w), x


def hard_simple_vfe(features, num_points, num_features):
    f = np.asarray(features, dtype=np.float64)
    return f[:, :, :num_features].sum(axis=1) / np.asarray(num_points, dtype=np.float64)[:, None]


def _scatter_mean(rows, inv, m):
    out = np.zeros((m, rows.shape[1]))
    np.add.at(out, inv, rows.astype(np.float64))
    cnt = np.maximum(np.bincount(inv, minlength=m), 1)
    return out / cnt[:, None]


def _keep(points, pc_range):
    p = points
    return ((p[:, 0] >= pc_range[0]) & (p[:, 0] <= pc_range[3]) & (p[:, 1] >= pc_range[1]) &
            (p[:, 1] <= pc_range[4]) & (p[:, 2] >= pc_range[2]) & (p[:, 2] <= pc_range[5]))


def _coords(xyz, pc_range, voxel_size):
    pc_range, voxel_size = np.asarray(pc_range, np.float32), np.asarray(voxel_size, np.float32)
    q = (xyz[:, [2, 1, 0]].astype(np.float32) - pc_range[[2, 1, 0]]) / voxel_size[[2, 1, 0]]
    return q.astype(np.int64)  # truncation toward zero, like .to(torch.int64)


def dyn_voxelization(points, pc_range, voxel_size):
    points = np.asarray(points, dtype=np.float32)
    points = points[_keep(points, np.asarray(pc_range, np.float32))]
    coords = _coords(points[:, :3], pc_range, voxel_size)
    uniq, inv = np.unique(coords, axis=0, return_inverse=True)
    return _scatter_mean(points, inv.reshape(-1), uniq.shape[0]), uniq


def dyn_voxelization_virtual(points, pc_range, voxel_size):
    points = np.asarray(points, dtype=np.float32)
    points = points[_keep(points, np.asarray(pc_range, np.float32))]
    real = points[points[:, -2] == 1][:, [0, 1, 2, 3, 4, -1]]
    painted = points[points[:, -2] == 0]
    virtual = points[points[:, -2] == -1]
    n = len(points)
    padded = np.zeros((n, 24), dtype=np.float32)
    r, p = len(real), len(painted)
    padded[:r, :6] = real
    padded[:r, -1] = 1
    padded[r:r + p, 6:21] = painted[:, :-2]
    padded[r:r + p, 21] = painted[:, -2]
    padded[r:r + p, 22] = 1
    padded[r + p:, 6:21] = virtual[:, :-2]
    padded[r + p:, 21] = virtual[:, -2]
    xyz = np.concatenate([real[:, :3], painted[:, :3], virtual[:, :3]], axis=0)
    coords = _coords(xyz, pc_range, voxel_size)
    uniq, inv = np.unique(coords, axis=0, return_inverse=True)
    vox = _scatter_mean(padded, inv.reshape(-1), uniq.shape[0])
    ind = vox[:, -1].copy()
    mix = (ind > 0) & (ind < 1)
    vox = vox[:, :-1]
    vox[mix, :6] = vox[mix, :6] / ind[mix, None]
    vox[mix, 6:] = vox[mix, 6:] / (1 - ind[mix, None])
    return vox, uniq


def dynamic_voxel_encoder(points_list, pc_range, voxel_size, virtual=False):
    """DynamicVoxelEncoder.forward (:83-102): per-sample voxelization, batch index prepended."""
    fn = dyn_voxelization_virtual if virtual else dyn_voxelization
    vs, cs = [], []
    for i, pts in enumerate(points_list):
        v, c = fn(pts, pc_range, voxel_size)
        vs.append(v)
        cs.append(np.concatenate([np.full((len(c), 1), i, dtype=np.int64), c], axis=1))
    pr, vsz = np.asarray(pc_range, np.float32), np.asarray(voxel_size, np.float32)
    shape = np.round((pr[3:] - pr[:3]) / vsz).astype(np.int32)
    return np.concatenate(vs, 0), np.concatenate(cs, 0), shape


# ---------------------------------------------------------------------------------------------
# Flattened description of a SparseEncoder and seeded parameters for it (shared by the golden
# generator and the tests, so the fixtures need not store megabytes of weights).
# ---------------------------------------------------------------------------------------------
def encoder_layer_specs(in_channels, base_channels=16, output_channels=128,
                        encoder_channels=((16,), (32, 32, 32), (64, 64, 64), (64, 64, 64)),
                        encoder_paddings=((1,), (1, 1, 1), (1, 1, 1), ((0, 1, 1), 1, 1)),
                        block_type="conv_module"):
    """Layer list in execution order, following SparseEncoder.__init__ / make_encoder_layers
    (sparse_encoder.py:60-93,130-204) for order=('conv','norm','act')."""
    def conv(cin, cout, ksize, stride, padding, subm, key):
        return dict(kind="conv", cin=cin, cout=cout, ksize=_triple(ksize), stride=_triple(stride),
                    padding=_triple(padding), subm=subm, indice_key=key, relu=True, has_bn=True)

    specs = [conv(in_channels, base_channels, 3, 1, 1, True, "subm1")]
    cin = base_channels
    n_stage = len(encoder_channels)
    for i, blocks in enumerate(encoder_channels):
        for j, cout in enumerate(tuple(blocks)):
            padding = tuple(encoder_paddings[i])[j]
            if i != 0 and j == 0 and block_type == "conv_module":
                specs.append(conv(cin, cout, 3, 2, padding, False, "spconv%d" % (i + 1)))
            elif block_type == "basicblock":
                if j == len(blocks) - 1 and i != n_stage - 1:
                    specs.append(conv(cin, cout, 3, 2, padding, False, "spconv%d" % (i + 1)))
                else:
                    c1 = conv(cout, cout, 3, 1, 1, True, None)
                    c2 = conv(cout, cout, 3, 1, 1, True, None)
                    specs.append(dict(kind="block", conv1=c1, conv2=c2))
            else:
                specs.append(conv(cin, cout, 3, 1, padding, True, "subm%d" % (i + 1)))
            cin = cout
    specs.append(conv(cin, output_channels, (3, 1, 1), (2, 1, 1), 0, False, "spconv_down2"))
    return specs


def _iter_convs(specs):
    for l in specs:
        if l["kind"] == "conv":
            yield l
        else:
            yield l["conv1"]
            yield l["conv2"]


def fill_params(specs, seed, eps=1e-3):
    """Seeded weights [kz,ky,kx,cin,cout] (fp32) and eval-BN statistics for every conv."""
    rs = np.random.RandomState(seed)
    for c in _iter_convs(specs):
        kvol = int(np.prod(c["ksize"]))
        scale = 1.0 / np.sqrt(c["cin"] * max(kvol / 4.0, 1.0))
        c["weight"] = (rs.standard_normal(tuple(c["ksize"]) + (c["cin"], c["cout"])) * scale).astype(np.float32)
        c["bn"] = dict(weight=rs.uniform(0.5, 1.5, c["cout"]).astype(np.float32),
                       bias=(rs.standard_normal(c["cout"]) * 0.1).astype(np.float32),
                       mean=(rs.standard_normal(c["cout"]) * 0.1).astype(np.float32),
                       var=rs.uniform(0.5, 1.5, c["cout"]).astype(np.float32), eps=eps)
    return specs
