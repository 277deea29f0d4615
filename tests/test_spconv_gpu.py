"""GPU parity: sparse 3D convolution path (csrc/spconv.cu through the C-ABI shims) vs
tests/golden/sparse_small.npz (the reference's own extension + Python modules) and vs the oracle.

Bars: rulebooks are compared as integers — output coordinate sets identical and in lexicographic
(b,z,y,x) order (the reference's CUDA-branch order; its CPU branch is first-appearance, so rows are
matched by coordinate), pair sets per kernel offset identical. Features: rtol 1e-4 (fp32 sums in a
different but fixed order)."""
import os

import numpy as np
import pytest
import torch

import distill_bev_b200 as dbev
from distill_bev_b200.plugin.ops import spconv as sp
from oracle import spconv_oracle as so
from test_oracle_sparse import ENC

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "sparse_small.npz"))


def _t(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _geom(g, name):
    v = g["rb_%s_geom" % name].tolist()
    return v[0:3], v[3:6], v[6:9], v[9:12], bool(v[12])


def _lex(c):
    return np.lexsort((c[:, 3], c[:, 2], c[:, 1], c[:, 0]))


def test_rulebooks_match_reference(g, cuda):
    for name in g["rb_names"]:
        shape, k, s, p, subm = _geom(g, name)
        coors = g["rb_%s_coors" % name]
        outids, pairs, num = sp.get_indice_pairs(_t(coors, cuda), 2, shape, k, s, p, 1, 0, subm)
        outids, pairs, num = outids.cpu().numpy(), pairs.cpu().numpy(), num.cpu().numpy()
        ref_out, ref_pairs, ref_num = g["rb_%s_outids" % name], g["rb_%s_pairs" % name], g["rb_%s_num" % name]
        assert pairs.shape == ref_pairs.shape, name
        assert np.array_equal(num, ref_num), name
        if subm:
            assert np.array_equal(outids, coors), name
            remap = np.arange(len(coors))
        else:
            order = _lex(ref_out)
            assert np.array_equal(outids, ref_out[order]), name     # same set, lexicographic order
            remap = np.empty(len(order), dtype=np.int64)
            remap[order] = np.arange(len(order))                    # reference row -> our row
        for kk in range(pairs.shape[0]):
            h = int(num[kk])
            ours = set(zip(pairs[kk, 0, :h].tolist(), pairs[kk, 1, :h].tolist()))
            ref = set(zip(ref_pairs[kk, 0, :h].tolist(), remap[ref_pairs[kk, 1, :h]].tolist()))
            assert ours == ref, (name, kk)
            assert (pairs[kk, :, h:] == -1).all(), (name, kk)
            assert (np.diff(pairs[kk, 1, :h]) > 0).all(), (name, kk)   # ascending output row


def test_indice_conv_on_reference_rulebook(g, cuda):
    """The reference's own (indice_pairs, indice_pair_num) fed to our indice_conv."""
    for name in g["rb_names"]:
        _, _, _, _, subm = _geom(g, name)
        y = sp.indice_conv(_t(g["rb_%s_feats" % name], cuda), _t(g["rb_%s_w" % name], cuda),
                           _t(g["rb_%s_pairs" % name], cuda), _t(g["rb_%s_num" % name], cuda),
                           len(g["rb_%s_outids" % name]), False, subm)
        ref = g["rb_%s_y" % name]
        np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=1e-4, atol=1e-5 * np.abs(ref).max())


def test_conv_on_own_rulebook_and_epilogue(g, cuda):
    name = "s2p1"
    shape, k, s, p, subm = _geom(g, name)
    coors, feats, w = g["rb_%s_coors" % name], g["rb_%s_feats" % name], g["rb_%s_w" % name]
    rb = sp.build_rulebook(_t(coors, cuda), 2, shape, k, s, p, 1, subm)
    ref_out, ref = g["rb_%s_outids" % name], g["rb_%s_y" % name]
    order = _lex(ref_out)
    y = sp.conv_table(_t(feats, cuda), _t(w, cuda), rb.nbr, rb.n_out).cpu().numpy()
    np.testing.assert_allclose(y, ref[order], rtol=1e-4, atol=1e-5 * np.abs(ref).max())
    rs = np.random.RandomState(0)
    scale, shift = rs.uniform(0.5, 1.5, 16).astype(np.float32), rs.standard_normal(16).astype(np.float32)
    res = rs.standard_normal(y.shape).astype(np.float32)
    y2 = sp.conv_table(_t(feats, cuda), _t(w, cuda), rb.nbr, rb.n_out, _t(scale, cuda), _t(shift, cuda),
                       _t(res, cuda), True).cpu().numpy()
    np.testing.assert_allclose(y2, np.maximum(ref[order] * scale + shift + res, 0), rtol=1e-4, atol=1e-4)


def _load_params(enc, specs):
    convs = [m for m in enc.modules() if isinstance(m, sp.SparseConvolution)]
    bns = [m for m in enc.modules() if isinstance(m, torch.nn.BatchNorm1d)]
    flat = list(so._iter_convs(specs))
    assert len(convs) == len(flat) == len(bns)
    with torch.no_grad():
        for m, bn, c in zip(convs, bns, flat):
            assert tuple(m.weight.shape) == c["weight"].shape
            m.weight.copy_(torch.from_numpy(c["weight"]))
            bn.weight.copy_(torch.from_numpy(c["bn"]["weight"]))
            bn.bias.copy_(torch.from_numpy(c["bn"]["bias"]))
            bn.running_mean.copy_(torch.from_numpy(c["bn"]["mean"]))
            bn.running_var.copy_(torch.from_numpy(c["bn"]["var"]))


@pytest.mark.parametrize("tag", ["lf", "sec"])
def test_sparse_encoder_matches_reference(g, cuda, tag):
    cfg = ENC[tag]
    enc = dbev.SparseEncoder(**cfg)
    specs = so.fill_params(so.encoder_layer_specs(cfg["in_channels"], 16, cfg["output_channels"],
                                                  cfg["encoder_channels"], cfg["encoder_paddings"],
                                                  cfg["block_type"]), seed=11)
    _load_params(enc, specs)
    enc = enc.to(cuda).eval()
    y = enc(_t(g["enc_%s_feats" % tag], cuda), _t(g["enc_%s_coors" % tag], cuda), 2).cpu().numpy()
    ref = g["enc_%s_out" % tag]
    assert y.shape == ref.shape
    assert np.array_equal(y != 0, ref != 0) or np.abs(y - ref).max() < 1e-4 * np.abs(ref).max()
    np.testing.assert_allclose(y, ref, rtol=1e-3, atol=1e-4 * np.abs(ref).max())


def test_train_mode_bn_path_equals_fused(g, cuda):
    """Unfused path (conv kernel, then torch BatchNorm1d / ReLU on the features) == fused epilogue."""
    tag = "sec"
    cfg = ENC[tag]
    enc = dbev.SparseEncoder(**cfg)
    _load_params(enc, so.fill_params(so.encoder_layer_specs(
        cfg["in_channels"], 16, cfg["output_channels"], cfg["encoder_channels"],
        cfg["encoder_paddings"], cfg["block_type"]), seed=11))
    enc = enc.to(cuda).eval()
    feats, coors = _t(g["enc_%s_feats" % tag], cuda), _t(g["enc_%s_coors" % tag], cuda)
    x = sp.SparseConvTensor(feats, coors, cfg["sparse_shape"], 2)
    conv, bn, relu = enc.conv_input[0], enc.conv_input[1], enc.conv_input[2]
    with torch.no_grad():
        fused = enc.conv_input(x)
        plain = conv(sp.SparseConvTensor(feats, coors, cfg["sparse_shape"], 2))
        want = relu(bn(plain.features.clone()))
    np.testing.assert_allclose(fused.features.cpu().numpy(), want.cpu().numpy(), rtol=1e-5, atol=1e-5)


def test_dense_matches_scatter(cuda):
    rs = np.random.RandomState(3)
    B, Z, Y, X, C = 2, 3, 17, 19, 32
    lin = rs.permutation(B * Z * Y * X)[:700]
    coors = np.stack(np.unravel_index(lin, (B, Z, Y, X)), 1).astype(np.int32)
    feats = rs.standard_normal((700, C)).astype(np.float32)
    d = sp.dense_from_sparse(_t(feats, cuda), _t(coors, cuda), [Z, Y, X], B).cpu().numpy()
    ref = so.dense(feats, coors, [Z, Y, X], B).reshape(B, C * Z, Y, X)
    assert np.array_equal(d, ref)


def test_full_size_lidarformer_grid(cuda):
    """LidarFormer geometry (41 x 1600 x 1600, configs/teacher_transformer/lidarformer.py:45): checks
    that do not need the CPU oracle at this size — output set == torch.unique of the candidate
    cells, submanifold centre column is the identity, linearity of the conv, and determinism."""
    rs = np.random.RandomState(5)
    B, shape, n = 2, [41, 1600, 1600], 60000
    coors = []
    for b in range(B):
        xy = np.clip(rs.standard_normal((n, 2)) * 250 + 800, 0, 1599).astype(np.int64)
        z = rs.randint(0, 41, n)
        c = np.unique(np.stack([z, xy[:, 0], xy[:, 1]], 1), axis=0)
        c = c[rs.permutation(len(c))]
        coors.append(np.concatenate([np.full((len(c), 1), b), c], 1))
    coors = np.concatenate(coors, 0).astype(np.int32)
    ct = _t(coors, cuda)
    rb = sp.build_rulebook(ct, B, shape, 3, 1, 1, 1, True)
    assert torch.equal(rb.nbr[13], torch.arange(len(coors), device=cuda, dtype=torch.int32))
    # symmetry of a submanifold rulebook: o sees i through k  <=>  i sees o through 26 - k
    nbr = rb.nbr.cpu().numpy()
    for k in (0, 5, 12):
        o = np.nonzero(nbr[k] >= 0)[0]
        assert np.array_equal(nbr[26 - k][nbr[k][o]], o)
    rb2 = sp.build_rulebook(ct, B, shape, 3, 2, 1, 1, False)
    c64 = ct.long()
    cand = []
    for kz in range(3):
        for ky in range(3):
            for kx in range(3):
                num = c64[:, 1:] + 1 - torch.tensor([kz, ky, kx], device=cuda)
                ok = ((num >= 0) & (num % 2 == 0)).all(1)
                o = num // 2
                ok &= (o[:, 0] < 21) & (o[:, 1] < 800) & (o[:, 2] < 800)
                cand.append(((c64[ok, 0] * 21 + o[ok, 0]) * 800 + o[ok, 1]) * 800 + o[ok, 2])
    uniq = torch.unique(torch.cat(cand))
    oi = rb2.out_indices.long()
    ours = ((oi[:, 0] * 21 + oi[:, 1]) * 800 + oi[:, 2]) * 800 + oi[:, 3]
    assert torch.equal(ours, uniq)
    assert rb2.out_shape == [21, 800, 800]
    w = _t(rs.standard_normal((3, 3, 3, 16, 32)).astype(np.float32) * 0.1, cuda)
    xa = _t(rs.standard_normal((len(coors), 16)).astype(np.float32), cuda)
    xb = _t(rs.standard_normal((len(coors), 16)).astype(np.float32), cuda)
    ya = sp.conv_table(xa, w, rb2.nbr, rb2.n_out)
    yb = sp.conv_table(xb, w, rb2.nbr, rb2.n_out)
    yab = sp.conv_table(2 * xa - 3 * xb, w, rb2.nbr, rb2.n_out)
    torch.testing.assert_close(yab, 2 * ya - 3 * yb, rtol=1e-4, atol=1e-4)
    assert torch.equal(ya, sp.conv_table(xa, w, rb2.nbr, rb2.n_out))   # fixed summation order


@pytest.mark.parametrize("cin,cout,kshape", [(32, 32, (3, 3, 3)), (32, 64, (3, 3, 3)), (64, 64, (3, 3, 3)),
                                             (64, 128, (3, 3, 3)), (128, 128, (3, 3, 3)),
                                             (128, 128, (3, 1, 1)), (5, 16, (3, 3, 3)),
                                             (16, 16, (3, 3, 3)), (16, 32, (3, 3, 3)), (23, 16, (3, 3, 3))])
def test_kernel_variants_agree_with_fp64(cuda, cin, cout, kshape):
    """Every kernel variant (tcgen05 3xTF32 / lane-group rows / tile FMA) against an fp64 gather-matmul
    on a clustered cloud: 1e-5 of the largest output (the 3xTF32 split must keep fp32 accuracy)."""
    rs = np.random.RandomState(cin * 1000 + cout)
    B, shape, n = 2, [11, 40, 40], 2500
    c = np.unique(np.stack([rs.randint(0, 11, n), rs.randint(5, 30, n), rs.randint(5, 30, n)], 1), axis=0)
    coors = np.concatenate([np.concatenate([np.full((len(c), 1), b), c], 1) for b in range(B)], 0)
    coors = coors[rs.permutation(len(coors))].astype(np.int32)
    subm = kshape == (3, 3, 3)
    rb = sp.build_rulebook(_t(coors, cuda), B, shape, list(kshape), [2, 1, 1], 0, 1, subm)
    feats = rs.standard_normal((len(coors), cin)).astype(np.float32)
    w = (rs.standard_normal(kshape + (cin, cout)) / np.sqrt(cin * 4)).astype(np.float32)
    nbr = rb.nbr.cpu().numpy()
    f64 = np.concatenate([feats.astype(np.float64), np.zeros((1, cin))], 0)
    want = np.zeros((rb.n_out, cout))
    w64 = w.reshape(-1, cin, cout).astype(np.float64)
    for k in range(nbr.shape[0]):
        want += f64[nbr[k]] @ w64[k]      # index -1 hits the appended zero row
    scale = rs.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = rs.standard_normal(cout).astype(np.float32)
    res = rs.standard_normal((rb.n_out, cout)).astype(np.float32)
    want_ep = np.maximum(want * scale + shift + res, 0)
    impls = ["fma"] + (["tc"] if cin % 32 == 0 and cout % 32 == 0 else [])
    for impl in impls + [None]:
        y = sp.conv_table(_t(feats, cuda), _t(w, cuda), rb.nbr, rb.n_out, impl=impl).cpu().numpy()
        # fp32 FMA ~1e-6; 3xTF32 drops the lo*lo term and keeps ~2^-20 per product
        tol = 5e-5 if impl != "fma" and cin % 32 == 0 and cout % 32 == 0 else 1e-5
        assert np.abs(y - want).max() <= tol * np.abs(want).max(), (impl, np.abs(y - want).max())
        y = sp.conv_table(_t(feats, cuda), _t(w, cuda), rb.nbr, rb.n_out, _t(scale, cuda), _t(shift, cuda),
                          _t(res, cuda), True, impl=impl).cpu().numpy()
        assert np.abs(y - want_ep).max() <= tol * np.abs(want_ep).max() + 1e-6, impl


def test_tc_many_tiles_deterministic(cuda):
    """More tiles than SMs (persistent loop, both TMEM buffers, stage ring wrap-around)."""
    rs = np.random.RandomState(9)
    B, shape = 2, [21, 400, 400]
    coors = []
    for b in range(B):
        xy = np.clip(rs.standard_normal((40000, 2)) * 60 + 200, 0, 399).astype(np.int64)
        c = np.unique(np.stack([rs.randint(8, 12, 40000), xy[:, 0], xy[:, 1]], 1), axis=0)
        coors.append(np.concatenate([np.full((len(c), 1), b), c], 1))
    coors = np.concatenate(coors, 0).astype(np.int32)
    rb = sp.build_rulebook(_t(coors, cuda), B, shape, 3, 1, 1, 1, True)
    assert rb.n_out > 148 * 128 * 2
    feats = _t(rs.standard_normal((len(coors), 64)).astype(np.float32), cuda)
    w = _t((rs.standard_normal((3, 3, 3, 64, 64)) / 16).astype(np.float32), cuda)
    a = sp.conv_table(feats, w, rb.nbr, rb.n_out, impl="tc")
    b = sp.conv_table(feats, w, rb.nbr, rb.n_out, impl="tc")
    ref = sp.conv_table(feats, w, rb.nbr, rb.n_out, impl="fma")
    assert torch.equal(a, b)
    assert (a - ref).abs().max() <= 2e-5 * ref.abs().max()


def test_errors(cuda):
    with pytest.raises(RuntimeError):
        sp.build_rulebook(torch.zeros((4, 4), dtype=torch.int32), 1, [4, 4, 4], 3, 1, 1, 1, True)  # CPU tensor
    with pytest.raises(NotImplementedError):
        sp.get_indice_pairs(torch.zeros((4, 4), dtype=torch.int32, device=cuda), 1, [4, 4, 4], transpose=True)
    w = torch.zeros((3, 3, 3, 8, 24), device=cuda)
    rb = sp.build_rulebook(torch.zeros((1, 4), dtype=torch.int32, device=cuda), 1, [4, 4, 4], 3, 1, 1, 1, True)
    with pytest.raises(RuntimeError):
        sp.conv_table(torch.zeros((1, 8), device=cuda), w, rb.nbr, 1)   # c_out 24 unsupported
