"""GPU: the stable radix sort / scan primitives (through the C-ABI) vs numpy."""
import numpy as np
import pytest
import torch

from distill_bev_b200 import _lib

pytestmark = pytest.mark.gpu


def _sort(keys_np, num_bits, dev):
    lib = _lib.load()
    n = keys_np.shape[0]
    keys = torch.from_numpy(keys_np.astype(np.int64)).to(dev).to(torch.int32)  # bit pattern of u32 < 2^31
    ko = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    oo = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    wsb = lib.dbev_sort_workspace_bytes(n)
    ws = _lib.workspace(wsb, dev)
    rc = lib.dbev_sort_keys_iota(_lib.ptr(keys), n, num_bits, _lib.ptr(ko), _lib.ptr(oo),
                                 _lib.ptr(ws), wsb, _lib.stream_ptr(dev))
    _lib.check(rc, "dbev_sort_keys_iota")
    torch.cuda.synchronize()
    return ko[:n].cpu().numpy(), oo[:n].cpu().numpy()


@pytest.mark.parametrize("n,bits", [(0, 8), (1, 1), (31, 5), (4096, 8), (4097, 9), (100003, 17),
                                    (1 << 20, 21), (3000001, 24), (250000, 31)])
def test_radix_sort_is_stable_and_sorted(cuda, n, bits):
    rng = np.random.RandomState(n % 1000 + bits)
    keys = rng.randint(0, 1 << min(bits, 31), size=n, dtype=np.int64) if n else np.zeros(0, np.int64)
    if n > 10:
        keys[rng.randint(0, n, n // 3)] = keys[0]  # heavy duplicates
    ko, oo = _sort(keys, bits, cuda)
    ref_order = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(oo.astype(np.int64), ref_order)
    np.testing.assert_array_equal(ko.astype(np.int64), keys[ref_order])


def test_radix_sort_all_equal_and_presorted(cuda):
    keys = np.full(70000, 5, dtype=np.int64)
    ko, oo = _sort(keys, 3, cuda)
    np.testing.assert_array_equal(oo, np.arange(70000))
    keys = np.arange(50000, dtype=np.int64)[::-1].copy()
    ko, oo = _sort(keys, 16, cuda)
    np.testing.assert_array_equal(ko, np.arange(50000))


@pytest.mark.parametrize("n", [0, 1, 7, 2048, 2049, 100000, 2048 * 2048 + 5])
def test_exclusive_scan(cuda, n):
    lib = _lib.load()
    rng = np.random.RandomState(n % 977)
    a = rng.randint(0, 5, size=n).astype(np.int32)
    t = torch.from_numpy(a).to(cuda)
    out = torch.empty(max(n, 1), dtype=torch.int32, device=cuda)
    tot = torch.full((1,), -1, dtype=torch.int32, device=cuda)
    wsb = lib.dbev_scan_workspace_bytes(n)
    ws = _lib.workspace(wsb, cuda)
    rc = lib.dbev_exclusive_scan_i32(_lib.ptr(t), _lib.ptr(out), n, _lib.ptr(tot), _lib.ptr(ws),
                                     wsb, _lib.stream_ptr(cuda))
    _lib.check(rc, "dbev_exclusive_scan_i32")
    ref = np.concatenate([[0], np.cumsum(a, dtype=np.int64)])
    np.testing.assert_array_equal(out[:n].cpu().numpy(), ref[:-1].astype(np.int32))
    assert int(tot.item()) == int(ref[-1])
