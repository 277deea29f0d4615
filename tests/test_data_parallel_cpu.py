"""CPU, world_size 2 over gloo: the gradient all-reduce of the data-parallel training step (SURVEY.md §8 rows C1 / (e);
reference: MMDistributedDataParallel, tools/distributed.py:11-79). Two ranks, each with half of a batch, must end
up with the gradients of ONE process that saw the whole batch (losses are sums / B per rank, DDP averages over
ranks) - flat buckets, bf16 wire compression and the overlapped (hook-driven) mode included."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(8, 16, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(16, 4, 1),
                               torch.nn.Flatten(), torch.nn.Linear(4 * 6 * 6, 3))


def _data():
    g = torch.Generator().manual_seed(1)
    return torch.randn(4, 8, 6, 6, generator=g), torch.randn(4, 3, generator=g)


def _loss(model, x, y):
    return ((model(x) - y) ** 2).sum() / x.shape[0]      # sum / per-rank batch, like bevdet_distill.py:1259-1262


def _worker(rank, world, port, mode, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from distill_bev_b200.plugin.data_parallel import GradientAllReduce
    model = _model()
    x, y = _data()
    xs, ys = x[rank * 2:rank * 2 + 2], y[rank * 2:rank * 2 + 2]
    red = GradientAllReduce(model.parameters(), world, bucket_bytes=2048,
                            comm_dtype=torch.bfloat16 if mode == "bf16" else None, overlap=(mode == "overlap"))
    for _ in range(2):                      # two steps: begin() must reset the flat buffers
        red.begin()
        _loss(model, xs, ys).backward()
        red.finish()
    grads = [p.grad.detach().clone() for p in model.parameters()]
    info = red.describe()
    gathered = [None] * world
    dist.all_gather_object(gathered, [g.tolist() for g in grads])
    if rank == 0:
        out.put((gathered, info))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["f32", "bf16", "overlap"])
def test_two_rank_gradients_equal_single_process_full_batch(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + {"f32": 0, "bf16": 1, "overlap": 2}[mode]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, info = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    model = _model()
    x, y = _data()
    _loss(model, x, y).backward()
    want = [p.grad for p in model.parameters()]
    tol = dict(rtol=2e-2, atol=2e-2) if mode == "bf16" else dict(rtol=1e-5, atol=1e-6)
    for r in range(2):
        for got, w in zip(gathered[r], want):
            torch.testing.assert_close(torch.tensor(got).reshape(w.shape), w, **tol)
    # both ranks hold identical gradients after the all-reduce
    for a, b in zip(gathered[0], gathered[1]):
        assert a == b
    assert len(info["buckets"]) > 1 and info["parameters"] == 6
    assert info["overlap_with_backward"] == (mode == "overlap")
    assert info["comm_dtype"] == ("bfloat16" if mode == "bf16" else "float32")


def test_single_process_is_a_no_op():
    sys.path.insert(0, ROOT)
    from distill_bev_b200.plugin.data_parallel import GradientAllReduce
    model = _model()
    x, y = _data()
    red = GradientAllReduce(model.parameters(), world_size=1)
    red.begin()
    _loss(model, x, y).backward()
    red.finish()
    ref = _model()
    _loss(ref, x, y).backward()
    for p, q in zip(model.parameters(), ref.parameters()):
        torch.testing.assert_close(p.grad, q.grad)
