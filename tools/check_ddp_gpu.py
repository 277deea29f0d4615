"""2-GPU check of the data-parallel training step (torchrun --nproc-per-node 2 tools/check_ddp_gpu.py): after one step with
the overlapped, graph-capturable GradientAllReduce every rank holds the SAME gradients, equal to the average of the ranks'
local gradients (bf16 wire compression: 1e-2 of each tensor's max entry; fp32 wire: 1e-6)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    res = {}

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))

    for comm in ("f32", "bf16"):
        hp = bench.HotPath(dev, seed=bench.rank_seed(rank), world=world, allreduce="after", comm_dtype=comm)
        packed = hp.dbev.fgd.PackedBoxes(hp.d_boxes, hp.d_box_offs, hp.max_boxes)
        args = (hp.d_calib, hp.d_points, hp.d_labels, packed)
        # (1) the reducer itself, exactly: 'after' mode leaves the local gradients in p.grad until finish()
        hp._forward_backward(*args)
        torch.cuda.synchronize()
        local_grads = [p.grad.detach().clone() for p in hp.trainable]
        hp.reducer.finish()
        torch.cuda.synchronize()
        worst_avg, worst_same = 0.0, 0.0
        for p, lg in zip(hp.trainable, local_grads):
            gathered = [torch.empty_like(lg) for _ in range(world)]
            dist.all_gather(gathered, lg)
            want = sum(gathered) / world
            worst_avg = max(worst_avg, rel(p.grad, want))
            mine = [torch.empty_like(lg) for _ in range(world)]
            dist.all_gather(mine, p.grad.contiguous())
            worst_same = max(worst_same, float((mine[0] - mine[-1]).abs().max()))
        after = [p.grad.detach().clone() for p in hp.trainable]
        # (2) run-to-run noise of the step itself (sort-free splat: float reductions in arbitrary order, amplified by the
        # TF32 encoder's ReLU masks): a second 'after' step against the first
        hp._forward_backward(*args)
        hp.reducer.finish()
        torch.cuda.synchronize()
        noise = max(rel(p.grad, a) for p, a in zip(hp.trainable, after))
        # (3) the overlapped mode (hooks, side streams) against the 'after' result
        hp.set_allreduce("overlap")
        hp._forward_backward(*args)
        hp.reducer.finish()
        torch.cuda.synchronize()
        overlap = max(rel(p.grad, a) for p, a in zip(hp.trainable, after))
        res[comm] = {"reducer_max_rel_err_vs_average_of_local_grads": worst_avg, "max_abs_diff_between_ranks": worst_same,
                     "run_to_run_noise_of_the_step": noise, "overlap_vs_after": overlap, "collective": hp.reducer.describe()}
        assert worst_same == 0.0, worst_same
        assert worst_avg <= (1e-2 if comm == "bf16" else 1e-6), worst_avg
        assert overlap <= max(3.0 * noise, 2e-2 if comm == "bf16" else 1e-5), (overlap, noise)
        del hp
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(res))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ddp_check.json"), "w"), indent=1)
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)


if __name__ == "__main__":
    main()
