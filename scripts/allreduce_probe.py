#!/usr/bin/env python3
"""Time of the bench's all-reduce alone (int32 96 x 40,000 matrix), CUDA events, under torchrun."""
import os
import torch
import torch.distributed as dist


def main():
    rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    m = torch.zeros((96, 40000), dtype=torch.int32, device="cuda")
    for _ in range(10):
        dist.all_reduce(m)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 100
    e0.record()
    for _ in range(n):
        dist.all_reduce(m)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / n
    # reduce to rank 0 only
    e0.record()
    for _ in range(n):
        dist.reduce(m, dst=0)
    e1.record(); torch.cuda.synchronize()
    t2 = e0.elapsed_time(e1) / n
    if rank == 0:
        print("world %d: all_reduce %.1f us, reduce %.1f us per call (15.4 MB int32)" % (dist.get_world_size(), t * 1e3, t2 * 1e3))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
