#!/usr/bin/env python3
"""Time of the path's one collective alone, CUDA events, under torchrun: tdg_allreduce_matrix (one
ncclAllReduce(int32, sum) on the counting stream) for the matrices of config 2 (96 x 40,000 =
15.4 MB) and config 4 (384 x 500,000 = 768 MB: combineReadCounts' sum, tagdigger_fun.py:1088-1095,
at its largest named shape)."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import torch
    import torch.distributed as dist
    from tagdigger_b200 import _native, counting
    rank = int(os.environ["RANK"])
    local = int(os.environ["LOCAL_RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = _native.Engine(local)
    counting.init_comm(eng, rank, world)
    out = []
    for rows, cols in ((96, 40000), (384, 500000)):
        m = torch.ones((rows, cols), dtype=torch.int32, device="cuda")
        eng.bind_matrix(m.data_ptr(), rows, cols)
        n = 20 if rows * cols > 10 ** 7 else 100
        for _ in range(5):
            eng.allreduce_matrix()
        eng.sync()
        m.fill_(1)
        torch.cuda.synchronize()
        dist.barrier()
        ts = torch.cuda.ExternalStream(eng.stream())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        for _ in range(n):
            eng.allreduce_matrix()
        e1.record(ts)
        eng.sync()
        t = torch.tensor([e0.elapsed_time(e1) / n], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = bool((m == world ** n).all().item()) if world ** n < 2 ** 31 else None
        out.append({"matrix": "%d x %d int32" % (rows, cols), "MB": round(rows * cols * 4 / 1e6, 1),
                    "ms_per_allreduce": round(float(t.item()), 4),
                    "bus_GBps": round(rows * cols * 4 * 2 * (world - 1) / world / (float(t.item()) * 1e-3) / 1e9, 1),
                    "sum_check": ok})
    if rank == 0:
        print(json.dumps({"world": world, "collective": "tdg_allreduce_matrix (ncclAllReduce int32 sum, in place, counting stream)",
                          "results": out}))
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
