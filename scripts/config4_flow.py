#!/usr/bin/env python3
"""BASELINE config 4 at its size: 1 B reads (100 bp), 384-plex, 500,000 variable-length tags (20-64 bp;
or the Stacks-style 80-140 bp set with 150 bp reads), sharded over the GPUs of one box -- every rank
counts its share batch by batch (a batch is generated on the device, counted where it lies, dropped),
then ONE all-reduce of the 384 x 500,000 int32 matrix (768 MB) on the counting streams
(tdg_allreduce_matrix; combineReadCounts, /root/reference/tagdigger_fun.py:1088-1095).

    python scripts/config4_flow.py [total_reads] [batch_reads] [C4|C4-stacks]           # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/config4_flow.py ...

Reported: reads/s of the counting kernels (CUDA events of the library, max over ranks), the all-reduce
by itself, and the job = kernels + all-reduce.  Checked: by construction on every batch (the generator
accumulates the expected matrix while it writes the FASTQ: counts >= expected, sum == tag hits of all
ranks) and exactly against the C oracle on a 1 M-read slice of rank 0's first batch.
"""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from oracle import c_oracle
    from tagdigger_b200 import _native, _synth_native, counting, matchset, synth
    total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
    batch = int(float(sys.argv[2])) if len(sys.argv) > 2 else 50_000_000
    shape = sys.argv[3] if len(sys.argv) > 3 else "C4"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bcs, tags, cutsite, site, readlen, mix = synth.shape_tables(shape)
    plan = matchset.plan(bcs, tags, cutsite)
    eng = _native.Engine(local)
    if world > 1:
        counting.init_comm(eng, rank, world)
    matrix = torch.zeros((plan.barnum, plan.ntags), dtype=torch.int32, device="cuda")
    expected = torch.zeros_like(matrix)
    eng.set_tags(plan.tags.patterns, plan.tags.index, any_base=plan.tags.any_base)
    eng.bind_matrix(matrix.data_ptr(), plan.barnum, plan.ntags)
    eng.begin_file(plan.bar.patterns, plan.bar.index, plan.bar_tag_off, any_base=plan.bar.any_base)
    if world > 1:
        eng.allreduce_matrix()             # (the matrix is still all zeros: NCCL sets up its channels on the first call)
        eng.sync()
    gen = _synth_native.Generator(bcs, tags, site, readlen=readlen, seed=4, **mix)
    # one small launch first: the kernel's one-time set-up (module load, shared-memory attributes) is not the job's
    wdev, wbytes = gen.generate(local, 0, 200_000)
    eng.count_device(wdev, wbytes, 0, _native.TDG_PREV_NONE)
    eng.sync()
    gen.free(local, wdev)
    eng.zero_matrix()
    eng.reset_file()
    mine = total // world
    first = rank * mine
    kernel_ms, nbytes_all, done, exact = 0.0, 0, 0, None
    t_wall = time.perf_counter()
    while done < mine:
        n = min(batch, mine - done)
        dev, nbytes = gen.generate(local, first + done, n, expected.data_ptr())
        if rank == 0 and exact is None:
            # a slice of the first batch through the C oracle, cell by cell
            m = min(n, 1_000_000)
            sdev, sbytes = gen.generate(local, first, m)
            img = np.empty(sbytes, dtype=np.uint8)
            eng.memcpy_d2h(img.ctypes.data, sdev, sbytes)
            probe = _native.Engine(local)
            counting.load_plan(probe, plan, nrows=plan.barnum)
            probe.count_device(sdev, sbytes, 0, _native.TDG_PREV_NONE)
            ptot = probe.file_totals()
            got = probe.read_matrix()
            probe.close()
            gen.free(local, sdev)
            want, wtot = c_oracle.count_sharded(img, c_oracle.Counter(bcs, tags, cutsite))
            exact = bool((got == want).all()) and ptot[:3] == wtot
            del img, got, want
        torch.cuda.synchronize()
        eng.timing_begin()
        eng.count_device(dev, nbytes, 0, _native.TDG_PREV_NONE)        # every batch is its own "file": lines numbered from 0
        ms, _ = eng.timing_end()
        kernel_ms += ms
        nbytes_all += nbytes
        done += n
        gen.free(local, dev)
    tot = eng.file_totals()
    wall = time.perf_counter() - t_wall
    # the one exchange
    ar_ms = 0.0
    if world > 1:
        expected_sum = expected.clone()
        dist.all_reduce(expected_sum)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        eng.allreduce_matrix()
        eng.sync()
        ar_ms = (time.perf_counter() - t0) * 1e3
    else:
        expected_sum = expected
    stats = torch.tensor([kernel_ms, ar_ms, float(tot[0]), float(tot[1]), float(tot[2]), float(nbytes_all)], dtype=torch.float64, device="cuda")
    mx = stats.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats)
    ok = bool((matrix >= expected_sum).all().item()) and int(matrix.sum(dtype=torch.int64).item()) == int(stats[4].item()) \
        and int(stats[2].item()) == mine * world
    if rank == 0:
        k_ms, a_ms = float(mx[0].item()), float(mx[1].item())
        reads = mine * world
        p_bar, p_tag = float(stats[3].item()) / reads, float(stats[4].item()) / reads
        alg = float(stats[5].item()) / reads + 32.0 * p_bar + 8.0 * p_tag
        peak = 6545.9
        try:
            peak = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except (OSError, ValueError, KeyError):
            pass
        print(json.dumps({
            "config": "configs[3]: %d reads (%d bp), 384-plex, %d tags of %d-%d bp, %d GPU(s), batches of %d reads per GPU"
                      % (reads, readlen, len(tags), min(map(len, tags)), max(map(len, tags)), world, batch),
            "gpus": world, "reads": reads, "text_bytes": int(stats[5].item()), "kernel_ms_max_over_ranks": round(k_ms, 2),
            "allreduce_768MB_ms": round(a_ms, 3), "reads_per_s_kernels": round(reads / (k_ms * 1e-3), 1),
            "reads_per_s_job": round(reads / ((k_ms + a_ms) * 1e-3), 1),
            "frac_of_hbm_peak": round(alg * (reads / world) / (k_ms * 1e-3) / 1e9 / peak, 4),
            "p_bar": round(p_bar, 3), "p_tag": round(p_tag, 3), "check_by_construction": "ok" if ok else "FAILED",
            "exact_vs_c_oracle_1M_slice": "ok" if exact else "FAILED", "wall_s_with_generation": round(wall, 1)}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0 if ok and (exact is None or exact) else 1


if __name__ == "__main__":
    sys.exit(main())
