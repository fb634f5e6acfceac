#!/usr/bin/env python3
"""BASELINE config 3 as a flow: K pre-split per-sample FASTQ files (blank Barcode column) counted by
counting.count_files on every GPU of the box -- whole files dealt to the ranks by size, every file
counted straight into its GLOBAL sample row (tagdigger_script.py:123-128 without per-file matrices),
ONE ncclAllReduce of the sample x tag matrix on the counting streams.  Run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        scripts/config3_flow.py [files] [reads_per_file]

Every rank writes its share of the files first (device generator -> host -> /dev/shm or TDG_TMP).
Checked: the reduced matrix against the generator's by-construction counts (>=, and the sum against
the tag hits of all ranks) and, for the first files, every cell against the C oracle.
"""
import contextlib
import io
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
    import bench
    from oracle import c_oracle
    from tagdigger_b200 import _synth_native, counting
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 384
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 250_000
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    _, tags = bench.workload_tables()
    eng = counting.get_engine(local)
    if world > 1:
        counting.init_comm(eng, rank, world)
    gen = _synth_native.Generator([""], tags, bench.CUTSITE, readlen=bench.READLEN, seed=5, p_nobar=0.05, p_unknown=0.30)
    base = os.environ.get("TDG_TMP") or ("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp")
    tmp = os.path.join(base, "tdg_config3")
    os.makedirs(tmp, exist_ok=True)
    names = [os.path.join(tmp, "s%03d.fq" % i) for i in range(k)]
    nsamples = max(1, k // 2)                       # two files per sample: rows are shared between files
    bckeys = {names[i]: [[""], ["Sample%03d" % (i % nsamples)]] for i in range(k)}
    expected = torch.zeros((1, len(tags)), dtype=torch.int32, device="cuda")
    want_rows = np.zeros((nsamples, len(tags)), dtype=np.int64)
    exact = {}
    t0 = time.perf_counter()
    for i in range(rank, k, world):
        expected.zero_()
        dev, nbytes = gen.generate(local, i * m, m, expected.data_ptr())
        host = np.empty(nbytes, dtype=np.uint8)
        eng.memcpy_d2h(host.ctypes.data, dev, nbytes)
        gen.free(local, dev)
        host.tofile(names[i])
        want_rows[i % nsamples] += expected.cpu().numpy()[0]
        if i % nsamples < 2:                        # the files of two samples exactly, through the C oracle
            exact[i] = c_oracle.Counter([""], tags, bench.CUTSITE).count(host)[0][0]
    made = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    runs = []
    for label in ("first", "second"):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        tot = {}
        with contextlib.redirect_stdout(io.StringIO()):
            samples, counts = counting.count_files(bckeys, tags, bench.CUTSITE, device=local, rank=rank, world=world,
                                                   reduce=counting.engine_reduce if world > 1 else None,
                                                   gather=counting.dist_gather if world > 1 else None, totals=tot, as_array=True)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        runs.append(float(dt.item()))
    hits = torch.tensor([sum(v[2] for v in tot.values())], dtype=torch.int64, device="cuda")
    w = torch.from_numpy(want_rows).cuda()
    if world > 1:
        dist.all_reduce(hits)
        dist.all_reduce(w)
    counts = np.asarray(counts, dtype=np.int64)
    ok = bool((counts >= w.cpu().numpy()).all()) and int(counts.sum()) == int(hits.item()) and len(samples) == nsamples
    # exact rows: a sample whose two files were both counted exactly by the oracle
    ex_ok, ex_n = True, 0
    got = [None] * k
    if world > 1:
        allex = [None] * world
        dist.all_gather_object(allex, exact)
        exact = {i: v for d in allex for i, v in d.items()}
    for s_idx in range(nsamples):
        files_of = [i for i in range(k) if i % nsamples == s_idx]
        if all(i in exact for i in files_of):
            ex_n += 1
            ex_ok = ex_ok and bool((counts[s_idx] == sum(exact[i] for i in files_of)).all())
    if rank == 0:
        total_bytes = sum(os.path.getsize(f) for f in names)
        best = min(runs)
        print(json.dumps({"config": "configs[2] as a flow: %d pre-split files x %d reads (blank Barcode), 40,000 tags, "
                                    "counting.count_files(world=%d)" % (k, m, world),
                          "gpus": world, "files": k, "reads": k * m, "bytes": total_bytes, "samples": nsamples,
                          "seconds": [round(x, 3) for x in runs], "reads_per_s": round(k * m / best, 1),
                          "text_GBps": round(total_bytes / best / 1e9, 2), "files_made_in_s": round(made, 1),
                          "check_by_construction": "ok" if ok else "FAILED",
                          "exact_vs_c_oracle": ("ok (%d sample rows)" % ex_n) if ex_ok and ex_n else ("FAILED" if not ex_ok else "none"),
                          "where": tmp}))
    if world > 1:
        dist.barrier()
    for i in range(rank, k, world):
        os.remove(names[i])
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0 if ok and ex_ok else 1


if __name__ == "__main__":
    sys.exit(main())
