#!/usr/bin/env python3
"""The gzip leg of the path users run, device feed against host feed: a config-2 FASTQ (bench.py's
tables) as ONE ordinary gzip member, (a) inflated by tdg_gz_inflate_host with the time per stage,
(b) counted by tdg_count_file with the device feed (default) and with TDG_GZDEV=0 (host threads).

    python scripts/gzdev_bench.py [reads] [level] [bgzf]      # bgzf: the file as BGZF members (bgzip's format) instead
"""
import json
import os
import shutil
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402


def _bgzf_piece(job):
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from feed_check import bgzf_compress
    path, a, b, last, level = job
    with open(path, "rb") as fh:
        fh.seek(a)
        data = fh.read(b - a)
    return bgzf_compress(data, eof_marker=last, level=level)


def write_bgzf_parallel(src, dst, level=6, piece=65280 * 256):
    import multiprocessing as mp
    size = os.path.getsize(src)
    cuts = list(range(0, size, piece)) + [size]
    jobs = [(src, a, b, b == size, level) for a, b in zip(cuts[:-1], cuts[1:])]
    with mp.get_context("fork").Pool(min(len(jobs), os.cpu_count() or 1)) as pool, open(dst, "wb") as out:
        for blob in pool.imap(_bgzf_piece, jobs):
            out.write(blob)


def main():
    import bench
    from oracle import c_oracle
    from tagdigger_b200 import _synth_native, counting, matchset
    reads = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
    level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    bcs, tags = bench.workload_tables()
    plan = matchset.plan(bcs, tags, bench.CUTSITE)
    eng = counting.get_engine(0)
    counting.load_plan(eng, plan, nrows=plan.barnum)
    gen = _synth_native.Generator(bcs, tags, bench.CUTSITE, readlen=bench.READLEN, seed=bench.SEED)
    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > (16 << 30) else None
    tmp = tempfile.mkdtemp(prefix="tdg_gzdev_", dir=base)
    res = {"reads": reads, "gzip_level": level, "host_cpus": os.cpu_count()}
    try:
        dev, nbytes = gen.generate(0, 0, reads)
        img = np.empty(nbytes, dtype=np.uint8)
        eng.memcpy_d2h(img.ctypes.data, dev, nbytes)
        gen.free(0, dev)
        plain = os.path.join(tmp, "c2.fq")
        img.tofile(plain)
        want, wtot = c_oracle.count_sharded(img, c_oracle.Counter(bcs, tags, bench.CUTSITE))
        gz = plain + ".gz"
        if len(sys.argv) > 3 and sys.argv[3] == "bgzf":
            write_bgzf_parallel(plain, gz, level=level)
            res["format"] = "BGZF"
        else:
            bench.write_gzip_parallel(plain, gz, level=level)
        res["text_bytes"] = int(nbytes)
        res["gzip_bytes"] = os.path.getsize(gz)
        runs = []
        for rep in range(3):
            t0 = time.perf_counter()
            out, info, ms = eng.gz_inflate_host(gz, nbytes + 4096)
            dt = time.perf_counter() - t0
            runs.append({"seconds": round(dt, 3), "info": info, "ms": ms, "same_bytes": bool(out == img.tobytes()) if rep == 0 else None})
            del out
        res["inflate_to_host"] = runs
        for label, env in (("device_feed", None), ("host_feed", "0")):
            if env is None:
                os.environ.pop("TDG_GZDEV", None)
            else:
                os.environ["TDG_GZDEV"] = env
            best = None
            for rep in range(3 if env is None else 2):
                eng.zero_matrix()
                eng.reset_file()
                t0 = time.perf_counter()
                tot = eng.count_file(gz, True)
                got = eng.read_matrix()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            res[label] = {"seconds": round(best, 3), "reads_per_s": round(reads / best, 1), "text_GBps": round(nbytes / best / 1e9, 2),
                          "exact_vs_c_oracle": "ok" if bool((got == want).all()) and tot[:3] == wtot else "FAILED"}
        os.environ.pop("TDG_GZDEV", None)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
