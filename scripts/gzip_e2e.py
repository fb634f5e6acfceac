#!/usr/bin/env python3
"""End-to-end rate of tdg_count_file on ORDINARY gzip files (one deflate stream, no index):
zlib on one thread against the speculative parallel inflater (csrc/tdg_pgz.h) at several
thread counts.  Synthetic config-2 FASTQ written to local disk.

    python scripts/gzip_e2e.py [reads]
"""
import gzip
import json
import os
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402


def main():
    import bench
    from tagdigger_b200 import _native, _synth_native, counting, matchset
    reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    bcs, tags = bench.workload_tables()
    plan = matchset.plan(bcs, tags, bench.CUTSITE)
    eng = _native.Engine(0)
    gen = _synth_native.Generator(bcs, tags, bench.CUTSITE, readlen=bench.READLEN, seed=bench.SEED)
    dev, nbytes = gen.generate(0, 0, reads)
    host = np.empty(nbytes, dtype=np.uint8)
    eng.memcpy_d2h(host.ctypes.data, dev, nbytes)
    gen.free(0, dev)
    raw = host.tobytes()
    del host
    tmp = tempfile.mkdtemp(dir=os.environ.get("TDG_TMP", "/tmp"))
    counting.load_plan(eng, plan, nrows=plan.barnum)
    out = {"reads": reads, "uncompressed_MB": round(nbytes / 1e6, 1), "host_cpus": os.cpu_count(), "files": {}}
    want = None
    for level in (1, 6):
        path = os.path.join(tmp, "reads_l%d.fastq.gz" % level)
        t0 = time.perf_counter()
        with open(path, "wb") as fh:
            fh.write(gzip.compress(raw, level))
        rec = {"file_MB": round(os.path.getsize(path) / 1e6, 1), "python_compress_s": round(time.perf_counter() - t0, 1), "runs": {}}
        for threads in (1, 4, 8, 16, 32):
            if threads > (os.cpu_count() or 1):
                continue
            os.environ["TDG_IO_THREADS"] = str(threads)
            best = None
            for rep in range(2):
                eng.zero_matrix()
                eng.reset_file()
                t0 = time.perf_counter()
                tot = eng.count_file(path, True)
                m = eng.read_matrix()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            if want is None:
                want = m.copy()                       # zlib path, one thread
            assert tot[0] == reads and np.array_equal(m, want), "parallel inflate changed the counts"
            rec["runs"]["zlib, 1 thread" if threads == 1 else "parallel, %d threads" % threads] = {
                "seconds": round(best, 3), "reads_per_s": round(reads / best, 1), "uncompressed_MB_per_s": round(nbytes / best / 1e6, 1)}
        out["files"]["gzip -%d single member" % level] = rec
        os.remove(path)
    os.rmdir(tmp)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
