#!/usr/bin/env python3
"""BASELINE config 3 in small: K pre-split per-sample FASTQ files (blank Barcode column) counted
through counting.count_files (one tag table, global sample rows, no per-file matrices).

    python scripts/multifile_e2e.py [files] [reads_per_file]
"""
import contextlib
import io
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
    from tagdigger_b200 import _native, _synth_native, counting
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    _, tags = bench.workload_tables()
    eng = counting.get_engine(0)
    gen = _synth_native.Generator([""], tags, bench.CUTSITE, readlen=bench.READLEN, seed=5, p_nobar=0.05, p_unknown=0.30)
    tmp = tempfile.mkdtemp(dir=os.environ.get("TDG_TMP", "/tmp"))
    bckeys = {}
    total_bytes = 0
    for i in range(k):
        dev, nbytes = gen.generate(0, i * m, m)
        host = np.empty(nbytes, dtype=np.uint8)
        eng.memcpy_d2h(host.ctypes.data, dev, nbytes)
        gen.free(0, dev)
        name = os.path.join(tmp, "s%03d.fq" % i)
        host.tofile(name)
        bckeys[name] = [[""], ["Sample%03d" % (i % (k // 2 or 1))]]
        total_bytes += nbytes
    res = {}
    for label in ("first", "second"):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            samples, counts = counting.count_files(bckeys, tags, bench.CUTSITE)
        dt = time.perf_counter() - t0
        res[label] = {"seconds": round(dt, 3), "reads_per_s": round(k * m / dt, 1), "GB_per_s": round(total_bytes / dt / 1e9, 2)}
    res["files"] = k
    res["reads_per_file"] = m
    res["samples"] = len(samples)
    res["tag_hits"] = int(np.asarray(counts).sum())
    print(json.dumps(res, indent=1))
    for f in bckeys:
        os.remove(f)
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
