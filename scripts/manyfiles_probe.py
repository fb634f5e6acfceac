#!/usr/bin/env python3
"""Where a key of many small files spends its time: N plain files of M reads through counting.count_files,
per-file phases printed by the library (TDG_FILE_DEBUG=1) and the Python side's share.

    TDG_FILE_DEBUG=1 python scripts/manyfiles_probe.py [files] [reads_per_file]
"""
import contextlib
import io
import os
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402


def main():
    from tagdigger_b200 import counting, synth
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 80000
    rng = np.random.default_rng(3)
    _, _, seqs = synth.make_marker_pairs(2000, rng)
    tags = [s for p in seqs for s in p]
    fq, _ = synth.make_fastq(m, [""], tags, rng)
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = tempfile.mkdtemp(prefix="tdg_many_", dir=base)
    names = [os.path.join(tmp, "s%03d.fq" % i) for i in range(k)]
    for n in names:
        with open(n, "wb") as fh:
            fh.write(fq)
    bckeys = {n: [[""], ["S%03d" % i]] for i, n in enumerate(names)}
    for rep in range(3):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            counting.count_files(bckeys, tags, "TGCAG", as_array=True)
        dt = time.perf_counter() - t0
        print("count_files: %d files x %d reads (%.1f MB each): %.3f s = %.2f ms per file, %.1f M reads/s"
              % (k, m, len(fq) / 1e6, dt, dt / k * 1e3, k * m / dt / 1e6), file=sys.stderr)
    for n in names:
        os.remove(n)
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
