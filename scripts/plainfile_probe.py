#!/usr/bin/env python3
"""One plain FASTQ file of the config-2 shape through tdg_count_file, five times, with the library's per-file phase
line (TDG_FILE_DEBUG=1): is the reader or the copy the wall on this box?

    TDG_FILE_DEBUG=1 python scripts/plainfile_probe.py [reads]
"""
import os
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402


def main():
    import bench
    from tagdigger_b200 import _synth_native, counting, matchset
    reads = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
    bcs, tags = bench.workload_tables()
    plan = matchset.plan(bcs, tags, bench.CUTSITE)
    eng = counting.get_engine(0)
    counting.load_plan(eng, plan, nrows=plan.barnum)
    gen = _synth_native.Generator(bcs, tags, bench.CUTSITE, readlen=bench.READLEN, seed=bench.SEED)
    dev, nbytes = gen.generate(0, 0, reads)
    img = np.empty(nbytes, dtype=np.uint8)
    eng.memcpy_d2h(img.ctypes.data, dev, nbytes)
    gen.free(0, dev)
    tmp = tempfile.mkdtemp(prefix="tdg_plain_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    path = os.path.join(tmp, "c2.fq")
    img.tofile(path)
    for rep in range(5):
        eng.zero_matrix()
        eng.reset_file()
        t0 = time.perf_counter()
        eng.count_file(path, False)
        dt = time.perf_counter() - t0
        print("plain file: %.3f s = %.1f M reads/s, %.1f GB/s" % (dt, reads / dt / 1e6, nbytes / dt / 1e9), file=sys.stderr)
    os.remove(path)
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
