#!/usr/bin/env python3
"""Where the host feed spends its time on an ordinary gzip file (CPU only): phase A (speculative
parallel inflate), phase B (serial chaining) and the drain (marker resolution + CRC) of tdg_pgz.h,
summed over the rounds, for a few thread counts.  TDG_PGZ_DEBUG=1 makes the reader print the
per-round timings this script adds up.

    python scripts/pgz_profile.py [reads]
"""
import os
import re
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

CHILD = r"""
import os, sys, time
sys.path.insert(0, %r); sys.path.insert(0, %r)
from feed_check import read_file
path, size = sys.argv[1], int(sys.argv[2])
t = time.time()
got, mode = read_file(path, True, 64 << 20, cap=size + 4096)
dt = time.time() - t
print("RESULT", mode, len(got), dt)
""" % (REPO, os.path.join(REPO, "tests"))


def main():
    import numpy as np
    import bench
    from tagdigger_b200 import synth
    reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = tempfile.mkdtemp(dir=base)
    bcs, tags = bench.workload_tables()
    rng = np.random.default_rng(1)
    plain = os.path.join(tmp, "a.fq")
    with open(plain, "wb") as fh:
        for _ in range(max(1, reads // 500000)):
            fh.write(synth.make_fastq(500000, bcs, tags, rng)[0])
    gz = plain + ".gz"
    bench.write_gzip_parallel(plain, gz, level=6)
    size = os.path.getsize(plain)
    print("text %d bytes, gzip %d bytes, host cpus %d" % (size, os.path.getsize(gz), os.cpu_count()))
    for thr in (1, 4, 8, 16, 32):
        if thr > (os.cpu_count() or 1) and thr != 1:
            continue
        env = dict(os.environ, TDG_IO_THREADS=str(thr), TDG_PGZ_DEBUG="1")
        p = subprocess.run([sys.executable, "-c", CHILD, gz, str(size)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           universal_newlines=True)
        a = sum(float(x) for x in re.findall(r"pgz round: A ([\d.]+) ms", p.stderr))
        b = sum(float(x) for x in re.findall(r"ms, B ([\d.]+) ms", p.stderr))
        d = sum(float(x) for x in re.findall(r"tasks, ([\d.]+) ms", p.stderr))
        m = re.search(r"RESULT (\S+) (\d+) ([\d.]+)", p.stdout)
        if not m:
            print(thr, "FAILED", p.stderr[-300:])
            continue
        dt = float(m.group(3))
        print("threads %2d  mode %-6s  %.2f GB/s of text   total %.3f s: phase A %.3f, chain %.3f, drain %.3f, other %.3f"
              % (thr, m.group(1), int(m.group(2)) / dt / 1e9, dt, a / 1e3, b / 1e3, d / 1e3, dt - (a + b + d) / 1e3))
    for f in (plain, gz):
        os.remove(f)
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
