#!/usr/bin/env python3
"""Rate of the scan-only kernel (count_kernel<false>: TMA ring + line-end scan, no matching) over a
resident FASTQ image: how fast the warps' bulk copies pull the stream out of HBM when little work is
done per tile -- the ceiling the counting kernel's feed can reach with this access pattern.

    python scripts/lines_rate.py [reads]
"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    reads = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000_000
    from tagdigger_b200 import _native, _synth_native
    bcs, tags = bench.workload_tables()
    eng = _native.Engine(0)
    gen = _synth_native.Generator(bcs, tags, bench.CUTSITE, readlen=bench.READLEN, seed=bench.SEED)
    dev, nbytes = gen.generate(0, 0, reads)
    for _ in range(3):
        lines, _ = eng.count_lines_device(dev, nbytes, 0, _native.TDG_PREV_NONE)
    eng.timing_begin()
    steps = 5
    for _ in range(steps):
        lines, _ = eng.count_lines_device(dev, nbytes, 0, _native.TDG_PREV_NONE)
    ms, n = eng.timing_end()
    k = ms / n
    print(json.dumps({"kernel": "count_kernel<false> + verify_kernel", "reads": reads, "bytes": nbytes, "lines": lines,
                      "ms": round(k, 4), "stream_GBps": round(nbytes / (k * 1e-3) / 1e9, 1),
                      "lib": os.environ.get("TDG_LIB", "default")}))
    gen.free(0, dev)
    eng.close()


if __name__ == "__main__":
    main()
