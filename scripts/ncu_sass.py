#!/usr/bin/env python3
"""SASS of a source-line range of count_kernel<true> with per-tile execution counts and stall samples.

    python scripts/ncu_sass.py report.ncu-rep lib.so reads tile_bytes first_line last_line [min_per_tile]
"""
import csv, io, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines


def main():
    rep, lib, reads, tile, l0, l1 = sys.argv[1], sys.argv[2], float(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
    minf = float(sys.argv[7]) if len(sys.argv) > 7 else 0.3
    ncu_lines.LIB = os.path.abspath(lib)
    tiles = reads * 249.82 / tile
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, universal_newlines=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]
    ia, ii, ist, isrc = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
    amap = ncu_lines.line_map("count_kernelILb1")
    base = None
    tot = 0.0
    for r in rows[hi + 1:]:
        addr = int(r[ia], 16)
        if base is None:
            base = addr
        key, _ = amap.get(addr - base, ((None, 0), "?"))
        f, l = key if key else ("?", 0)
        n = int(r[ii] or 0) / tiles
        if f == "tdg_kernel.cuh" and l0 <= l <= l1 and n >= minf:
            tot += n
            print("%04x %4d %6.2f %5s  %s" % (addr - base, l, n, r[ist], r[isrc].strip()))
    print("total %.1f instr/tile" % tot)


if __name__ == "__main__":
    main()
