#!/usr/bin/env python3
"""Group the SASS instructions of an ncu report by how often they execute per tile and by
source line (read here, no GPU).  Shows what the once-per-tile bookkeeping, the candidate
loop and the batch code cost in warp instructions.

    python scripts/ncu_freq.py gpurun_out/prof.ncu-rep reads_per_launch [bytes_per_read] [tile_bytes]
"""
import collections
import csv
import io
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines  # noqa: E402


def main():
    rep = sys.argv[1]
    reads = float(sys.argv[2])
    bpr = float(sys.argv[3]) if len(sys.argv) > 3 else 249.82
    tile = float(sys.argv[4]) if len(sys.argv) > 4 else 5632.0
    tiles = reads * bpr / tile
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, universal_newlines=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ia, ii = hdr.index("Address"), hdr.index("Instructions Executed")
    amap = ncu_lines.line_map("count_kernelILb1")
    src = open(os.path.join(ncu_lines.REPO, "tagdigger_b200", "csrc", "tdg_kernel.cuh")).read().splitlines()
    base = None
    groups = collections.defaultdict(lambda: collections.Counter())
    tot = 0
    for r in rows[hi + 1:]:
        addr = int(r[ia], 16)
        if base is None:
            base = addr
        n = int(r[ii] or 0)
        tot += n
        f = n / tiles
        bucket = "once per tile" if 0.97 <= f <= 1.03 else "%.2f per tile" % (round(f * 4) / 4) if f >= 0.2 else "rare"
        key, _ = amap.get(addr - base, ((None, 0), "?"))
        groups[bucket][key] += f
    print("tiles %.0f, warp instructions per tile %.1f, per read %.2f" % (tiles, tot / tiles, tot / reads))
    for bucket, per in sorted(groups.items(), key=lambda kv: -sum(kv[1].values())):
        total = sum(per.values())
        print("== %s: %.1f instr/tile" % (bucket, total))
        for key, v in per.most_common(int(os.environ.get("TOPN","14"))):
            f, ln = key if key else ("?", 0)
            text = src[ln - 1].strip()[:80] if f == "tdg_kernel.cuh" and 0 < ln <= len(src) else ""
            print("   %6.1f  %s:%d  %s" % (v, f, ln, text))


if __name__ == "__main__":
    main()
