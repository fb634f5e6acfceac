#!/usr/bin/env python3
"""Instructions and stall samples of count_kernel<true> by code region (source-line markers).

    python scripts/ncu_regions.py report.ncu-rep lib.so reads tile_bytes

Regions are found by searching csrc/tdg_kernel.cuh for marker text, so the table follows edits.
"""
import collections, csv, io, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines

MARKS = [("tma/mbar helpers", "__device__ __forceinline__ uint32_t smem_u32"),
         ("scan helpers", "// Line-end candidates, 16 bytes at a time"),
         ("fetch/general", "// Unaligned 32-character window"),
         ("pack_word", "// One 4-character word of the fast matcher"),
         ("prologue", "template <bool MATCH>\n__global__"),
         ("produce", "// ---- producer (lane 0)"),
         ("setup", "// ---- matcher set-up"),
         ("front", "auto batch_front = [&]"),
         ("back", "auto batch_back = [&]"),
         ("loop top", "    uint32_t s = 0, parity = 0;"),
         ("classify", "auto classify = [&]()"),
         ("prefix", "            uint32_t incl, total;"),
         ("guess", "                if (need_guess) {"),
         ("emit-arith", "// ---- emission: queue the starts of sequence lines"),
         ("walk", "                    if (simple) {"),
         ("drain", "                    if (!verified) {"),
         ("seg end/refill", "            seg_lines += total;"),
         ("open", "// ---- open the next tile"),
         ("scan", "// ---- scan: control-character mask"),
         ("last tile", "        if (last_tile) {"),
         ("maskstore", "// the masks go to shared memory right away"),
         ("epilogue", "    if (MATCH) {\n        // totals: one set of atomics per warp")]


def main():
    rep, lib, reads, tile = sys.argv[1], sys.argv[2], float(sys.argv[3]), float(sys.argv[4])
    ncu_lines.LIB = os.path.abspath(lib)
    src = open(os.path.join(ncu_lines.REPO, "tagdigger_b200", "csrc", "tdg_kernel.cuh")).read()
    starts = []
    for name, text in MARKS:
        i = src.find(text)
        if i < 0:
            print("marker not found:", name)
            continue
        starts.append((src.count("\n", 0, i) + 1, name))
    starts.sort()
    tiles = reads * 249.82 / tile

    def region(f, l):
        if f != "tdg_kernel.cuh":
            return "intrinsics (" + str(f) + ")"
        name = "?"
        for ln, nm in starts:
            if ln <= l:
                name = nm
        return name
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, universal_newlines=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hi]
    ia, ii, ist = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
    amap = ncu_lines.line_map("count_kernelILb1")
    ci, cs, base = collections.Counter(), collections.Counter(), None
    for r in rows[hi + 1:]:
        addr = int(r[ia], 16)
        if base is None:
            base = addr
        key, _ = amap.get(addr - base, ((None, 0), "?"))
        f, l = key if key else ("?", 0)
        g = region(f, l)
        ci[g] += int(r[ii] or 0)
        cs[g] += int(r[ist] or 0)
    ti, ts = sum(ci.values()), sum(cs.values())
    print("tiles %.0f; warp instructions per tile %.1f, per read %.2f" % (tiles, ti / tiles, ti / reads))
    print("%-34s %10s %7s %9s %6s" % ("region", "instr/tile", "instr%", "samples%", "rel"))
    for g, n in ci.most_common():
        print("%-34s %10.1f %7.1f %9.1f %6.2f" % (g, n / tiles, 100 * n / ti, 100 * cs[g] / ts, (cs[g] / ts) / (n / ti) if n else 0))


if __name__ == "__main__":
    main()
