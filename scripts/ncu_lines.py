#!/usr/bin/env python3
"""Per-source-line summary of an ncu report (run here, no GPU needed).

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep [kernel-substring] [top-N]

Joins `ncu --page source --csv` (per-SASS-instruction counts and stall samples)
with `nvdisasm -g` line info of the library's cubin, and prints the source
lines that execute the most warp instructions / collect the most stall samples.
"""

import csv
import io
import os
import re
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "tagdigger_b200", "libtagdigger_b200.so")


def line_map(kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], stdout=subprocess.PIPE,
                         universal_newlines=True).stdout
    amap, cur, infn = {}, None, False
    for ln in txt.splitlines():
        if ln.startswith("//-") and ".text." in ln:
            infn = kernel_sub in ln
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            amap[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return amap


def main():
    rep = sys.argv[1]
    ksub = sys.argv[2] if len(sys.argv) > 2 else "count_kernelILb1"
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + ksub], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, universal_newlines=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
    amap = line_map(ksub)
    base = None
    per = {}
    tot_i = tot_s = 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            addr = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        except ValueError:
            continue
        if base is None:
            base = addr
        key, sass = amap.get(addr - base, ((None, 0), "?"))
        n = int(r[ii] or 0)
        s = int(r[isamp] or 0)
        d = per.setdefault(key, {"inst": 0, "samp": 0, "stall": {}})
        d["inst"] += n
        d["samp"] += s
        for k in stalls:
            v = int(r[k] or 0)
            if v:
                d["stall"][hdr[k]] = d["stall"].get(hdr[k], 0) + v
        tot_i += n
        tot_s += s
    src = {}
    print("total warp instructions %d, stall samples %d" % (tot_i, tot_s))
    for key, d in sorted(per.items(), key=lambda kv: -kv[1]["inst"])[:top]:
        f, ln = key if key else ("?", 0)
        if f and f not in src:
            p = os.path.join(REPO, "tagdigger_b200", "csrc", f)
            src[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src.get(f, [])[ln - 1].strip()[:90] if f and 0 < ln <= len(src.get(f, [])) else ""
        st = ",".join("%s:%d" % (k[6:], v) for k, v in sorted(d["stall"].items(), key=lambda kv: -kv[1])[:3])
        print("%5.1f%% inst %5.1f%% samp  %s:%d  %s   [%s]" % (100.0 * d["inst"] / max(tot_i, 1),
              100.0 * d["samp"] / max(tot_s, 1), f, ln, text, st))


if __name__ == "__main__":
    main()
