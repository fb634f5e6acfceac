#!/usr/bin/env python3
"""Extended differential fuzz of the counting kernel's FAST path against the C oracle (needs a GPU).

Well-formed synthetic FASTQ of random shapes (read length, barcode count, uniform / variable tag
lengths, LF / CRLF) is damaged at random places -- N, lower case, tabs, NUL, a non-ASCII pair, a
lone CR, a blank line, a byte of a header turned into a line feed -- and counted through the C ABI
both as one device chunk and through the streaming entry point; counts, totals and line counts
must equal the oracle's.

    python scripts/gpu_fuzz.py [iterations] [first_seed]
"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from oracle import c_oracle                                   # noqa: E402  (test infrastructure)
from tagdigger_b200 import _native, counting, matchset, synth  # noqa: E402


def damage(buf, rng, kinds, rate):
    a = np.frombuffer(buf, dtype=np.uint8).copy()
    n = int(len(a) * rate)
    if n == 0:
        return a.tobytes()
    pos = rng.integers(0, len(a) - 1, size=n)
    pool = {"N": 0x4E, "tab": 0x09, "nul": 0x00, "cr": 0x0D, "lf": 0x0A, "space": 0x20, "vt": 0x0B, "us": 0x1F}
    pairs = []
    for p in pos:
        k = kinds[int(rng.integers(0, len(kinds)))]
        if k == "lower":
            if 65 <= a[p] <= 90:
                a[p] |= 0x20
        elif k == "hi":
            pairs.append(int(p))
        else:
            a[p] = pool[k]
    # a two-byte UTF-8 character (e-acute), written last and only over two ASCII bytes, so that the
    # image stays valid UTF-8 (what the reference reads as text)
    for p in pairs:
        if a[p] < 0x80 and a[p + 1] < 0x80 and (p == 0 or a[p - 1] < 0x80):
            a[p], a[p + 1] = 0xC3, 0xA9
    return a.tobytes()


def run(iters, seed0=0):
    """Number of mismatching iterations."""
    eng = _native.Engine(0)
    t0 = time.time()
    bad = 0
    for it in range(iters):
        rng = np.random.default_rng(7000 + seed0 + it)
        cutsite = "TGCAG"
        nbar = int(rng.choice([4, 24, 96, 200]))
        bcs = synth.make_barcodes(nbar, rng, cutsite=cutsite)
        if rng.random() < 0.5:
            L = None
            tl = int(rng.choice([20, 33, 48, 63, 64]))
            _, _, seqs = synth.make_marker_pairs(150, rng, length=tl, cutsite=cutsite)
        else:
            L = rng.integers(20, 65, size=150)
            _, _, seqs = synth.make_marker_pairs(150, rng, cutsite=cutsite, lengths=L)
        tags = [s for p in seqs for s in p]
        readlen = int(rng.choice([40, 76, 100, 101, 125, 150]))
        newline = [b"\n", b"\r\n"][int(rng.random() < 0.25)]
        nreads = int(rng.choice([3000, 20000, 60000]))
        fq, _ = synth.make_fastq(nreads, bcs, tags, rng, cutsite=cutsite, newline=newline, readlen=readlen)
        kinds = [["N"], ["N", "lower"], ["N", "lower", "tab", "space"], ["N", "nul", "vt", "us", "hi"],
                 ["N", "cr"], ["N", "lf"], ["N", "lower", "tab", "nul", "cr", "lf", "hi", "space"]][it % 7]
        rate = float(rng.choice([0.0, 1e-5, 1e-4, 1e-3, 1e-2]))
        data = damage(fq, rng, kinds, rate)
        try:
            cnt = c_oracle.Counter(bcs, tags, cutsite)
        except AssertionError:
            continue                          # a tag that prefixes another: the reference refuses the set
        want, wtot = cnt.count(data, 5e9)
        want = want.tolist()
        wlines = c_oracle.count_lines(data)
        plan = matchset.plan(bcs, tags, cutsite)
        counting.load_plan(eng, plan, nrows=plan.barnum)
        dev, nb = eng.upload(data)
        eng.count_device(dev, nb)
        got = eng.read_matrix().tolist()
        tot = eng.file_totals()
        eng.device_free(dev)
        ok = got == want and tot[:3] == wtot and tot[3] == wlines
        tot2 = []
        cuts = sorted(int(c) for c in rng.integers(1, len(data), size=3))
        ok2 = counting.find_tags_bytes(data, bcs, tags, cutsite, totals=tot2, pieces=cuts) == want and tot2[:3] == wtot
        if not (ok and ok2):
            bad += 1
            print("MISMATCH seed", 7000 + seed0 + it, "nbar", nbar, "readlen", readlen, "kinds", kinds, "rate", rate,
                  "device", ok, "stream", ok2, tot, wtot, wlines, flush=True)
    eng.close()
    print("gpu_fuzz: %d iterations, %d mismatches, %.0f s" % (iters, bad, time.time() - t0))
    return bad


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    return 1 if run(iters, seed0) else 0


if __name__ == "__main__":
    sys.exit(main())
