#!/usr/bin/env python3
"""Differential fuzz of the device gzip feed's per-lane inflater on the CPU (tests/native/gzlane_check.cpp runs
tagdigger_b200/csrc/tdg_gzlane.h + tdg_gzchain.h one lane after the other): random FASTQ-like streams --
levels, strategies, sync / full flush points, two members -- optionally damaged (bit flips, overwritten runs,
truncation, inserted bytes), at random chunk geometries, some lanes made blind (the host fills in).  Undamaged
streams must come out exactly; damaged ones must give Python's bytes, or stop with a correct prefix (of what a
sequential inflater hands out) on the path that leads to the host reader / zlib.  Never wrong bytes.

    python scripts/gzlane_fuzz.py [cases] [seed]
"""
import gzip
import os
import random
import sys
import time
import zlib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def main():
    from gzlane_check import inflate
    from test_feed_cpu import _fastq_like
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 12345
    r = random.Random(seed)
    base = _fastq_like(41, 3 << 20)
    bad = 0
    kinds = {}
    t0 = time.time()
    for case in range(cases):
        n = r.randint(100000, len(base))
        data = base[:n]
        level = r.choice((1, 4, 6, 9))
        style = r.random()
        if style < 0.25:
            cut = r.randint(1, n - 1)
            blob = bytearray(gzip.compress(data[:cut], level) + gzip.compress(data[cut:], level))
        elif style < 0.4:
            co = zlib.compressobj(level, zlib.DEFLATED, 31, 9, r.choice((zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY,
                                                                         zlib.Z_RLE, zlib.Z_FIXED)))
            parts, pos = [], 0
            while pos < n:
                step = r.randint(1000, 400000)
                parts.append(co.compress(data[pos:pos + step]))
                if r.random() < 0.5:
                    parts.append(co.flush(r.choice((zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH))))
                pos += step
            parts.append(co.flush())
            blob = bytearray(b"".join(parts))
        else:
            blob = bytearray(gzip.compress(data, level))
        kind = r.choice(("flip", "run", "truncate", "insert", "none", "none"))
        if kind == "flip":
            for _ in range(r.randint(1, 3)):
                blob[r.randrange(len(blob))] ^= 1 << r.randrange(8)
        elif kind == "run":
            at = r.randrange(len(blob))
            blob[at:at + r.randint(1, 300)] = r.randbytes(r.randint(1, 300))
        elif kind == "truncate":
            del blob[r.randint(len(blob) // 2, len(blob) - 1):]
        elif kind == "insert":
            at = r.randrange(len(blob))
            blob[at:at] = r.randbytes(r.randint(1, 50))
        blob = bytes(blob)
        try:
            want = gzip.decompress(blob)
        except Exception:  # noqa: BLE001
            want = None
        seq = b""                                  # what a sequential inflater hands out before it notices the damage
        try:
            rest = blob
            while rest:
                d = zlib.decompressobj(31)
                seq += d.decompress(rest)
                if not d.eof:
                    break
                rest = d.unused_data
                if rest[:2] != b"\x1f\x8b":
                    break
        except zlib.error:
            pass
        chunk = r.choice((4096, 8192, 1 << 15, 1 << 16, 1 << 17))
        out, code, info = inflate(blob, chunk=chunk, max_chunks=r.randint(1, 60), cap=2 * len(base) + 100,
                                  blind_every=r.choice((0, 0, 0, 5, 11)))
        if code >= 0:
            ok = (want is None or out == want) and (want is not None or seq.startswith(out) or out.startswith(seq))
        else:
            ok = code in (-1, -2, -10)
            if ok and code != -10:
                truth = want if want is not None else seq
                ok = truth.startswith(out) or (want is None and out.startswith(seq))
        if kind == "none" and (code < 0 or out != data):
            ok = False
        k = kinds.setdefault(kind, [0, 0, 0])
        k[0] += 1
        k[1] += code >= 0
        k[2] += info["repairs"]
        if not ok:
            bad += 1
            print("MISMATCH case", case, kind, level, chunk, code, info, len(out), None if want is None else len(want))
    print("%d streams (seed %d), %d mismatches, %d s" % (cases, seed, bad, round(time.time() - t0)))
    for kind, (n, whole, rep) in sorted(kinds.items()):
        print("  %-8s %4d streams, %4d read to the end by the feed, %d chunks inflated by the host in place of a blind lane" % (kind, n, whole, rep))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
