#!/usr/bin/env python3
"""End-to-end rate of the barcode splitter (config-5 shape: PstI-MspI library, adapter
read-through, 96 barcodes) on a synthetic FASTQ on local disk: the streaming device path
against the CPU restatement of the reference's loop on a sample.

    python scripts/splitter_e2e.py [reads]
"""
import contextlib
import io
import json
import os
import random
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def make_reads(rng, barcodes, cutsite, adapter, n):
    from tagdigger_b200 import hostio
    full0 = adapter[0][0].replace("^", "")
    a0 = adapter[0][0][:adapter[0][0].find("^")] + adapter[0][1]
    rc = {bc: hostio.reverseComplement(bc) for bc in barcodes}
    out = []
    for i in range(n):
        bc = barcodes[rng.randrange(len(barcodes))]
        a1 = adapter[1][0][:adapter[1][0].find("^")] + adapter[1][1].replace("[barcode]", rc[bc])
        k = rng.random()
        insert = "".join(rng.choices("ACGT", k=rng.randint(20, 160)))
        if k < 0.35:
            tail = a1 + "".join(rng.choices("ACGT", k=40))                    # read-through into the adapter
        elif k < 0.40:
            tail = full0 + "".join(rng.choices("ACGT", k=60))                  # chimera
        elif k < 0.45:
            tail = a0
        else:
            tail = "".join(rng.choices("ACGT", k=100))
        head = bc + cutsite if rng.random() < 0.9 else "".join(rng.choices("ACGT", k=10))
        s = (head + insert + tail)[:100]
        q = "".join(rng.choices("ABCDEFGHIJ", k=len(s)))
        out.append("@INST:7:FC:1:%d:%d:%d 1:N:0:ATCACG\n%s\n+\n%s\n" % (1100 + i % 90, rng.randrange(30000), rng.randrange(30000), s, q))
    return "".join(out)


def main():
    from oracle import tagdigger_oracle as orc
    from tagdigger_b200 import hostio, splitter
    from tagdigger_b200 import synth
    import numpy as np
    reads = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
    rng = random.Random(5)
    adapter = hostio.adapters["PstI-MspI-Hall"]
    barcodes = synth.make_barcodes(96, np.random.default_rng(5))
    unit = 250_000
    text = make_reads(rng, barcodes, "TGCAG", adapter, unit)
    tmp = tempfile.mkdtemp(dir=os.environ.get("TDG_TMP", "/tmp"))
    inp = os.path.join(tmp, "reads.fastq")
    with open(inp, "w") as fh:
        for _ in range(reads // unit):
            fh.write(text)
    reads = reads // unit * unit
    outs = [os.path.join(tmp, "s%02d.fq" % i) for i in range(len(barcodes))]
    res = {"reads": reads, "input_MB": round(os.path.getsize(inp) / 1e6, 1), "host_cpus": os.cpu_count(), "runs": []}
    for rep in range(3):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            splitter.barcodeSplitter(inp, barcodes, outs, adapter=adapter)
        dt = time.perf_counter() - t0
        res["runs"].append({"seconds": round(dt, 3), "reads_per_s": round(reads / dt, 1)})
    res["output_MB"] = round(sum(os.path.getsize(o) for o in outs) / 1e6, 1)
    # CPU: the restated reference loop on the first 100,000 reads, and a byte-for-byte check of those
    sample = 100_000
    lines = text.splitlines(keepends=True)[:4 * sample]
    t0 = time.perf_counter()
    bufs = [[] for _ in barcodes]
    for b, ls, _ in orc.split_records(iter(lines), barcodes, "TGCAG", adapter):
        bufs[b].append("".join(x + "\n" for x in ls))
    dt = time.perf_counter() - t0
    res["cpu_port"] = {"reads": sample, "seconds": round(dt, 2), "reads_per_s": round(sample / dt, 1), "cores": 1}
    ok = True
    for o, b in zip(outs, bufs):
        want = "".join(b).encode()
        with open(o, "rb") as fh:
            ok &= fh.read(len(want)) == want
    res["prefix_identical_to_cpu_port"] = bool(ok)
    print(json.dumps(res, indent=1))
    for p in outs + [inp]:
        os.remove(p)
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
