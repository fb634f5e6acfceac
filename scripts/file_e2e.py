#!/usr/bin/env python3
"""End-to-end rate of the FILE entry point (tdg_count_file: read()/inflate on host threads ->
pinned buffers -> H2D -> kernel), plain and gzip, on a synthetic FASTQ written to local disk.

    python scripts/file_e2e.py [reads] [gz_reads]
"""
import gzip
import json
import os
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402


def main():
    import bench
    from tagdigger_b200 import _native, _synth_native, counting, matchset
    reads = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
    gz_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
    bcs, tags = bench.workload_tables()
    plan = matchset.plan(bcs, tags, bench.CUTSITE)
    eng = _native.Engine(0)
    gen = _synth_native.Generator(bcs, tags, bench.CUTSITE, readlen=bench.READLEN, seed=bench.SEED)
    dev, nbytes = gen.generate(0, 0, reads)
    host = np.empty(nbytes, dtype=np.uint8)
    eng.memcpy_d2h(host.ctypes.data, dev, nbytes)
    gen.free(0, dev)
    tmp = tempfile.mkdtemp(dir=os.environ.get("TDG_TMP", "/tmp"))
    plain = os.path.join(tmp, "reads.fastq")
    host.tofile(plain)
    cut = bytes(host[:int(nbytes * gz_reads / reads)])
    cut = cut[:cut.rfind(b"\n@") + 1]
    gz = os.path.join(tmp, "reads_small.fastq.gz")
    with gzip.open(gz, "wb", compresslevel=1) as fh:
        fh.write(cut)
    multi = os.path.join(tmp, "reads_members.fastq.gz")      # many gzip members (bgzip/pigz -i style)
    with open(multi, "wb") as fh:
        step = 1 << 22
        for i in range(0, len(cut), step):
            fh.write(gzip.compress(cut[i:i + step], compresslevel=1))
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from feed_check import bgzf_compress
    bg = os.path.join(tmp, "reads_bgzf.fastq.gz")
    with open(bg, "wb") as fh:
        fh.write(bgzf_compress(cut, level=1))
    del host
    counting.load_plan(eng, plan, nrows=plan.barnum)
    out = {}
    for name, path, isgz in (("plain", plain, False), ("plain_again", plain, False), ("gzip_single_member", gz, True),
                             ("gzip_many_members", multi, True), ("bgzf", bg, True), ("bgzf_again", bg, True)):
        eng.zero_matrix()
        eng.reset_file()
        t0 = time.perf_counter()
        tot = eng.count_file(path, isgz)
        m = eng.read_matrix()
        dt = time.perf_counter() - t0
        out[name] = {"reads": tot[0], "seconds": round(dt, 3), "reads_per_s": round(tot[0] / dt, 1),
                     "file_MB": round(os.path.getsize(path) / 1e6, 1), "tag_hits": int(m.sum())}
    print(json.dumps(out, indent=1))
    for p in (plain, gz, multi, bg):
        os.remove(p)
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
