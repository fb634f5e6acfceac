#!/usr/bin/env python3
"""Kernel-only timing of the counting path on the OTHER BASELINE.json shapes (not bench
lines: a check that no table shape or record length falls off a performance cliff).
Every run is verified by construction (counts >= the generator's expected matrix,
sum(counts) == tag hits, reads seen == reads generated).

    python scripts/bench_configs.py [reads]
"""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402


def run(name, nbar, npairs, readlen, reads, lengths=None, blank=False, cutsite="TGCAG", site=None, **mix):
    import torch
    from tagdigger_b200 import _native, _synth_native, matchset, synth
    rng = np.random.default_rng(7)
    site = site or cutsite            # the concrete site written into tags and reads (IUPAC cut sites)
    bcs = [""] if blank else synth.make_barcodes(nbar, rng, cutsite=site)
    t0 = time.time()
    lens = None if lengths is None else rng.integers(lengths[0], lengths[1] + 1, size=npairs)
    mnames, _, seqs = synth.make_marker_pairs(npairs, rng, cutsite=site, lengths=lens)
    tags = [s for p in seqs for s in p]
    if lengths is not None and lengths[0] != lengths[1]:
        # random variable-length tags overlap now and then: drop those markers as the script would
        import contextlib
        import io
        from tagdigger_b200 import hostio
        names = ["%s_%d" % (m, k) for m in mnames for k in (0, 1)]
        with contextlib.redirect_stdout(io.StringIO()):
            tags = hostio.sanitizeTags([names, tags])[1]
    plan = matchset.plan(bcs, tags, cutsite)
    eng = _native.Engine(0)
    matrix = torch.zeros((plan.barnum, plan.ntags), dtype=torch.int32, device="cuda")
    eng.set_tags(plan.tags.patterns, plan.tags.index, any_base=plan.tags.any_base)
    eng.bind_matrix(matrix.data_ptr(), plan.barnum, plan.ntags)
    eng.begin_file(plan.bar.patterns, plan.bar.index, plan.bar_tag_off, any_base=plan.bar.any_base)
    setup = time.time() - t0
    gen = _synth_native.Generator(bcs, tags, site, readlen=readlen, seed=11, **mix)
    expected = torch.zeros_like(matrix)
    dev, nbytes = gen.generate(0, 0, reads, expected.data_ptr())
    torch.cuda.synchronize()
    for _ in range(2):
        eng.zero_matrix()
        eng.count_device(dev, nbytes, 0, _native.TDG_PREV_NONE)
    eng.sync()
    eng.reset_file()
    eng.timing_begin()
    steps = 3
    for _ in range(steps):
        eng.zero_matrix()
        eng.count_device(dev, nbytes, 0, _native.TDG_PREV_NONE)
    ms, n = eng.timing_end()
    tot = eng.file_totals()
    ok = bool((matrix >= expected).all().item()) and int(matrix.sum(dtype=torch.int64).item()) == tot[2] // steps \
        and tot[0] // steps == reads
    k = ms / n
    out = {"config": name, "reads": reads, "barcodes": len(bcs), "tags": len(tags), "readlen": readlen,
           "bytes_per_read": round(nbytes / reads, 1), "kernel_ms": round(k, 3),
           "reads_per_s": round(reads / (k * 1e-3), 1), "stream_GBps": round(nbytes / (k * 1e-3) / 1e9, 1),
           "p_bar": round(tot[1] / tot[0], 3), "p_tag": round(tot[2] / tot[0], 3), "check": "ok" if ok else "FAILED",
           "host_setup_s": round(setup, 1)}
    print(json.dumps(out), flush=True)
    gen.free(0, dev)
    eng.close()
    del matrix, expected
    torch.cuda.empty_cache()


def main():
    reads = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
    run("C2 96-plex, 40k tags x 64 bp, 100 bp reads", 96, 20000, 100, reads)
    run("C3 pre-split (blank barcode), 40k tags, 100 bp reads", 1, 20000, 100, reads, blank=True, p_nobar=0.05, p_unknown=0.30)
    run("C4 384-plex, 500k tags of 20-64 bp, 100 bp reads", 384, 250000, 100, reads, lengths=(20, 64))
    run("C4 384-plex, 500k tags of 20-64 bp, 150 bp reads", 384, 250000, 150, reads // 2, lengths=(20, 64))
    run("C5-like 96-plex, 40k tags of 30-64 bp, 100 bp reads", 96, 20000, 100, reads, lengths=(30, 64))
    run("ApeKI (CWGC, two cut sites) 96-plex, 40k tags, 100 bp reads", 96, 20000, 100, reads, cutsite="CWGC", site="CAGC")
    run("short reads: 96-plex, 40k tags of 30 bp, 50 bp reads", 96, 20000, 50, reads, lengths=(30, 30))


if __name__ == "__main__":
    main()
