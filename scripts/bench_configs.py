#!/usr/bin/env python3
"""Kernel-only timing of the counting path on the OTHER BASELINE.json shapes (not bench
lines: a check that no table shape or record length falls off a performance cliff).
Every run is verified twice: by construction on the whole image (counts >= the generator's expected
matrix, sum(counts) == tag hits, reads seen == reads generated) and EXACTLY against the C oracle on a
slice of it (every cell and the three totals equal).

    python scripts/bench_configs.py [reads] [shape ...]
"""
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

PEAK = 6545.9
try:
    PEAK = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
except (OSError, ValueError, KeyError):
    pass


def run(name, reads, general=False, exact_reads=2_000_000):
    import torch
    from oracle import c_oracle
    from tagdigger_b200 import _native, _synth_native, matchset, synth
    t0 = time.time()
    bcs, tags, cutsite, site, readlen, mix = synth.shape_tables(name)
    plan = matchset.plan(bcs, tags, cutsite)
    if general:
        os.environ["TDG_GENERAL"] = "1"
    else:
        os.environ.pop("TDG_GENERAL", None)
    eng = _native.Engine(0)
    matrix = torch.zeros((plan.barnum, plan.ntags), dtype=torch.int32, device="cuda")
    eng.set_tags(plan.tags.patterns, plan.tags.index, any_base=plan.tags.any_base)
    eng.bind_matrix(matrix.data_ptr(), plan.barnum, plan.ntags)
    eng.begin_file(plan.bar.patterns, plan.bar.index, plan.bar_tag_off, any_base=plan.bar.any_base)
    setup = time.time() - t0
    gen = _synth_native.Generator(bcs, tags, site, readlen=readlen, seed=11, **mix)
    # exact: a slice against the C oracle
    dev, nbytes = gen.generate(0, 777, exact_reads)
    img = np.empty(nbytes, dtype=np.uint8)
    eng.memcpy_d2h(img.ctypes.data, dev, nbytes)
    eng.count_device(dev, nbytes, 0, _native.TDG_PREV_NONE)
    tot = eng.file_totals()
    got = matrix.cpu().numpy().astype(np.int64)
    gen.free(0, dev)
    want, wtot = c_oracle.count_sharded(img, c_oracle.Counter(bcs, tags, cutsite))
    exact = bool((got == want).all()) and tot[:3] == wtot
    del img, got, want
    # speed: the whole image
    expected = torch.zeros_like(matrix)
    dev, nbytes = gen.generate(0, 0, reads, expected.data_ptr())
    torch.cuda.synchronize()
    for _ in range(3):
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
    p_bar, p_tag = tot[1] / tot[0], tot[2] / tot[0]
    alg = nbytes / reads + 32.0 * p_bar + 8.0 * p_tag
    out = {"config": name + (" (general matcher forced)" if general else ""), "reads": reads, "barcodes": len(bcs),
           "tags": len(tags), "tag_len": [min(map(len, tags)), max(map(len, tags))], "readlen": readlen,
           "bytes_per_read": round(nbytes / reads, 1), "kernel_ms": round(k, 3),
           "reads_per_s": round(reads / (k * 1e-3), 1), "stream_GBps": round(nbytes / (k * 1e-3) / 1e9, 1),
           "algorithmic_GBps": round(alg * reads / (k * 1e-3) / 1e9, 1),
           "frac_of_hbm_peak": round(alg * reads / (k * 1e-3) / 1e9 / PEAK, 4),
           "p_bar": round(p_bar, 3), "p_tag": round(p_tag, 3),
           "check": "ok" if ok else "FAILED", "exact_vs_c_oracle": "ok (%d reads)" % exact_reads if exact else "FAILED",
           "matrix_MB": round(plan.barnum * plan.ntags * 4 / 1e6, 1), "host_setup_s": round(setup, 1)}
    print(json.dumps(out), flush=True)
    gen.free(0, dev)
    eng.close()
    del matrix, expected
    torch.cuda.empty_cache()


def main():
    reads = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
    from tagdigger_b200 import synth
    names = sys.argv[2:] or list(synth.SHAPES)
    for name in names:
        general = name.endswith("+general")
        base = name[:-len("+general")] if general else name
        n = reads // 2 if synth.SHAPES[base]["readlen"] > 100 else reads
        run(base, n // 4 if general else n, general=general)


if __name__ == "__main__":
    main()
