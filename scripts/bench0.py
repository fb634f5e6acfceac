import sys, time, json
sys.path.insert(0, '.')
import numpy as np
from tagdigger_b200 import _native, counting, matchset, synth
rng = np.random.default_rng(20162)
bcs = synth.make_barcodes(96, rng)
_, _, seqs = synth.make_marker_pairs(20000, rng)
tags = [s for p in seqs for s in p]
t0=time.time()
fq, truth = synth.make_fastq(2000000, bcs, tags, rng)
print("gen", time.time()-t0, len(fq))
eng = counting.get_engine(0)
p = matchset.plan(bcs, tags, "TGCAG")
counting.load_plan(eng, p, nrows=p.barnum)
dev, n = eng.upload(fq)
for it in range(3):
    eng.zero_matrix(); eng.count_device(dev, n)
eng.sync()
eng.timing_begin()
K=10
for it in range(K):
    eng.zero_matrix(); eng.count_device(dev, n)
ms, nl = eng.timing_end()
tot = eng.file_totals()
print("kernel ms/step", ms/K, "GB/s", n*K/ms/1e6, "reads/s", 2000000*K/ms*1e3, tot)
m = eng.read_matrix()
print("sum", m.sum(), "expected>=", truth["expected"].sum())
