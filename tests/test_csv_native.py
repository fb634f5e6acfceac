"""The native CSV writers (tdg_write_counts_csv / tdg_write_geno_csv through hostio.writeCounts /
writeDiploidGeno with an int32 ndarray) produce the bytes of the pure-Python csv.writer path,
which is itself pinned to the reference's output (tests/test_hostio_golden.py).  CPU only."""

import random

import numpy as np
import pytest

from tagdigger_b200 import hostio


@pytest.mark.parametrize("seed", range(8))
def test_native_csv_equals_python_csv(seed, in_tmp):
    rng = random.Random(seed)
    nmark = rng.choice([1, 2, 7, 60, 700])
    names = []
    for m in range(nmark):
        marker = rng.choice(["M%d" % m, "x,y%d" % m, 'q"%d' % m, " sp%d" % m, "é%d" % m])
        names += [marker + "_A_0", marker + "_C_1"]
    if rng.random() < 0.3:
        names.reverse()
    rows = rng.choice([1, 3, 40])
    samples = [rng.choice(["s%d", "a,b%d", 'q"q%d', " x%d ", "ü%d", "line\nbreak%d"]) % i for i in range(rows)]
    matrix = np.array([[rng.choice([0, 0, 0, 1, 9, 10, 99, 12345, 2147483647]) for _ in names] for _ in range(rows)],
                      dtype=np.int32)
    hostio.writeCounts("a.csv", matrix.tolist(), samples, names)
    hostio.writeCounts("b.csv", matrix, samples, names)
    assert open("a.csv", "rb").read() == open("b.csv", "rb").read()
    hostio.writeDiploidGeno("ga.csv", matrix.tolist(), samples, names)
    hostio.writeDiploidGeno("gb.csv", matrix, samples, names)
    assert open("ga.csv", "rb").read() == open("gb.csv", "rb").read()


def test_native_csv_large_and_errors(in_tmp, capsys):
    rng = np.random.default_rng(1)
    matrix = rng.integers(0, 5000, size=(37, 60000), dtype=np.int32)
    names = ["M%d_%s_%d" % (i // 2, "AC"[i % 2], i % 2) for i in range(60000)]
    samples = ["S%02d" % i for i in range(37)]
    hostio.writeCounts("big_native.csv", matrix, samples, names)
    hostio.writeCounts("big_python.csv", matrix.tolist(), samples, names)
    assert open("big_native.csv", "rb").read() == open("big_python.csv", "rb").read()
    # a marker without allele 1: message, no file -- on both paths
    bad = ["M0_A_0", "M1_A_0", "M1_C_1"]
    m3 = np.ones((2, 3), dtype=np.int32)
    hostio.writeDiploidGeno("x.csv", m3, ["a", "b"], bad)
    hostio.writeDiploidGeno("y.csv", m3.tolist(), ["a", "b"], bad)
    out = capsys.readouterr().out.splitlines()
    assert out[-1] == out[-2] == "'1' is not in list"
    import os
    assert not os.path.exists("x.csv") and not os.path.exists("y.csv")
    hostio.writeDiploidGeno("/nonexistent-dir/z.csv", m3[:, 1:], ["a", "b"], bad[1:])
    assert capsys.readouterr().out.strip() == "Could not write file /nonexistent-dir/z.csv."
