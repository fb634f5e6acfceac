"""Exact parity with the C oracle at the table shapes the bench and BASELINE.json's configs use
(run with -m gpu on a B200): 96-plex x 40,000 tags (config 2), blank barcode x 40,000 tags (config
3), 384-plex x 500,000 tags of 20-64 bp and a Stacks-style 80-140 bp set (config 4), 30-64 bp tags
(config 5's tag shape), ApeKI's two cut sites, 50 bp reads.  Every image is generated on the device
(the bench's generator), counted through the C ABI (tdg_count_device and tdg_submit), copied to the
host and counted there by oracle/oracle.c on all host cores: the matrices and the three totals must
be EQUAL (integer work: bit-exact)."""

import numpy as np
import pytest

from oracle import c_oracle
from tagdigger_b200 import _native, _synth_native, counting, matchset, synth

pytestmark = pytest.mark.gpu

READS = 2_000_000


def _image(eng, gen, first, nreads):
    dev, nbytes = gen.generate(0, first, nreads)
    img = np.empty(nbytes, dtype=np.uint8)
    eng.memcpy_d2h(img.ctypes.data, dev, nbytes)
    return dev, nbytes, img


@pytest.mark.parametrize("shape,npairs,general", [
    ("C2", None, False),
    ("C3", None, False),
    ("C4", None, False),
    ("C4-150", 100000, False),
    ("C4-stacks", 100000, False),
    ("C5", None, False),
    ("ApeKI", None, False),
    ("short", None, False),
    ("C2", None, True),                 # the general matcher forced on the headline shape
])
def test_shape_exact(shape, npairs, general, monkeypatch):
    if general:
        monkeypatch.setenv("TDG_GENERAL", "1")
    bcs, tags, cutsite, site, readlen, mix = synth.shape_tables(shape, npairs=npairs)
    plan = matchset.plan(bcs, tags, cutsite)
    eng = _native.Engine(0)
    try:
        counting.load_plan(eng, plan, nrows=plan.barnum)
        gen = _synth_native.Generator(bcs, tags, site, readlen=readlen, seed=101, **mix)
        nreads = READS if not general else READS // 4
        dev, nbytes, img = _image(eng, gen, 12345, nreads)
        want, wtot = c_oracle.count_sharded(img, c_oracle.Counter(bcs, tags, cutsite))
        assert wtot[0] == nreads and wtot[2] > nreads // 4          # the image exercises the tables
        # device-resident image (the bench's path)
        eng.count_device(dev, nbytes, 0, _native.TDG_PREV_NONE)
        tot = eng.file_totals()
        got = eng.read_matrix()
        gen.free(0, dev)
        assert tot[:3] == wtot
        assert (got == want).all()
        # the same image from host memory (tdg_submit: H2D pieces, carried partial lines)
        eng.zero_matrix()
        eng.reset_file()
        eng.submit(img)
        eng.end_file()
        assert eng.file_totals()[:3] == wtot
        assert (eng.read_matrix() == want).all()
    finally:
        eng.close()


def test_config2_ten_million_reads_exact():
    """The headline tables at 10 M reads (2.5 GB): every cell and total equal to the oracle's."""
    bcs, tags, cutsite, site, readlen, mix = synth.shape_tables("C2", seed=20162)
    plan = matchset.plan(bcs, tags, cutsite)
    eng = _native.Engine(0)
    try:
        counting.load_plan(eng, plan, nrows=plan.barnum)
        gen = _synth_native.Generator(bcs, tags, site, readlen=readlen, seed=20162)
        dev, nbytes, img = _image(eng, gen, 0, 10_000_000)
        eng.count_device(dev, nbytes, 0, _native.TDG_PREV_NONE)
        tot = eng.file_totals()
        got = eng.read_matrix()
        gen.free(0, dev)
        want, wtot = c_oracle.count_sharded(img, c_oracle.Counter(bcs, tags, cutsite))
        assert tot[:3] == wtot
        assert (got == want).all()
    finally:
        eng.close()
