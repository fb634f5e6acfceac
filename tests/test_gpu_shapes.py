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


@pytest.mark.parametrize("maxlen", [65, 96, 97, 128, 129, 160, 161])
def test_long_tags_share_their_key(maxlen):
    """Tags longer than the 128-bit key (Stacks catalogs): alleles that differ only PAST base 64 share
    a key and are told apart by the tail compare of the matcher's long form; lengths on both sides of
    every word boundary of the tail, and one base beyond what the long form takes (161: the general
    matcher).  Reads: exact copies, copies with the last base changed, copies cut one base short,
    copies with an N inside the tail."""
    import random
    r = random.Random(maxlen)
    cutsite = "TGCAG"
    bcs = ["ACGT", "TTGCA", "GGATCC", "CATGA"]
    tags, reads = [], []
    for k in range(300):
        L = r.choice([maxlen, maxlen, max(66, maxlen - r.randint(0, 40))])
        body = "".join(r.choice("ACGT") for _ in range(L - len(cutsite)))
        t0 = cutsite + body
        pos = r.randint(64, L - 1)                       # SNP past the key
        t1 = t0[:pos] + r.choice([c for c in "ACGT" if c != t0[pos]]) + t0[pos + 1:]
        tags += [t0, t1]
    # drop tags that are prefixes of others (the reference's trie would refuse them)
    keep = [t for t in tags if not any(u != t and u.startswith(t) for u in tags)]
    tags = keep
    for _ in range(20000):
        t = r.choice(tags)
        b = r.choice(bcs)
        kind = r.random()
        s = t
        if kind < 0.15:
            s = t[:-1] + r.choice([c for c in "ACGT" if c != t[-1]])
        elif kind < 0.25:
            s = t[:-1]
        elif kind < 0.35:
            p = r.randint(64, len(t) - 1)
            s = t[:p] + "N" + t[p + 1:]
        tail = "".join(r.choice("ACGT") for _ in range(r.randint(0, 12))) if kind >= 0.25 or kind < 0.15 else ""
        reads.append(b + s + tail)
    fq = "".join("@r%d\n%s\n+\n%s\n" % (i, s, "I" * len(s)) for i, s in enumerate(reads)).encode()
    want, wtot = c_oracle.Counter(bcs, tags, cutsite).count(fq)
    assert wtot[2] > 5000
    tot = []
    got = np.asarray(counting.find_tags_bytes(fq, bcs, tags, cutsite, totals=tot))
    assert tot[:3] == wtot
    assert (got == want).all()
