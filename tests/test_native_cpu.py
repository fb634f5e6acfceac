"""CPU-side checks of the native library: it builds/loads, exports every symbol
the header declares and refuses to count without a GPU; and the product's packed
tables + per-read match code (csrc/tdg_tables.h, tdg_match.h, compiled for the host
by the TEST harness tests/native/table_check.cpp) agree with the oracle.  No
compute call needs a GPU here."""

import os
import random
import re

import pytest

from conftest import REPO, file_bytes, load_golden
from helpers import rand_seq
from oracle import tagdigger_oracle as orc
from table_check import TableError, Tables
from tagdigger_b200 import _native, counting, matchset


def test_header_symbols_are_exported():
    hdr = open(os.path.join(REPO, "include", "tagdigger_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(tdg_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    L = _native.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.tdg_abi_version() == _native.ABI_VERSION == 2


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_native.TdgError) as e:
        _native.Engine()
    assert "no CPU counting path" in str(e.value)
    with pytest.raises(_native.TdgError):
        counting.find_tags_fastq(__file__, ["AACG"], ["TGCAGCCCC"])


def test_table_builders_refuse_bad_sets():
    t = Tables()
    with pytest.raises(TableError):
        t.set_tags(["ACGT", "ACGTA"])          # not prefix-free
    with pytest.raises(TableError):
        t.set_tags(["ACGT", "ACGT"])           # duplicate
    with pytest.raises(TableError):
        t.set_tags(["ACNT"])
    with pytest.raises(TableError):
        t.set_tags([])
    t.set_tags(["ACGT"])
    with pytest.raises(TableError):
        t.set_bars(["A" * 33], [0], [33])
    with pytest.raises(TableError):
        t.set_bars(["AC", "ACG"], [0, 1], [2, 3])


def test_probe_sequences_end_in_their_first_line():
    """Biallelic pairs fill their slot pair exactly; only pairs that something was
    stored beyond carry the 'more' flag, and they are few at the table's load."""
    r = random.Random(3)
    tags = []
    for _ in range(3000):
        a = "TGCAG" + rand_seq(r, 59)
        j = r.randrange(40, 64)
        b = a[:j] + ("A" if a[j] != "A" else "C") + a[j + 1:]
        tags += [a, b]
    t = Tables()
    t.set_tags(tags)
    st = t.stats()
    assert st["used"] == len(tags) and st["slots"] >= 8 * len(tags)
    assert st["more"] < 0.03 * len(tags)
    for k, s in enumerate(tags[:200]):
        t.set_bars(["AC"], [0], [2])
        assert t.match("AC" + s + "GG") == k


def _expected(seq, bt, tt, offs, ntags):
    b = orc.lookup(seq, bt)
    if b < 0:
        return -2
    t = orc.lookup(seq[offs[b]:], tt)
    return -1 if t < 0 else b * ntags + t


def _host_engine(barcodes, tags, cutsite):
    t = Tables()
    t.load_plan(matchset.plan(barcodes, tags, cutsite))
    return t


FIND = [c for c in load_golden("find_tags.json")
        if c["exc"] is None and not c["kwargs"].get("tassel_tagcount")]


@pytest.mark.parametrize("i", range(len(FIND)))
def test_selftest_match_on_golden_reads(i):
    """Every sequence line of every golden find_tags case, one by one."""
    case = FIND[i]
    name, barcodes, tags = case["args"]
    raw = file_bytes(case["files"][name])
    if name[-2:].lower() == "gz":
        import gzip
        raw = gzip.decompress(raw)
    cutsite = case["kwargs"].get("cutsite", "TGCAG")
    bt, tt, offs, barnum, ntags = orc.prepare(barcodes, tags, cutsite)
    if bt.root_is_leaf or tt.root_is_leaf:
        pytest.skip("degenerate tree")
    eng = _host_engine(barcodes, tags, cutsite)
    lines = raw.decode("utf-8").replace("\r\n", "\n").replace("\r", "\n").split("\n")
    for n, line in enumerate(lines):
        if n % 4 == 1:
            assert eng.match(line) == _expected(line.strip().upper(), bt, tt, offs, ntags), line


@pytest.mark.parametrize("seed", range(120))
def test_selftest_match_random_sets(seed):
    r = random.Random(1000 + seed)
    cutsite = r.choice(["TGCAG", "CWGC", "", "CATGG", "TGCAGG", "RY"])
    barcodes = [rand_seq(r, r.randint(0 if r.random() < 0.05 else 1, 9)) for _ in range(r.randint(1, 12))]
    if r.random() < 0.1:
        barcodes = [""]
    sites = orc.expand_cut_site(cutsite)
    maxl = r.choice([8, 20, 40, 70, 150])
    tags = []
    for _ in range(r.randint(1, 30)):
        t = rand_seq(r, r.randint(1, maxl))
        if r.random() < 0.7:
            t = r.choice(sites) + t
        tags.append(t)
    if r.random() < 0.2:
        tags.append(tags[0][:max(1, len(tags[0]) // 2)])
    if r.random() < 0.2:
        tags.append(tags[0])
    try:
        bt, tt, offs, barnum, ntags = orc.prepare(barcodes, tags, cutsite)
        want_exc = None
    except (AssertionError, IndexError) as e:
        want_exc = (type(e), str(e))
    if want_exc is None and (bt.root_is_leaf or tt.root_is_leaf):
        with pytest.raises(matchset.DegenerateTree):
            matchset.plan(barcodes, tags, cutsite)
        return
    if want_exc is not None:
        with pytest.raises(want_exc[0]) as e:
            matchset.plan(barcodes, tags, cutsite)
        assert str(e.value) == want_exc[1]
        return
    eng = _host_engine(barcodes, tags, cutsite)
    for _ in range(60):
        k = r.random()
        bc, site, tg = r.choice(barcodes), r.choice(sites), r.choice(tags)
        stripped = tg[len(site):] if tg.startswith(site) else tg
        if k < 0.5:
            rd = bc + site + stripped + rand_seq(r, r.randint(0, 20))
        elif k < 0.6:
            rd = bc + tg + rand_seq(r, 5)
        elif k < 0.7:
            rd = bc + site + rand_seq(r, 30)
        elif k < 0.8:
            rd = rand_seq(r, 50)
        else:
            full = bc + site + stripped
            rd = full[:r.randint(0, len(full))]
        if r.random() < 0.1 and rd:
            j = r.randrange(len(rd))
            rd = rd[:j] + r.choice("Nn.x ") + rd[j + 1:]
        if r.random() < 0.1:
            rd = rd.lower()
        if r.random() < 0.1:
            rd = r.choice([" ", "\t", "  ", "\x0b\x1c", " ", "  "]) + rd + r.choice(["", " ", "\r"])
        assert eng.match(rd) == _expected(rd.strip().upper(), bt, tt, offs, ntags), rd


@pytest.mark.parametrize("seed", range(40))
def test_matchset_equals_trie(seed):
    """Effective sets against the oracle's trie, including the exact
    AssertionError text."""
    r = random.Random(seed)
    for _ in range(30):
        pool = [rand_seq(r, r.randint(0 if r.random() < 0.1 else 1, 5)) for _ in range(6)]
        pats = []
        for _ in range(r.randint(1, 10)):
            p = r.choice(pool)
            if r.random() < 0.4:
                p += rand_seq(r, r.randint(1, 3))
            pats.append(p)
        numseq = r.choice([len(pats), max(1, len(pats) // 2)])
        try:
            trie, want_exc = orc.build_trie(pats, numseq), None
        except (AssertionError, IndexError) as e:
            trie, want_exc = None, (type(e).__name__, str(e))
        try:
            eff, got_exc = matchset.effective_set(pats, numseq), None
        except (AssertionError, IndexError) as e:
            eff, got_exc = None, (type(e).__name__, str(e))
        except matchset.DegenerateTree:
            assert trie is not None and trie.root_is_leaf
            continue
        assert want_exc == got_exc, (pats, numseq)
        if eff is None:
            continue
        for _ in range(40):
            q = r.choice(pats)[:r.randint(0, 8)] + rand_seq(r, r.randint(0, 4))
            want = orc.lookup(q, trie)
            if eff.any_base:
                got = 0 if q[:1] in ("A", "C", "G", "T") else -1
            else:
                hits = [i for p, i in zip(eff.patterns, eff.index) if q.startswith(p)]
                assert len(hits) <= 1
                got = hits[0] if hits else -1
            assert got == want, (pats, numseq, q)


def test_limit_from_maxreads():
    f = _native.limit_from_maxreads
    assert f(5e9) == 5000000000
    assert f(0) == 1 and f(-3) == 1 and f(0.2) == 1
    assert f(2.5) == 3 and f(3) == 3
    assert f(float("inf")) == _native.NO_LIMIT
