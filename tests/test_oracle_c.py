"""The C half of the oracle against the golden vectors and the Python oracle."""

import gzip
import random

import numpy as np
import pytest

from conftest import file_bytes, load_golden
from oracle import c_oracle
from oracle import tagdigger_oracle as orc

FIND = [c for c in load_golden("find_tags.json")
        if not c["kwargs"].get("tassel_tagcount")]


def _image(case):
    """Uncompressed bytes the reference iterated over, or None if it could not open the file."""
    name = case["args"][0]
    if name not in case["files"]:
        return None
    raw = file_bytes(case["files"][name])
    if name[-2:].lower() == "gz":
        try:
            return gzip.decompress(raw)
        except OSError:
            return None
    return raw


@pytest.mark.parametrize("i", range(len(FIND)))
def test_c_oracle_golden(i):
    case = FIND[i]
    data = _image(case)
    if data is None or data[:2] == b"\x1f\x8b":
        pytest.skip("file-open error case (host layer, not the loop)")
    _, barcodes, tags = case["args"]
    try:
        got = c_oracle.find_tags_bytes(data, barcodes, tags, **case["kwargs"])
        exc = None
    except (AssertionError, IndexError, TypeError) as e:
        got, exc = None, [type(e).__name__, str(e)]
    if case["exc"] is not None:
        assert exc is not None and exc[0] == case["exc"][0]
        if exc[0] == "AssertionError":
            assert exc[1] == case["exc"][1]
    else:
        assert exc is None, exc
        assert got == case["ret"]


def test_c_oracle_matches_python_oracle_on_synthetic():
    from tagdigger_b200 import synth
    rng = np.random.default_rng(11)
    for cutsite, nl in (("TGCAG", b"\n"), ("CWGC", b"\r\n"), ("TGCAG", b"\r")):
        sites = orc.expand_cut_site(cutsite)
        bcs = synth.make_barcodes(16, rng, cutsite=sites[0])
        _, _, seqs = synth.make_marker_pairs(100, rng, cutsite=sites[0],
                                             lengths=rng.integers(20, 80, size=100))
        tags = [s for p in seqs for s in p]
        fq, _ = synth.make_fastq(6000, bcs, tags, rng, cutsite=sites[0], newline=nl)
        tot = [0, 0, 0]
        want = orc.find_tags_text(fq, bcs, tags, cutsite=cutsite, totals=tot)
        cnt = c_oracle.Counter(bcs, tags, cutsite)
        got, gtot = cnt.count(fq)
        assert got.tolist() == want
        assert gtot == tot
        assert c_oracle.count_lines(fq) == len(fq.decode().splitlines()) or nl != b"\n"


def test_c_oracle_sharded_count_is_exact():
    """Counting line-aligned shards with first_line set reproduces the whole."""
    from tagdigger_b200 import synth
    rng = np.random.default_rng(12)
    bcs = synth.make_barcodes(8, rng)
    _, _, seqs = synth.make_marker_pairs(50, rng)
    tags = [s for p in seqs for s in p]
    fq, _ = synth.make_fastq(3000, bcs, tags, rng)
    cnt = c_oracle.Counter(bcs, tags)
    whole, _ = cnt.count(fq)
    cut = fq.index(b"\n", len(fq) // 3) + 1
    first = c_oracle.count_lines(fq[:cut])
    a, ta = cnt.count(fq[:cut])
    b, _ = cnt.count(fq[cut:], first_line=first, reads_before=ta[0])
    assert (a + b == whole).all()
