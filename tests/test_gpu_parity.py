"""Parity of the CUDA path with the oracle (run with -m gpu on a B200).

Everything goes through the C ABI (ctypes -> libtagdigger_b200.so): whole files
via tdg_count_file, host images via tdg_submit, device-resident chunks via
tdg_count_device.  The bar is bit-exact counts (integer work)."""

import gzip
import os
import random

import numpy as np
import pytest

from conftest import file_bytes, load_golden, materialize
from helpers import fastq_of, line_soup, rand_seq, small_setup
from oracle import c_oracle
from oracle import tagdigger_oracle as orc
from tagdigger_b200 import _native, counting, matchset, synth

pytestmark = pytest.mark.gpu

FIND = load_golden("find_tags.json") + load_golden("find_tags_text.json")   # tassel_tagcount=True and text-mode cases included


@pytest.fixture(scope="module")
def eng():
    return counting.get_engine(0)


@pytest.mark.parametrize("i", range(len(FIND)))
def test_find_tags_fastq_golden(i, in_tmp):
    """The cases recorded from the reference itself, through the drop-in."""
    case = FIND[i]
    materialize(case["files"], in_tmp)
    try:
        ret, exc = counting.find_tags_fastq(*case["args"], **case["kwargs"]), None
    except matchset.DegenerateTree:
        assert case["exc"] is not None and case["exc"][0] in ("IndexError", "TypeError")
        return
    except BaseException as e:  # noqa: BLE001
        ret, exc = None, [type(e).__name__, str(e)]
    if case["exc"] is not None:
        assert exc is not None and exc[0] == case["exc"][0], (exc, case["exc"])
        if exc[0] == "AssertionError":
            assert exc[1] == case["exc"][1]
    else:
        assert exc is None, exc
        assert ret == case["ret"]


def _oracle(data, barcodes, tags, cutsite="TGCAG", maxreads=5e9):
    cnt = c_oracle.Counter(barcodes, tags, cutsite)
    m, tot = cnt.count(data, maxreads)
    return m.tolist(), tot, c_oracle.count_lines(data)


@pytest.mark.parametrize("cutsite,newline,lengths", [
    ("TGCAG", b"\n", None),
    ("TGCAG", b"\r\n", None),
    ("TGCAG", b"\r", None),
    ("CWGC", b"\n", (20, 64)),
    ("TGCAG", b"\n", (20, 64)),
    ("TGCAG", b"\n", (70, 150)),
    ("", b"\n", (25, 40)),
])
def test_synthetic_against_c_oracle(cutsite, newline, lengths):
    rng = np.random.default_rng(abs(hash((cutsite, newline, lengths))) % 2**32)
    site = orc.expand_cut_site(cutsite)[0]
    bcs = synth.make_barcodes(24, rng, cutsite=site)
    L = None if lengths is None else rng.integers(lengths[0], lengths[1] + 1, size=300)
    _, _, seqs = synth.make_marker_pairs(300, rng, cutsite=site, lengths=L)
    tags = [s for p in seqs for s in p]
    readlen = 100 if lengths is None or lengths[1] <= 64 else 180
    fq, truth = synth.make_fastq(60000, bcs, tags, rng, cutsite=site, newline=newline, readlen=readlen)
    want, wtot, wlines = _oracle(fq, bcs, tags, cutsite)
    tot = []
    got = counting.find_tags_bytes(fq, bcs, tags, cutsite, totals=tot)
    assert got == want
    assert tot[:3] == wtot and tot[3] == wlines
    assert (np.asarray(got) >= truth["expected"]).all()


@pytest.mark.parametrize("nbar,newline", [(384, b"\n"), (384, b"\r\n"), (200, b"\n")])
def test_many_barcodes(nbar, newline):
    """384-plex: the barcode table does not fit beside the kernel's rings (it is read through L1
    instead of shared memory) and many of its 256 buckets hold more than two patterns (the
    matcher's loop behind the two unrolled compares)."""
    rng = np.random.default_rng(384 + nbar + len(newline))
    site = orc.expand_cut_site("TGCAG")[0]
    bcs = synth.make_barcodes(nbar, rng, cutsite=site)
    _, _, seqs = synth.make_marker_pairs(200, rng, cutsite=site)
    tags = [s for p in seqs for s in p]
    fq, truth = synth.make_fastq(50000, bcs, tags, rng, cutsite=site, newline=newline, readlen=100)
    want, wtot, wlines = _oracle(fq, bcs, tags, "TGCAG")
    tot = []
    got = counting.find_tags_bytes(fq, bcs, tags, "TGCAG", totals=tot)
    assert got == want
    assert tot[:3] == wtot and tot[3] == wlines
    assert (np.asarray(got) >= truth["expected"]).all()
    assert sum(map(sum, got)) > 20000


@pytest.mark.parametrize("seed", range(12))
def test_line_soup(seed):
    """Arbitrary text: mixed line ends, blank lines, whitespace, truncated
    lines, no final newline -- the line index must track Python's."""
    r = random.Random(seed)
    cutsite = r.choice(["TGCAG", "CWGC", "TGCAG"])
    site = orc.expand_cut_site(cutsite)[0]
    barcodes, tags = small_setup(r, site)
    kinds = [("\n",), ("\r\n",), ("\r",), ("\n", "\r\n", "\r"), ("\n", "\r")][seed % 5]
    data = line_soup(r, barcodes, tags, site, r.randint(500, 4000), kinds,
                     final_newline=r.random() < 0.5, max_pad=r.choice([4, 40, 700]))
    want, wtot, wlines = _oracle(data, barcodes, tags, cutsite)
    assert want == orc.find_tags_text(data, barcodes, tags, cutsite)      # C oracle == Python oracle here
    tot = []
    got = counting.find_tags_bytes(data, barcodes, tags, cutsite, totals=tot)
    assert got == want
    assert tot[:3] == wtot and tot[3] == wlines
    # the same image in random pieces (carry-over of partial lines)
    cuts = sorted(r.sample(range(1, len(data)), min(7, len(data) - 1)))
    tot2 = []
    assert counting.find_tags_bytes(data, barcodes, tags, cutsite, totals=tot2, pieces=cuts) == want
    assert tot2 == tot


def test_unicode_whitespace_and_non_ascii():
    barcodes, tags = ["AACG"], ["TGCAGCCCC", "TGCAGAAAA"]
    reads = ["\u00a0AACGTGCAGCCCC", " \u3000 AACGTGCAGAAAAT", "AACGTGCAG\u00e9CCCC", "\u00e9AACGTGCAGCCCC",
             "AACGTGCAGCCCC\u2009", "aacgtgcagcccc", "\u2003\u2028AACGTGCAGCCCC",
             "\u0085\u1680\u205f\u202fAACGTGCAGCCCC"]
    data = fastq_of([s.encode("utf-8") for s in reads])
    data = data.replace(b"@r3", "@rü3".encode("utf-8"))
    want = orc.find_tags_text(data, barcodes, tags)
    assert c_oracle.find_tags_bytes(data, barcodes, tags) == want
    assert counting.find_tags_bytes(data, barcodes, tags) == want
    assert want == [[5, 1]]


KERNEL_TILES = [4608, 5632]      # csrc/tdg_kernel.cuh: bytes per warp tile (15 warps x 9 units; 12 warps x 11 units)


@pytest.mark.parametrize("tile", KERNEL_TILES + [_native.TDG_TILE_BYTES])
@pytest.mark.parametrize("delta", [-3, -2, -1, 0, 1, 2, 3, 17])
def test_device_chunk_sizes_around_tile_edges(eng, delta, tile):
    """tdg_count_device with n just below / at / above multiples of the tile
    size, and line ends falling exactly on tile and chunk boundaries."""
    r = random.Random(7 + delta)
    barcodes, tags = small_setup(r)
    for ntiles in (1, 2, 5):
        n = ntiles * tile + delta
        data = line_soup(r, barcodes, tags, "TGCAG", 3000, ("\n", "\r\n", "\r") if delta % 2 else ("\n",))
        data = (data * (n // len(data) + 1))[:n]
        # force interesting bytes at the edges
        for pos, ch in ((tile - 1, b"\n"), (tile, b"\r"), (n - 1, b"\r" if delta % 2 else b"\n")):
            if 0 <= pos < n:
                data = data[:pos] + ch + data[pos + 1:]
        want, wtot, wlines = _oracle(data, barcodes, tags)
        p = matchset.plan(barcodes, tags, "TGCAG")
        counting.load_plan(eng, p, nrows=p.barnum)
        dev, nb = eng.upload(data)
        try:
            eng.count_device(dev, nb)
            tot = eng.file_totals()
            got = eng.read_matrix().tolist()
            lines_only = eng.count_lines_device(dev, nb)
        finally:
            eng.device_free(dev)
        assert got == want
        assert tot[:3] == wtot
        last = data[-1:]
        # next_line = number of lines that have started = lines Python iterates
        assert tot[3] == wlines == lines_only[0]
        assert lines_only[1] == (1 if last == b"\n" else 2 if last == b"\r" else 3)


def test_chunk_chaining_matches_whole(eng):
    """A file given as several device chunks cut at line ends, chained on the
    device, with explicit line bases, equals the whole."""
    r = random.Random(99)
    barcodes, tags = small_setup(r, ntag=20)
    reads = [r.choice(barcodes) + r.choice(tags) + rand_seq(r, 20) for _ in range(5000)]
    data = fastq_of(reads)
    want, wtot, _ = _oracle(data, barcodes, tags)
    p = matchset.plan(barcodes, tags, "TGCAG")
    cuts = [0]
    for frac in (0.2, 0.5, 0.9):
        cuts.append(data.index(b"\n", int(len(data) * frac)) + 1)      # arbitrary LINE (not record) ends
    cuts.append(len(data))
    for mode in ("chained", "explicit"):
        counting.load_plan(eng, p, nrows=p.barnum)
        line = 0
        bufs = []
        for k, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
            dev, nb = eng.upload(data[a:b])
            bufs.append(dev)
            if mode == "chained" and k > 0:
                eng.count_device(dev, nb, _native.TDG_LINE_CHAINED, 0)
            else:
                eng.count_device(dev, nb, line, _native.TDG_PREV_NONE if k == 0 else _native.TDG_PREV_LF)
            line += data[a:b].count(b"\n")
        got = eng.read_matrix().tolist()
        tot = eng.file_totals()
        for d in bufs:
            eng.device_free(d)
        assert got == want, mode
        if mode == "chained":
            assert tot[:3] == wtot


@pytest.mark.parametrize("maxreads", [0, 1, 2.5, 7, 1000, 4999, 5000, 1e12])
def test_maxreads(maxreads):
    r = random.Random(5)
    barcodes, tags = small_setup(r)
    reads = [r.choice(barcodes) + r.choice(tags) + "ACGT" for _ in range(5000)]
    data = fastq_of(reads)
    want, wtot, _ = _oracle(data, barcodes, tags, maxreads=maxreads)
    tot = []
    assert counting.find_tags_bytes(data, barcodes, tags, maxreads=maxreads, totals=tot) == want
    assert tot[:3] == wtot


def test_files_plain_and_gzip(tmp_path):
    rng = np.random.default_rng(3)
    bcs = synth.make_barcodes(12, rng)
    _, _, seqs = synth.make_marker_pairs(80, rng)
    tags = [s for p in seqs for s in p]
    fq, _ = synth.make_fastq(30000, bcs, tags, rng)
    want, wtot, _ = _oracle(fq, bcs, tags)
    plain = str(tmp_path / "reads.fastq")
    with open(plain, "wb") as fh:
        fh.write(fq)
    gz = str(tmp_path / "reads.fastq.gz")
    with gzip.open(gz, "wb", compresslevel=1) as fh:
        fh.write(fq)
    multi = str(tmp_path / "multi.fq.GZ")          # two gzip members, upper-case suffix
    cut = fq.index(b"\n", len(fq) // 2) + 1
    with open(multi, "wb") as fh:
        fh.write(gzip.compress(fq[:cut]) + gzip.compress(fq[cut:]))
    from feed_check import bgzf_compress
    bg = str(tmp_path / "blocked.fastq.gz")         # BGZF: inflated in parallel by the host feeder
    with open(bg, "wb") as fh:
        fh.write(bgzf_compress(fq, block=20000))
    for path in (plain, gz, multi, bg):
        tot = []
        assert counting.find_tags_fastq(path, bcs, tags, totals=tot) == want, path
        assert tot == wtot
    notgz = str(tmp_path / "plain_named.gz")
    with open(notgz, "wb") as fh:
        fh.write(fq[:1000])
    with pytest.raises(gzip.BadGzipFile):
        counting.find_tags_fastq(notgz, bcs, tags)
    with pytest.raises(FileNotFoundError):
        counting.find_tags_fastq(str(tmp_path / "missing.fq"), bcs, tags)


def test_gzip_inflated_in_parallel(tmp_path, monkeypatch):
    """Ordinary gzip files go through the speculative parallel inflater (csrc/tdg_pgz.h): same
    counts as the oracle on the uncompressed bytes; damaged files raise what gzip.open raises."""
    import zlib
    monkeypatch.setenv("TDG_PGZ_MIN", "0")
    monkeypatch.setenv("TDG_PGZ_CHUNK", "100000")
    monkeypatch.setenv("TDG_IO_THREADS", "8")
    rng = np.random.default_rng(31)
    bcs = synth.make_barcodes(24, rng)
    _, _, seqs = synth.make_marker_pairs(200, rng)
    tags = [s for p in seqs for s in p]
    fq, _ = synth.make_fastq(120000, bcs, tags, rng)
    want, wtot, _ = _oracle(fq, bcs, tags)
    cut = fq.index(b"\n", len(fq) // 3) + 1
    blobs = {"l1.fq.gz": gzip.compress(fq, 1), "l9.fq.gz": gzip.compress(fq, 9),
             "members.fq.gz": gzip.compress(fq[:cut], 6) + gzip.compress(b"") + gzip.compress(fq[cut:], 4) + b"\0" * 100}
    for name, blob in blobs.items():
        path = str(tmp_path / name)
        with open(path, "wb") as fh:
            fh.write(blob)
        tot = []
        assert counting.find_tags_fastq(path, bcs, tags, totals=tot) == want, name
        assert tot == wtot
        tot = []
        assert counting.find_tags_fastq(path, bcs, tags, maxreads=50000, totals=tot) == _oracle(fq, bcs, tags, maxreads=50000)[0]
    good = blobs["l9.fq.gz"]
    bad = bytearray(good)
    bad[len(bad) // 2] ^= 0x40
    for name, blob, exc in (("trunc.gz", good[:len(good) // 2], EOFError), ("crc.gz", good[:-8] + b"\0" * 8, gzip.BadGzipFile),
                            ("flip.gz", bytes(bad), (zlib.error, gzip.BadGzipFile))):
        path = str(tmp_path / name)
        with open(path, "wb") as fh:
            fh.write(blob)
        with pytest.raises(exc):
            gzip.open(path, "rb").read()                  # the reference's reader
        with pytest.raises(exc):
            counting.find_tags_fastq(path, bcs, tags)


def test_small_chunks_force_many_pieces():
    """A context with tiny chunk_bytes: every piece boundary lands mid-line."""
    rng = np.random.default_rng(8)
    bcs = synth.make_barcodes(8, rng)
    _, _, seqs = synth.make_marker_pairs(40, rng)
    tags = [s for p in seqs for s in p]
    fq, _ = synth.make_fastq(20000, bcs, tags, rng)
    want, wtot, wlines = _oracle(fq, bcs, tags)
    eng = _native.Engine(0, chunk_bytes=50000)
    p = matchset.plan(bcs, tags, "TGCAG")
    counting.load_plan(eng, p, nrows=p.barnum)
    eng.submit(fq)
    eng.end_file()
    assert eng.read_matrix().tolist() == want
    assert eng.file_totals() == wtot + [wlines]
    eng.close()


def test_two_million_reads_config1_shape():
    """BASELINE config 1 shape (96 barcodes, 1000 biallelic pairs, PstI) at
    2 M reads against the C oracle, plus the by-construction lower bound."""
    rng = np.random.default_rng(20161)
    bcs = synth.make_barcodes(96, rng)
    _, _, seqs = synth.make_marker_pairs(1000, rng)
    tags = [s for p in seqs for s in p]
    fq, truth = synth.make_fastq(2000000, bcs, tags, rng)
    cnt = c_oracle.Counter(bcs, tags)
    want, wtot = cnt.count(fq)
    tot = []
    got = np.asarray(counting.find_tags_bytes(fq, bcs, tags, totals=tot))
    assert (got == want).all()
    assert tot[:3] == wtot
    assert (got >= truth["expected"]).all()
    assert got.sum() == tot[2]


@pytest.mark.parametrize("seg_tiles", [2, 5])
def test_multi_tile_segments_and_fix_pass(seg_tiles, monkeypatch):
    """Segments of several tiles (TDG_SEG_TILES test hook) on text whose FASTQ
    structure guess is wrong most of the time: exercises verify + fix pass, and
    a read limit that falls inside a segment."""
    monkeypatch.setenv("TDG_SEG_TILES", str(seg_tiles))
    eng = _native.Engine(0)
    r = random.Random(40 + seg_tiles)
    barcodes, tags = small_setup(r)
    p = matchset.plan(barcodes, tags, "TGCAG")
    for kinds, limit in ((("\n",), None), (("\n", "\r\n", "\r"), None), (("\n",), 777), (("\r",), 1500)):
        data = line_soup(r, barcodes, tags, "TGCAG", 9000, kinds)
        want, wtot, wlines = _oracle(data, barcodes, tags, maxreads=limit or 5e9)
        counting.load_plan(eng, p, nrows=p.barnum)
        dev, nb = eng.upload(data)
        eng.count_device(dev, nb, reads_limit=_native.limit_from_maxreads(limit or 5e9))
        got = eng.read_matrix().tolist()
        tot = eng.file_totals()
        eng.device_free(dev)
        assert got == want
        assert tot[:3] == wtot
        if limit is None:
            assert tot[3] == wlines
    # well-formed FASTQ shifted by one and two extra leading lines: every guess is
    # right except where the structure is ambiguous; result must still be exact
    reads = [r.choice(barcodes) + r.choice(tags) + rand_seq(r, 30) for _ in range(4000)]
    for lead in (b"", b"x\n", b"x\ny\n", b"x\ny\nz\n"):
        data = lead + fastq_of(reads)
        want, wtot, _ = _oracle(data, barcodes, tags)
        counting.load_plan(eng, p, nrows=p.barnum)
        dev, nb = eng.upload(data)
        eng.count_device(dev, nb)
        assert eng.read_matrix().tolist() == want
        assert eng.file_totals()[:3] == wtot
        eng.device_free(dev)
    eng.close()


def _device_count(eng, data, barcodes, tags, cutsite="TGCAG", limit=None):
    p = matchset.plan(barcodes, tags, cutsite)
    counting.load_plan(eng, p, nrows=p.barnum)
    dev, nb = eng.upload(data)
    try:
        eng.count_device(dev, nb, reads_limit=_native.limit_from_maxreads(limit or 5e9))
        return eng.read_matrix().tolist(), eng.file_totals()
    finally:
        eng.device_free(dev)


@pytest.mark.parametrize("general", ["0", "1"])
def test_pathological_line_shapes(general, monkeypatch):
    """Inputs far from FASTQ: thousands of empty lines per tile (several emission
    rounds per tile), lines longer than a tile, a file without any line end, tabs
    and other control characters inside lines -- with the fast matcher and with
    the general matcher (TDG_GENERAL test hook)."""
    monkeypatch.setenv("TDG_GENERAL", general)
    eng = _native.Engine(0)
    r = random.Random(123)
    barcodes, tags = small_setup(r)
    hit = barcodes[0] + tags[0]
    cases = []
    cases.append(b"\n" * 40000)                                             # every line empty
    cases.append((hit.encode() + b"\n") * 3000)                             # every line a read-shaped line
    cases.append(b"\n".join([b"A", hit.encode(), b"", b"#"] * 5000) + b"\n")  # 2-byte records
    cases.append(b"@h\n" + hit.encode() + b"G" * 30000 + b"\n+\n" + b"I" * 30000 + b"\n" + fastq_of([hit] * 50))
    cases.append(hit.encode() * 2000)                                        # no line end at all
    cases.append(b"x\n" + hit.encode())                                     # last line without line end
    cases.append(fastq_of([hit[:12] + "\t" + hit[12:], "\x0b" + hit, hit + "\x1f", hit[:3] + "\x00" + hit[3:], hit] * 200))
    cases.append(fastq_of([hit] * 300, newline=b"\r\n") + fastq_of([hit] * 300, newline=b"\r") + fastq_of([hit] * 300))
    for data in cases:
        want, wtot, wlines = _oracle(data, barcodes, tags)
        got, tot = _device_count(eng, data, barcodes, tags)
        assert got == want
        assert tot[:3] == wtot
        assert tot[3] == wlines
    eng.close()


@pytest.mark.parametrize("general", ["0", "1"])
def test_invalid_bases_and_short_reads_near_tags(general, monkeypatch):
    """N and other non-bases at every position around the barcode / tag spans,
    reads cut at every length, variable-length tags (a shorter tag may match
    although a later base is invalid)."""
    monkeypatch.setenv("TDG_GENERAL", general)
    eng = _native.Engine(0)
    r = random.Random(321)
    barcodes, tags = small_setup(r, nbar=8, ntag=30, tag_len=(6, 59))
    reads = []
    for _ in range(300):
        s = r.choice(barcodes) + r.choice(tags) + rand_seq(r, r.randint(0, 12))
        k = r.random()
        if k < 0.4:
            j = r.randrange(len(s))
            s = s[:j] + r.choice("Nn.-*xX@") + s[j + 1:]
        elif k < 0.7:
            s = s[:r.randint(0, len(s))]
        elif k < 0.8:
            s = s.lower()
        reads.append(s)
    data = fastq_of(reads)
    want, wtot, _ = _oracle(data, barcodes, tags)
    got, tot = _device_count(eng, data, barcodes, tags)
    assert got == want
    assert tot[:3] == wtot
    eng.close()


def test_general_matcher_table_shapes(monkeypatch):
    """Table shapes outside the fast matcher: tags longer than 64 bases, several
    length classes, long barcodes, blank barcode with empty cut site."""
    eng = _native.Engine(0)
    r = random.Random(77)
    for shape in ("long_tags", "short_tags", "long_barcodes", "blank"):
        if shape == "long_tags":
            barcodes, tags = small_setup(r, tag_len=(60, 150))
            cutsite = "TGCAG"
        elif shape == "short_tags":
            barcodes, tags = small_setup(r, ntag=10, tag_len=(1, 30))
            cutsite = "TGCAG"
        elif shape == "long_barcodes":
            barcodes, tags = small_setup(r)
            barcodes = [b + rand_seq(r, 14) for b in barcodes]
            cutsite = "TGCAG"
        else:
            barcodes, tags = [""], ["AACGC", "TTG"]
            cutsite = ""
        reads = []
        for _ in range(2000):
            s = r.choice(barcodes) + (r.choice(tags) if r.random() < 0.7 else cutsite + rand_seq(r, 40)) + rand_seq(r, 20)
            if r.random() < 0.1:
                j = r.randrange(len(s))
                s = s[:j] + "N" + s[j + 1:]
            reads.append(s)
        data = fastq_of(reads)
        want, wtot, _ = _oracle(data, barcodes, tags, cutsite)
        got, tot = _device_count(eng, data, barcodes, tags, cutsite)
        assert got == want, shape
        assert tot[:3] == wtot, shape
    eng.close()


def test_tassel_tagcount_weights(tmp_path, monkeypatch):
    """tassel_tagcount=True: count= weights from the header lines, Python int() grammar,
    ValueError for a header without a parsable count, maxreads, several GPU blocks."""
    monkeypatch.setattr(counting, "TASSEL_BLOCK", 300)
    r = random.Random(17)
    barcodes, tags = small_setup(r, nbar=5, ntag=14)
    recs = []
    for i in range(1500):
        s = r.choice(barcodes) + (r.choice(tags) if r.random() < 0.7 else "TGCAG" + rand_seq(r, 30)) + rand_seq(r, 10)
        w = r.choice(["1", "7", " 12 ", "+3", "-2", "1_000", "00042", str(r.randrange(10 ** 12))])
        recs.append("@tag%d length=64 count=%s\n%s\n+\n%s\n" % (i, w, s, "I" * len(s)))
    path = str(tmp_path / "tassel.fq")
    for maxreads in (5e9, 777):
        with open(path, "w") as fh:
            fh.write("".join(recs))
        want = orc.find_tags_fastq(path, barcodes, tags, maxreads=maxreads, tassel_tagcount=True)
        tot = []
        got = counting.find_tags_fastq(path, barcodes, tags, maxreads=maxreads, tassel_tagcount=True, totals=tot)
        assert got == want
        assert sum(map(sum, want)) != 0
    with open(path, "w") as fh:
        fh.write("".join(recs[:40]) + "@no weight here\nACGT\n+\nIIII\n")
    with pytest.raises(ValueError):
        counting.find_tags_fastq(path, barcodes, tags, tassel_tagcount=True)
    with pytest.raises(ValueError):
        orc.find_tags_fastq(path, barcodes, tags, tassel_tagcount=True)


@pytest.mark.parametrize("seed", range(16))
def test_fuzz_byte_soup(seed):
    """Random bytes over a small alphabet (bases, line ends of all kinds, FASTQ punctuation,
    whitespace, NUL, a non-ASCII byte pair), random lengths around the kernel's tile and halo
    sizes, one and several device chunks: bit-exact against the C oracle, totals and line
    counts included."""
    r = random.Random(1000 + seed)
    eng = _native.Engine(0)
    barcodes, tags = small_setup(r, nbar=5, ntag=10, tag_len=(4, 30))
    p = matchset.plan(barcodes, tags, "TGCAG")
    alphabet = [b"A", b"C", b"G", b"T"] * 6 + [b"\n"] * 3 + [b"\r", b"\r\n", b"@", b"+", b" ", b"\t", b"N", b"a", b"c",
                                                          b"\x00", b"\x0b", "é".encode(), b"I", b"#"]
    hits = [(r.choice(barcodes) + t).encode() for t in tags]
    for _ in range(25):
        n = r.choice([0, 1, 5, 127, 128, 129, 143, 144, 145, 175, 176, 177, 4607, 4608, 4609, 4736, 5631, 5632, 5633, 5760,
                      9216, 11264, 17000, 40000]) + r.randint(0, 3)
        parts = []
        size = 0
        while size < n:
            piece = r.choice(hits) if r.random() < 0.08 else r.choice(alphabet)
            parts.append(piece)
            size += len(piece)
        data = b"".join(parts)[:n]
        try:
            data.decode("utf-8")
        except UnicodeDecodeError:
            data = data[:-1]                       # do not cut the two-byte character in half
        want, wtot, wlines = _oracle(data, barcodes, tags)
        counting.load_plan(eng, p, nrows=p.barnum)
        if data:
            dev, nb = eng.upload(data)
            eng.count_device(dev, nb)
            got = eng.read_matrix().tolist()
            tot = eng.file_totals()
            eng.device_free(dev)
            assert got == want, (seed, n)
            assert tot[:3] == wtot and tot[3] == wlines
        # the same bytes through the streaming entry point, cut at random places
        cuts = sorted(r.sample(range(1, max(2, len(data))), min(3, max(0, len(data) - 1)))) if len(data) > 2 else []
        tot2 = []
        assert counting.find_tags_bytes(data, barcodes, tags, totals=tot2, pieces=cuts) == want
        assert tot2[:3] == wtot
    eng.close()
