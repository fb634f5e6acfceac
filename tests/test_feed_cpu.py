"""The host feeder of tdg_count_file (tagdigger_b200/csrc/tdg_feed.h: parallel pread, parallel
BGZF inflate, zlib for ordinary gzip) delivers exactly the bytes Python reads from the same file.
CPU only, through the test harness tests/native/feed_check.cpp."""

import gzip
import os
import random

import pytest

from feed_check import FeedError, bgzf_block, bgzf_compress, read_file


def _data(seed, n):
    r = random.Random(seed)
    recs = []
    size = 0
    while size < n:
        s = "".join(r.choice("ACGT") for _ in range(r.randint(30, 120)))
        rec = "@r%d\n%s\n+\n%s\n" % (len(recs), s, "I" * len(s))
        recs.append(rec)
        size += len(rec)
    return "".join(recs).encode()


@pytest.mark.parametrize("chunk", [1 << 16, 100000, 1 << 20, 9 << 20])
def test_plain_parallel_pread(tmp_path, chunk):
    data = _data(1, 30 << 20) if chunk >= 1 << 20 else _data(1, 3 << 20)
    p = str(tmp_path / "x.fq")
    open(p, "wb").write(data)
    got, mode = read_file(p, False, chunk)
    assert mode == "plain" and got == data
    open(p, "wb").close()
    assert read_file(p, False, chunk)[0] == b""


@pytest.mark.parametrize("chunk", [70000, 1 << 20, 5 << 20])
def test_bgzf_parallel_inflate(tmp_path, chunk, monkeypatch):
    monkeypatch.setenv("TDG_IO_THREADS", "4")
    data = _data(2, 12 << 20)
    p = str(tmp_path / "x.fq.gz")
    open(p, "wb").write(bgzf_compress(data))
    assert gzip.open(p, "rb").read() == data                  # what the reference would read
    got, mode = read_file(p, True, chunk)
    assert mode == "bgzf" and got == data
    # without the end-of-file marker block, and with tiny members
    open(p, "wb").write(bgzf_compress(data[:200000], block=777, eof_marker=False))
    got, mode = read_file(p, True, chunk)
    assert mode == "bgzf" and got == data[:200000]


def test_bgzf_then_ordinary_members_and_small_chunks(tmp_path, monkeypatch):
    monkeypatch.setenv("TDG_IO_THREADS", "3")
    data = _data(3, 2 << 20)
    half = len(data) // 2
    p = str(tmp_path / "mixed.gz")
    open(p, "wb").write(bgzf_compress(data[:half], eof_marker=False) + gzip.compress(data[half:]))
    assert gzip.open(p, "rb").read() == data
    got, mode = read_file(p, True, 1 << 20)
    assert got == data and mode == "zlib"                     # continued sequentially
    got, mode = read_file(p, True, 4096)                      # chunk smaller than a member: zlib from the start
    assert got == data and mode == "zlib"


def test_ordinary_gzip_single_and_multi_member(tmp_path):
    data = _data(4, 3 << 20)
    p = str(tmp_path / "x.gz")
    open(p, "wb").write(gzip.compress(data))
    got, mode = read_file(p, True, 1 << 20)
    assert got == data and mode == "zlib"
    open(p, "wb").write(gzip.compress(data[:1000]) + gzip.compress(b"") + gzip.compress(data[1000:]))
    assert read_file(p, True, 1 << 18)[0] == data
    open(p, "wb").write(data[:5000])
    with pytest.raises(FeedError) as e:
        read_file(p, True, 1 << 20)
    assert e.value.code == -6 and "Not a gzipped file" in str(e.value)


def test_bgzf_corruption_is_reported(tmp_path):
    data = _data(5, 1 << 20)
    blob = bytearray(bgzf_compress(data))
    p = str(tmp_path / "bad.gz")
    blob[len(blob) // 2] ^= 0x55                                 # flip a bit inside some member's deflate data
    open(p, "wb").write(bytes(blob))
    with pytest.raises(FeedError) as e:
        read_file(p, True, 1 << 20)
    assert e.value.code == -6
    with pytest.raises(Exception):
        gzip.open(p, "rb").read()                             # the reference fails on it as well
    good = bgzf_block(data[:1000])
    open(p, "wb").write(good + good[:-8] + b"\0" * 8)         # second member: wrong CRC and ISIZE
    with pytest.raises(FeedError):
        read_file(p, True, 1 << 20)
    with pytest.raises(FeedError) as e:
        read_file(str(tmp_path / "missing.gz"), True)
    assert e.value.code == -3


# ---------------------------------------------------------------------------------------------
# ordinary gzip streams inflated in parallel (tagdigger_b200/csrc/tdg_pgz.h)

import functools


@functools.lru_cache(maxsize=None)
def _fastq_like(seed, n):
    import numpy as np
    r = np.random.default_rng(seed)
    out, size, i = [], 0, 0
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)
    while size < n:                       # blocks of 2,000 records of one length each
        L = int(r.integers(40, 130))
        m = 2000
        seqs = bases[r.integers(0, 4, (m, L))]
        quals = (35 + np.minimum(39, r.integers(20, 60, (m, L)))).astype(np.uint8)
        xs, ys = r.integers(1000, 30000, m), r.integers(1000, 30000, m)
        for j in range(m):
            rec = b"@M:%d:%d:%d 1:N:0\n%s\n+\n%s\n" % ((i + j) % 8, xs[j], ys[j], seqs[j].tobytes(), quals[j].tobytes())
            out.append(rec)
            size += len(rec)
        i += m
    return b"".join(out)


def _pgz_env(monkeypatch, threads, chunk):
    monkeypatch.setenv("TDG_IO_THREADS", str(threads))
    monkeypatch.setenv("TDG_PGZ_MIN", "0")
    monkeypatch.setenv("TDG_PGZ_CHUNK", str(chunk))


@pytest.mark.parametrize("level", [1, 6, 9])
@pytest.mark.parametrize("threads,pchunk,chunk", [(2, 65536, 1 << 20), (3, 40000, 333333), (8, 150000, 4 << 20),
                                                  (8, 4096, 70000)])
def test_pgzip_fastq(tmp_path, monkeypatch, level, threads, pchunk, chunk):
    _pgz_env(monkeypatch, threads, pchunk)
    data = _fastq_like(10 + level, 6 << 20)
    p = str(tmp_path / "x.fq.gz")
    open(p, "wb").write(gzip.compress(data, level))
    got, mode = read_file(p, True, chunk, cap=len(data) + 100)
    assert mode == "pgzip" and got == data


def _deflate_pieces(pieces):
    """One gzip member whose deflate stream is made of pieces (data, level, strategy, flush)."""
    import struct
    import zlib
    raw = b"".join(d for d, _, _, _ in pieces)
    body = []
    for d, level, strategy, flush in pieces:
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
        # every piece is its own raw deflate stream ended by a flush, so the pieces concatenate
        body.append(co.compress(d) + co.flush(flush))
    tail = zlib.compressobj(6, zlib.DEFLATED, -15)
    body.append(tail.compress(b"") + tail.flush())                     # the final (empty) block
    return (b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\xff" + b"".join(body) +
            struct.pack("<II", zlib.crc32(raw) & 0xFFFFFFFF, len(raw) & 0xFFFFFFFF)), raw


def test_pgzip_block_kinds(tmp_path, monkeypatch):
    """Stored, fixed-Huffman and dynamic blocks, flush points, incompressible and highly
    repetitive stretches: chunks the block finder cannot enter are inflated serially."""
    import zlib
    r = random.Random(7)
    text = _fastq_like(3, 3 << 20)
    noise = r.randbytes(1 << 20)
    pieces = [
        (text[:900000], 6, zlib.Z_DEFAULT_STRATEGY, zlib.Z_FULL_FLUSH),
        (noise, 6, zlib.Z_DEFAULT_STRATEGY, zlib.Z_FULL_FLUSH),          # stored blocks
        (text[900000:1500000], 6, zlib.Z_FIXED, zlib.Z_FULL_FLUSH),       # fixed-Huffman blocks
        (b"A" * (9 << 20), 9, zlib.Z_DEFAULT_STRATEGY, zlib.Z_FULL_FLUSH),   # 1000:1, fills the speculative buffers
        (text[1500000:], 1, zlib.Z_DEFAULT_STRATEGY, zlib.Z_FULL_FLUSH),
        (noise[:70000], 0, zlib.Z_DEFAULT_STRATEGY, zlib.Z_FULL_FLUSH),   # level 0
        (text[:300000], 9, zlib.Z_HUFFMAN_ONLY, zlib.Z_FULL_FLUSH),
        (text[:300000], 9, zlib.Z_RLE, zlib.Z_FULL_FLUSH),
    ]
    blob, raw = _deflate_pieces(pieces)
    p = str(tmp_path / "kinds.gz")
    open(p, "wb").write(blob)
    assert gzip.open(p, "rb").read() == raw
    for threads, pchunk in ((4, 65536), (8, 20000), (3, 1 << 20)):
        _pgz_env(monkeypatch, threads, pchunk)
        got, mode = read_file(p, True, 1 << 20, cap=len(raw) + 100)
        assert mode == "pgzip" and got == raw


def test_pgzip_sync_flush_without_window_reset(tmp_path, monkeypatch):
    """Z_SYNC_FLUSH keeps the window: later pieces refer back across the flush points."""
    import struct
    import zlib
    text = _fastq_like(4, 2 << 20)
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = b""
    for i in range(0, len(text), 50000):
        body += co.compress(text[i:i + 50000]) + co.flush(zlib.Z_SYNC_FLUSH)
    body += co.flush()
    blob = b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\xff" + body + struct.pack("<II", zlib.crc32(text), len(text))
    p = str(tmp_path / "sync.gz")
    open(p, "wb").write(blob)
    _pgz_env(monkeypatch, 8, 30000)
    got, mode = read_file(p, True, 1 << 20, cap=len(text) + 100)
    assert mode == "pgzip" and got == text


def test_pgzip_members_and_headers(tmp_path, monkeypatch):
    import io
    data = _fastq_like(5, 4 << 20)
    cuts = [0, 10, 10, 70000, 70001, 900000, 2500000, len(data)]
    blob = b""
    for a, b in zip(cuts, cuts[1:]):
        bio = io.BytesIO()
        with gzip.GzipFile(filename="part%d.fq" % a, mode="wb", fileobj=bio, compresslevel=1 + a % 9, mtime=a) as g:
            g.write(data[a:b])
        blob += bio.getvalue()
    # a member with FEXTRA and FCOMMENT
    import struct
    import zlib
    extra = b"XY\x03\x00abc"
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = co.compress(data[:5000]) + co.flush()
    blob += (b"\x1f\x8b\x08\x14\x00\x00\x00\x00\x00\xff" + struct.pack("<H", len(extra)) + extra + b"a comment\x00" + body +
             struct.pack("<II", zlib.crc32(data[:5000]), 5000))
    p = str(tmp_path / "members.gz")
    open(p, "wb").write(blob)
    want = data + data[:5000]
    assert gzip.open(p, "rb").read() == want
    for threads, pchunk in ((8, 50000), (2, 4096), (5, 1 << 20)):
        _pgz_env(monkeypatch, threads, pchunk)
        got, mode = read_file(p, True, 777777, cap=len(want) + 100)
        assert mode == "pgzip" and got == want
    # trailing zeros and trailing garbage: ignored, as by the zlib path
    for tail in (b"\0" * 1000, b"garbage that is not gzip", b"\x1f"):
        open(p, "wb").write(blob + tail)
        _pgz_env(monkeypatch, 1, 50000)
        ref = read_file(p, True, 1 << 20, cap=len(want) + 100)
        _pgz_env(monkeypatch, 8, 50000)
        got = read_file(p, True, 1 << 20, cap=len(want) + 100)
        assert ref == (want, "zlib") and got == (want, "pgzip")


def _outcome(path, cap):
    try:
        return read_file(path, True, 1 << 20, cap=cap)[0]
    except FeedError as e:
        return e.code


def test_pgzip_anomalies_end_like_the_zlib_path(tmp_path, monkeypatch):
    """Header CRC flag, corrupt deflate data, wrong CRC / ISIZE, truncation, a bad second header:
    same bytes or same error code as zlib's gzread on one thread."""
    import struct
    import zlib
    data = _fastq_like(6, 3 << 20)
    good = gzip.compress(data, 6)
    cases = {}
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = co.compress(data) + co.flush()
    head = b"\x1f\x8b\x08\x02\x00\x00\x00\x00\x00\xff"
    cases["fhcrc"] = head + struct.pack("<H", zlib.crc32(head) & 0xFFFF) + body + struct.pack("<II", zlib.crc32(data), len(data))
    bad = bytearray(good)
    bad[len(bad) // 2] ^= 0x10
    cases["flipped bit"] = bytes(bad)
    cases["wrong crc"] = good[:-8] + struct.pack("<II", zlib.crc32(data) ^ 1, len(data))
    cases["wrong isize"] = good[:-8] + struct.pack("<II", zlib.crc32(data), len(data) + 1)
    cases["truncated"] = good[:len(good) * 2 // 3]
    cases["truncated trailer"] = good[:-3]
    cases["bad second header"] = good + b"\x1f\x8b\x09\x00" + good[4:]
    cases["second member corrupt"] = good + bytes(bad)
    cases["reserved flag"] = good + b"\x1f\x8b\x08\x80" + good[4:]
    p = str(tmp_path / "odd.gz")
    for name, blob in cases.items():
        open(p, "wb").write(blob)
        _pgz_env(monkeypatch, 1, 40000)
        ref = _outcome(p, 2 * len(data) + 100)
        for threads, pchunk in ((8, 40000), (3, 1 << 20)):
            _pgz_env(monkeypatch, threads, pchunk)
            got = _outcome(p, 2 * len(data) + 100)
            assert got == ref, name
        if name not in ("fhcrc", "reserved flag"):
            assert isinstance(ref, int) and ref == -6, name


def test_pgzip_small_files_stay_on_zlib(tmp_path, monkeypatch):
    monkeypatch.setenv("TDG_IO_THREADS", "8")
    monkeypatch.delenv("TDG_PGZ_MIN", raising=False)
    data = _fastq_like(8, 1 << 20)
    p = str(tmp_path / "s.gz")
    open(p, "wb").write(gzip.compress(data))
    assert read_file(p, True, 1 << 20) == (data, "zlib")
