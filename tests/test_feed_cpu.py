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
