"""The per-lane inflater of the device-side gzip feed (tagdigger_b200/csrc/tdg_gzlane.h) and its
round / chain logic (tdg_gzchain.h), run on the CPU one lane after the other through
tests/native/gzlane_check.cpp: the bytes must be those Python's gzip module reads
(gzip.open of /root/reference/tagdigger_fun.py:240-241), whatever the block kinds, the chunk size
and the place where the host reader has to take over."""

import gzip
import io
import random
import struct
import zlib

import pytest

from gzlane_check import inflate
from test_feed_cpu import _deflate_pieces, _fastq_like


@pytest.mark.parametrize("level", [1, 6, 9])
@pytest.mark.parametrize("chunk,max_chunks,search", [(1 << 16, 32, 1 << 15), (40000 // 16 * 16, 7, 1 << 15), (1 << 17, 64, 1 << 17),
                                                     (4096, 300, 4096)])
def test_lane_inflater_fastq(level, chunk, max_chunks, search):
    data = _fastq_like(10 + level, 6 << 20)
    out, n, info = inflate(gzip.compress(data, level), chunk=chunk, max_chunks=max_chunks, search_bytes=search)
    assert n == len(data) and out == data
    if chunk >= 1 << 16:
        # the lanes did the work: nearly every chunk continued the stream, nothing went back to the host
        assert info["handover"] == 0 and info["accepted"] >= 0.9 * info["chunks"]


def test_lane_inflater_block_kinds():
    """Stored, fixed-Huffman and dynamic blocks, flush points, incompressible and 1000:1 stretches."""
    r = random.Random(7)
    text = _fastq_like(3, 3 << 20)
    noise = r.randbytes(1 << 20)
    pieces = [
        (text[:900000], 6, zlib.Z_DEFAULT_STRATEGY, zlib.Z_FULL_FLUSH),
        (noise, 6, zlib.Z_DEFAULT_STRATEGY, zlib.Z_FULL_FLUSH),
        (text[900000:1500000], 6, zlib.Z_FIXED, zlib.Z_FULL_FLUSH),
        (b"A" * (9 << 20), 9, zlib.Z_DEFAULT_STRATEGY, zlib.Z_FULL_FLUSH),
        (text[1500000:], 1, zlib.Z_DEFAULT_STRATEGY, zlib.Z_FULL_FLUSH),
        (noise[:70000], 0, zlib.Z_DEFAULT_STRATEGY, zlib.Z_FULL_FLUSH),
        (text[:300000], 9, zlib.Z_HUFFMAN_ONLY, zlib.Z_FULL_FLUSH),
        (text[:300000], 9, zlib.Z_RLE, zlib.Z_FULL_FLUSH),
    ]
    blob, raw = _deflate_pieces(pieces)
    assert gzip.decompress(blob) == raw
    for chunk, max_chunks in ((1 << 16, 16), (20000 // 16 * 16, 64), (1 << 20, 4)):
        out, n, info = inflate(blob, chunk=chunk, max_chunks=max_chunks, cap=len(raw) + 100)
        assert n == len(raw) and out == raw, (chunk, info)


def test_lane_inflater_each_kind_alone():
    """Every block kind as a whole stream (the first chunk of a round starts in it)."""
    text = _fastq_like(4, 1 << 20)
    for strategy, level in ((zlib.Z_FIXED, 6), (zlib.Z_HUFFMAN_ONLY, 6), (zlib.Z_RLE, 6), (zlib.Z_DEFAULT_STRATEGY, 0),
                            (zlib.Z_FILTERED, 9)):
        co = zlib.compressobj(level, zlib.DEFLATED, 31, 9, strategy)
        blob = co.compress(text) + co.flush()
        out, n, info = inflate(blob, chunk=1 << 15, max_chunks=16, cap=len(text) + 100)
        assert n == len(text) and out == text, (strategy, level, info)


def test_lane_inflater_long_codes():
    """Skewed symbol statistics give codes longer than the 10-bit / 8-bit primary tables."""
    r = random.Random(11)
    # geometric byte distribution: a few very frequent symbols and a long tail of rare ones
    vals = []
    for _ in range(3 << 20):
        v = 0
        while v < 255 and r.random() < 0.72:
            v += 1
        vals.append(v)
    data = bytes(vals)
    # plus rare long matches at many distances (long distance codes)
    parts = [data[:1 << 20]]
    for i in range(2000):
        d = r.randint(1, 32000)
        parts.append(data[(1 << 20) + i * 500:(1 << 20) + i * 500 + 300])
        parts.append(parts[-1][:r.randint(3, 200)])
    data = b"".join(parts)
    for level in (1, 9):
        out, n, info = inflate(gzip.compress(data, level), chunk=1 << 15, max_chunks=24, cap=len(data) + 100)
        assert n == len(data) and out == data, info


def test_lane_inflater_sync_flush_keeps_the_window():
    text = _fastq_like(4, 2 << 20)
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = b""
    for i in range(0, len(text), 50000):
        body += co.compress(text[i:i + 50000]) + co.flush(zlib.Z_SYNC_FLUSH)
    body += co.flush()
    blob = b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\xff" + body + struct.pack("<II", zlib.crc32(text), len(text))
    out, n, info = inflate(blob, chunk=30000 // 16 * 16, max_chunks=20, cap=len(text) + 100)
    assert n == len(text) and out == text


def test_lane_inflater_members_and_headers():
    data = _fastq_like(5, 4 << 20)
    cuts = [0, 10, 10, 70000, 70001, 900000, 2500000, len(data)]
    blob = b""
    for a, b in zip(cuts, cuts[1:]):
        bio = io.BytesIO()
        with gzip.GzipFile(filename="part%d.fq" % a, mode="wb", fileobj=bio, compresslevel=1 + a % 9, mtime=a) as g:
            g.write(data[a:b])
        blob += bio.getvalue()
    extra = b"XY\x03\x00abc"
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = co.compress(data[:5000]) + co.flush()
    blob += (b"\x1f\x8b\x08\x14\x00\x00\x00\x00\x00\xff" + struct.pack("<H", len(extra)) + extra + b"a comment\x00" + body +
             struct.pack("<II", zlib.crc32(data[:5000]), 5000))
    want = data + data[:5000]
    assert gzip.decompress(blob) == want
    for chunk, max_chunks in ((50000 // 16 * 16, 16), (4096, 64), (1 << 20, 8)):
        out, n, info = inflate(blob, chunk=chunk, max_chunks=max_chunks, cap=len(want) + 100)
        assert n == len(want) and out == want, info
    for tail in (b"\0" * 1000, b"garbage that is not gzip", b"\x1f"):
        out, n, info = inflate(blob + tail, chunk=1 << 16, max_chunks=16, cap=len(want) + 100)
        assert n == len(want) and out == want, (tail[:8], info)


def test_lane_inflater_hands_damaged_streams_to_the_host_reader():
    """Corrupt deflate data, wrong CRC / ISIZE, truncation, a bad second header: the lanes deliver a
    correct prefix and the stream ends on the host reader (-1: zlib judges it) or with a bad check
    (-2), never with wrong bytes."""
    data = _fastq_like(6, 3 << 20)
    good = gzip.compress(data, 6)
    bad = bytearray(good)
    bad[len(bad) // 2] ^= 0x10
    cases = {
        "flipped bit": (bytes(bad), (-1, -2)),
        "wrong crc": (good[:-8] + struct.pack("<II", zlib.crc32(data) ^ 1, len(data)), (-2,)),
        "wrong isize": (good[:-8] + struct.pack("<II", zlib.crc32(data), len(data) + 1), (-2,)),
        "truncated": (good[:len(good) * 2 // 3], (-1,)),
        "truncated trailer": (good[:-3], (-1,)),
        "bad second header": (good + b"\x1f\x8b\x09\x00" + good[4:], (-1,)),
        "second member corrupt": (good + bytes(bad), (-1, -2)),
        "reserved flag": (good + b"\x1f\x8b\x08\x80" + good[4:], (-1,)),
    }
    for name, (blob, codes) in cases.items():
        for chunk, max_chunks in ((40000 // 16 * 16, 12), (1 << 18, 4)):
            out, n, info = inflate(blob, chunk=chunk, max_chunks=max_chunks, cap=2 * len(data) + 100)
            assert n in codes, (name, n, info)
            # what was delivered before the defect is a prefix of the true text (twice, for two members)
            assert (data + data).startswith(out), (name, info)
    # a header with the FHCRC flag is not this code's to parse
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = co.compress(data) + co.flush()
    head = b"\x1f\x8b\x08\x02\x00\x00\x00\x00\x00\xff"
    blob = head + struct.pack("<H", zlib.crc32(head) & 0xFFFF) + body + struct.pack("<II", zlib.crc32(data), len(data))
    assert inflate(blob)[1] == -10


def test_lane_inflater_small_symbol_buffers_hand_over():
    """A chunk whose output does not fit its symbol buffer ends the device feed; the host reader finishes."""
    data = b"ACGT" * (3 << 20) + _fastq_like(2, 1 << 20)
    blob = gzip.compress(data, 6)
    out, n, info = inflate(blob, chunk=1 << 14, max_chunks=8, symcap=8 << 14, cap=len(data) + 100)
    assert n == len(data) and out == data and info["handover"] == 1


def test_lane_inflater_random_streams():
    """Random mixtures of literal runs, matches and noise at random chunk geometries."""
    r = random.Random(2024)
    for case in range(12):
        parts = []
        for _ in range(r.randint(3, 30)):
            kind = r.random()
            if kind < 0.3:
                parts.append(r.randbytes(r.randint(1, 40000)))
            elif kind < 0.6:
                parts.append(bytes(r.choice(b"ACGTN\n@+I") for _ in range(r.randint(1, 60000))))
            elif kind < 0.8 and parts:
                src = parts[r.randrange(len(parts))]
                parts.append(src[:r.randint(0, len(src))] * r.randint(1, 4))
            else:
                parts.append(bytes([r.randrange(256)]) * r.randint(1, 100000))
        data = b"".join(parts)
        blob = gzip.compress(data, r.choice((1, 4, 6, 9)))
        chunk = r.choice((4096, 8192, 1 << 15, 1 << 16))
        out, n, info = inflate(blob, chunk=chunk, max_chunks=r.randint(1, 40), search_bytes=r.choice((2048, 1 << 15)),
                               cap=len(data) + 100)
        assert n == len(data) and out == data, (case, info)


def test_chunks_without_a_confirmed_start_are_inflated_by_the_host():
    """A lane that finds no block start (or a false one) does not end the round: the host inflates that
    chunk into the same symbols and the chain goes on."""
    data = _fastq_like(21, 8 << 20)
    for level in (1, 6):
        blob = gzip.compress(data, level)
        ref = inflate(blob, chunk=1 << 16, max_chunks=200)
        out, n, info = inflate(blob, chunk=1 << 16, max_chunks=200, blind_every=7)
        assert n == len(data) and out == data
        assert info["repairs"] >= 5 and info["rounds"] == ref[2]["rounds"] and info["handover"] == 0, info


def test_lane_inflater_damage_fuzz():
    """Random damage (bit flips, byte runs overwritten, truncation, bytes inserted) at random places of
    random streams: the feed either delivers exactly what Python's gzip reads, or it stops -- with a
    correct prefix -- on a file Python's gzip refuses too.  Never wrong bytes."""
    r = random.Random(77)
    base = _fastq_like(31, 2 << 20)
    for case in range(40):
        n = r.randint(200000, len(base))
        data = base[:n]
        level = r.choice((1, 6, 9))
        if r.random() < 0.3:
            cut = r.randint(1, n - 1)
            blob = bytearray(gzip.compress(data[:cut], level) + gzip.compress(data[cut:], level))
        else:
            blob = bytearray(gzip.compress(data, level))
        kind = r.choice(("flip", "run", "truncate", "insert", "none"))
        if kind == "flip":
            for _ in range(r.randint(1, 3)):
                blob[r.randrange(len(blob))] ^= 1 << r.randrange(8)
        elif kind == "run":
            at = r.randrange(len(blob))
            blob[at:at + r.randint(1, 300)] = r.randbytes(r.randint(1, 300))
        elif kind == "truncate":
            del blob[r.randint(len(blob) // 2, len(blob) - 1):]
        elif kind == "insert":
            at = r.randrange(len(blob))
            blob[at:at] = r.randbytes(r.randint(1, 50))
        blob = bytes(blob)
        try:
            want = gzip.decompress(blob)
        except Exception:  # noqa: BLE001
            want = None
        # what a sequential inflater hands out before it notices the damage (garbage included: the
        # reference's text iteration would see those bytes, too, before gzip raises)
        seq = b""
        try:
            rest = blob
            while rest:
                d = zlib.decompressobj(31)
                seq += d.decompress(rest)
                if not d.eof:
                    break
                rest = d.unused_data
                if rest[:2] != b"\x1f\x8b":
                    break
        except zlib.error:
            pass
        chunk = r.choice((8192, 1 << 15, 1 << 16))
        out, code, info = inflate(blob, chunk=chunk, max_chunks=r.randint(2, 40), cap=2 * len(base) + 100)
        if code >= 0:
            assert want is None or out == want, (case, kind, info)
            assert want is not None or seq.startswith(out) or out.startswith(seq), (case, kind, info)
        else:
            assert code in (-1, -2, -10), (case, kind, code, info)
            if code != -10:
                truth = want if want is not None else seq
                assert truth.startswith(out) or (want is None and out.startswith(seq)), (case, kind, info)
