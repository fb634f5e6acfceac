"""ctypes wrapper of the test-side feeder harness (tests/native/feed_check.cpp) and a BGZF writer."""

import ctypes
import os
import struct
import subprocess
import zlib

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "native", "feed_check.cpp")
_LIB = os.path.join(_HERE, "native", "libfeed_check.so")
_DEPS = [os.path.join(os.path.dirname(_HERE), "tagdigger_b200", "csrc", h) for h in ("tdg_feed.h", "tdg_pgz.h", "tdg_text.h", "tdg_pool.h")]


def build(force=False):
    if not force and os.path.exists(_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(_LIB) for d in [_SRC] + _DEPS):
        return _LIB
    subprocess.check_call(["g++", "-O3", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", _LIB, _SRC, "-lz"])
    return _LIB


class FeedError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "%d: %s" % (code, msg))
        self.code = code


MODES = {0: "plain", 1: "plain_seq", 2: "bgzf", 3: "zlib", 4: "pgzip"}


def read_file(path, gz, chunk=1 << 20, cap=None):
    """(bytes delivered by the product's Feeder, final mode name)."""
    L = ctypes.CDLL(build())
    L.fck_read.restype = ctypes.c_longlong
    L.fck_read.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                           ctypes.POINTER(ctypes.c_int)]
    L.fck_error.restype = ctypes.c_char_p
    size = os.path.getsize(path) if os.path.exists(path) else 0
    cap = cap or max(1 << 20, 40 * size + (1 << 20))
    out = np.empty(cap, dtype=np.uint8)
    mode = ctypes.c_int(-1)
    n = L.fck_read(os.fsencode(path), 1 if gz else 0, chunk, out.ctypes.data, cap, ctypes.byref(mode))
    if n < 0:
        raise FeedError(int(n), L.fck_error().decode())
    return out[:n].tobytes(), MODES.get(mode.value, "?")


def bgzf_block(data, level=6):
    """One BGZF member (data <= 65280 bytes)."""
    assert len(data) <= 65280
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = co.compress(data) + co.flush()
    bsize = 12 + 6 + len(body) + 8 - 1
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + body +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def bgzf_compress(data, block=65280, eof_marker=True, level=6):
    out = [bgzf_block(data[i:i + block], level) for i in range(0, len(data), block)]
    if eof_marker:
        out.append(bgzf_block(b""))
    return b"".join(out)


def utf8_first_invalid(data, piece=1 << 16):
    L = ctypes.CDLL(build())
    L.fck_utf8.restype = ctypes.c_longlong
    L.fck_utf8.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t]
    return int(L.fck_utf8(data, len(data), piece))


def line_limit(data, lines, piece=1 << 16):
    L = ctypes.CDLL(build())
    L.fck_line_limit.restype = ctypes.c_longlong
    L.fck_line_limit.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint64]
    return int(L.fck_line_limit(data, len(data), piece, lines))
