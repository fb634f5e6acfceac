"""ctypes wrapper of the test-side table harness (tests/native/table_check.cpp)."""

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "native", "table_check.cpp")
_LIB = os.path.join(_HERE, "native", "libtable_check.so")
_CSRC = os.path.join(os.path.dirname(_HERE), "tagdigger_b200", "csrc")


def build(force=False):
    deps = [_SRC, os.path.join(_CSRC, "tdg_tables.h"), os.path.join(_CSRC, "tdg_match.h")]
    if not force and os.path.exists(_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(_LIB) for d in deps):
        return _LIB
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", _LIB, _SRC])
    return _LIB


class TableError(RuntimeError):
    pass


class Tables(object):
    """Packed barcode + tag tables on the host, with the kernel's own match code."""

    def __init__(self):
        L = ctypes.CDLL(build())
        vp = ctypes.c_void_p
        L.tck_create.restype = vp
        L.tck_destroy.argtypes = [vp]
        L.tck_error.restype = ctypes.c_char_p
        L.tck_error.argtypes = [vp]
        L.tck_set_tags.argtypes = [vp, vp, vp, vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
        L.tck_set_bars.argtypes = [vp, vp, vp, vp, vp, ctypes.c_uint32, ctypes.c_uint32]
        L.tck_match.restype = ctypes.c_longlong
        L.tck_match.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t]
        L.tck_table_stats.argtypes = [vp, vp]
        self._L = L
        self._h = L.tck_create()
        self.cols = 0

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.tck_destroy(self._h)
            self._h = None

    def _ck(self, rc):
        if rc:
            raise TableError(self._L.tck_error(self._h).decode())

    @staticmethod
    def _csr(seqs, dtype):
        blob = "".join(seqs).encode("ascii")
        off = np.zeros(len(seqs) + 1, dtype=dtype)
        if len(seqs):
            np.cumsum([len(s) for s in seqs], out=off[1:])
        return blob, off

    def set_tags(self, seqs, cols=None, any_base=False, ncols=None):
        cols = list(range(len(seqs))) if cols is None else list(cols)
        blob, off = self._csr(seqs, np.uint64)
        col = np.asarray(cols, dtype=np.int32)
        self.cols = ncols if ncols is not None else (max(cols) + 1 if cols else 0)
        self._ck(self._L.tck_set_tags(self._h, blob, off.ctypes.data, col.ctypes.data, len(seqs), 1 if any_base else 0,
                                      self.cols))

    def set_bars(self, patterns, rows, tag_offs, any_base=False):
        blob, off = self._csr(patterns, np.uint32)
        row = np.asarray(list(rows), dtype=np.int32)
        toff = np.asarray(list(tag_offs), dtype=np.uint32)
        self._ck(self._L.tck_set_bars(self._h, blob, off.ctypes.data, row.ctypes.data, toff.ctypes.data, len(patterns),
                                      1 if any_base else 0))

    def load_plan(self, p):
        self.set_tags(p.tags.patterns, p.tags.index, any_base=p.tags.any_base, ncols=p.ntags)
        self.set_bars(p.bar.patterns, p.bar.index, p.bar_tag_off, any_base=p.bar.any_base)

    def match(self, read):
        if isinstance(read, str):
            read = read.encode("utf-8")
        return int(self._L.tck_match(self._h, read, len(read)))

    def stats(self):
        out = np.zeros(3, dtype=np.uint64)
        self._L.tck_table_stats(self._h, out.ctypes.data)
        return {"slots": int(out[0]), "used": int(out[1]), "more": int(out[2])}
