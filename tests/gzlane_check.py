"""ctypes wrapper of the CPU harness over the device feed's per-lane inflater
(tests/native/gzlane_check.cpp; product code: tagdigger_b200/csrc/tdg_gzlane.h, tdg_gzchain.h)."""

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "native", "gzlane_check.cpp")
_LIB = os.path.join(_HERE, "native", "libgzlane_check.so")
_DEPS = [os.path.join(os.path.dirname(_HERE), "tagdigger_b200", "csrc", h) for h in ("tdg_gzlane.h", "tdg_gzchain.h", "tdg_pgz.h")]


def build(force=False):
    if not force and os.path.exists(_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(_LIB) for d in [_SRC] + _DEPS):
        return _LIB
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", _LIB, _SRC, "-lz"])
    return _LIB


def inflate(blob, chunk=1 << 16, max_chunks=64, search_bytes=1 << 15, symcap=None, cap=None, blind_every=0):
    """(bytes or None, code, info dict).  code >= 0: everything was inflated."""
    L = ctypes.CDLL(build())
    L.gzl_inflate.restype = ctypes.c_longlong
    L.gzl_inflate.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_size_t, ctypes.c_uint32,
                              ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint32]
    symcap = symcap or 8 * chunk
    cap = cap or max(1 << 20, 40 * len(blob) + (1 << 20))
    out = np.empty(cap, dtype=np.uint8)
    info = np.zeros(6, dtype=np.int64)
    why = ctypes.create_string_buffer(128)
    n = L.gzl_inflate(blob, len(blob), chunk, max_chunks, search_bytes, symcap, out.ctypes.data, cap, info.ctypes.data, why, 128, blind_every)
    d = {"rounds": int(info[0]), "accepted": int(info[1]), "handover": int(info[2]), "repairs": int(info[3]),
         "delivered": int(info[4]), "chunks": int(info[5]), "why": why.value.decode()}
    if n >= 0:
        return out[:n].tobytes(), int(n), d
    return out[:int(info[4])].tobytes(), int(n), d
