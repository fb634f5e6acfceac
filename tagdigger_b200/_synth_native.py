"""ctypes binding of the device-side synthetic FASTQ generator
(csrc/tdg_synth.cu -> libtdg_synth.so).  Bench / test support only: it makes
the 200 M-read images of BASELINE.json directly in HBM, deterministically and
shard by shard; it is not part of the counting path."""

import ctypes
import os
import subprocess

import numpy as np

from . import _native

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "csrc", "tdg_synth.cu")
LIB_PATH = os.path.join(_HERE, "libtdg_synth.so")
_lib = None


def build(force=False):
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= os.path.getmtime(_SRC):
        return LIB_PATH
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB_PATH, _SRC]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + proc.stdout)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        try:
            build()
        except (OSError, RuntimeError):
            if not os.path.exists(LIB_PATH):
                raise
        L = ctypes.CDLL(LIB_PATH)
        vp, u64, u32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32
        L.tdgs_generate.restype = ctypes.c_int
        L.tdgs_generate.argtypes = [ctypes.c_int, u64, u64, u64, u32, ctypes.c_char_p,
                                    vp, vp, u32, u32, vp, vp, vp, u32, u32, vp,
                                    ctypes.c_size_t, ctypes.c_size_t, ctypes.POINTER(vp), ctypes.POINTER(u64), vp]
        L.tdgs_free.argtypes = [ctypes.c_int, vp]
        L.tdgs_last_error.restype = ctypes.c_char_p
        _lib = L
    return _lib


def _matrix(seqs):
    lens = np.array([len(s) for s in seqs], dtype=np.uint32)
    stride = max(16, int(lens.max()) if len(seqs) else 16)
    m = np.full((len(seqs), stride), ord("A"), dtype=np.uint8)
    for i, s in enumerate(seqs):
        m[i, :len(s)] = np.frombuffer(s.encode(), dtype=np.uint8)
    return m, lens, stride


def _thresholds(p_hit, p_unknown, p_nobar, p_n, p_lower, p_short, p_qual_at):
    full = float(1 << 32)

    def t(x):
        return min(int(x * full), (1 << 32) - 1)
    return np.array([t(p_hit), t(p_hit + p_unknown), t(p_hit + p_unknown + p_nobar),
                     t(p_n), t(p_lower), t(p_short), t(p_qual_at), 0], dtype=np.uint32)


class Generator(object):
    """Synthetic read mix of SURVEY.md 8(d) for a barcode/tag set.  ``tags`` are
    full tag sequences INCLUDING the cut site (as they appear in the read)."""

    def __init__(self, barcodes, tags, cutsite="TGCAG", readlen=100, seed=20162, zipf=1.0,
                 p_hit=0.60, p_unknown=0.20, p_nobar=0.15, p_n=0.02, p_lower=0.001, p_short=0.001,
                 p_qual_at=0.01):
        self.bar, self.bar_len, self.bar_stride = _matrix(barcodes)
        self.tag, self.tag_len, self.tag_stride = _matrix(tags)
        w = 1.0 / np.arange(1, len(tags) + 1, dtype=np.float64) ** zipf
        cdf = np.cumsum(w / w.sum())
        self.cdf = np.minimum(cdf * float(1 << 32), float((1 << 32) - 1)).astype(np.uint32)
        self.cdf[-1] = (1 << 32) - 1
        self.probs = _thresholds(p_hit, p_unknown, p_nobar, p_n, p_lower, p_short, p_qual_at)
        self.cutsite = cutsite.encode()
        self.readlen = readlen
        self.seed = seed
        self.nbar, self.ntags = len(barcodes), len(tags)

    def generate(self, device, first_read, nreads, expected_ptr=None):
        """Image of reads [first_read, first_read + nreads) in device memory,
        padded for tdg_count_device.  Returns (device pointer, nbytes)."""
        L = lib()
        out = ctypes.c_void_p()
        nb = ctypes.c_uint64()
        rc = L.tdgs_generate(device, self.seed, first_read, nreads, self.readlen, self.cutsite,
                             self.bar.ctypes.data, self.bar_len.ctypes.data, self.nbar, self.bar_stride,
                             self.tag.ctypes.data, self.tag_len.ctypes.data, self.cdf.ctypes.data, self.ntags,
                             self.tag_stride, self.probs.ctypes.data,
                             _native.TDG_TILE_BYTES, _native.TDG_HALO_BYTES,
                             ctypes.byref(out), ctypes.byref(nb), expected_ptr)
        if rc != 0:
            raise RuntimeError("tdgs_generate: " + L.tdgs_last_error().decode())
        return out.value, nb.value

    def free(self, device, ptr):
        lib().tdgs_free(device, ptr)
