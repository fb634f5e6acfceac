"""ctypes wrapper over oracle/oracle.c.  TEST INFRASTRUCTURE ONLY (see the
header of oracle/tagdigger_oracle.py)."""

import ctypes
import os
import subprocess

import numpy as np

from . import tagdigger_oracle as py

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _SO, src])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.orc_trie_build.restype = ctypes.c_int
        L.orc_trie_build.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32,
                                     ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64)]
        L.orc_trie_free.argtypes = [ctypes.c_void_p]
        L.orc_trie_lookup.restype = ctypes.c_int32
        L.orc_trie_lookup.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
        L.orc_count.restype = ctypes.c_int
        L.orc_count.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
                                ctypes.c_void_p, ctypes.c_uint32, ctypes.c_double, ctypes.c_uint64,
                                ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_count_lines.restype = ctypes.c_uint64
        L.orc_count_lines.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
        _lib = L
    return _lib


class CTrie(object):
    def __init__(self, sequences, numseq):
        L = lib()
        blob = "".join(sequences).encode()
        off = np.zeros(len(sequences) + 1, dtype=np.uint64)
        np.cumsum([len(s) for s in sequences], out=off[1:])
        self._h = ctypes.c_void_p()
        prob = ctypes.c_int64(0)
        rc = L.orc_trie_build(blob, off.ctypes.data, len(sequences), numseq,
                              ctypes.byref(self._h), ctypes.byref(prob))
        if rc == 1:
            raise AssertionError("Problematic sequence: {}.  Likely due to overlapping tags."
                                 .format(prob.value))
        if rc == 2:
            raise IndexError("list index out of range")

    def lookup(self, seq):
        r = lib().orc_trie_lookup(self._h, seq.encode(), len(seq))
        if r == -2:
            raise IndexError("index out of range")
        if r == -3:
            raise TypeError("'int' object is not subscriptable")
        return r

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_trie_free(self._h)
            self._h = None


def _setup(barcodes, tags, cutsite):
    """Same preparation as tagdigger_oracle.prepare (tagdigger_fun.py:198-233),
    but building the C tries."""
    if not all(set(b.upper()) <= set("ACGT") for b in barcodes):
        raise AssertionError("Non-ACGT barcode.")
    cutsite = cutsite.upper()
    if not set(cutsite) <= set("ACGTNRYKMSWBDHV"):
        raise AssertionError("Invalid cut site.")
    tags = [t.upper() for t in tags]
    if not all(set(t) <= set("ACGT") for t in tags):
        raise AssertionError("Non-ACGT tag.")
    cutlen = len(cutsite)
    offsets = [len(b) + cutlen for b in barcodes]
    sites = py.expand_cut_site(cutsite)
    pats = []
    for s in sites:
        pats.extend(py.barcode_patterns(barcodes, s))
    bar = CTrie(pats, len(barcodes))
    if set(t[:cutlen] for t in tags) <= set(sites):
        if len(sites) == 1:
            tags = [t[cutlen:] for t in tags]
        else:
            offsets = [o - cutlen for o in offsets]
    tag = CTrie(tags, len(tags))
    return bar, tag, np.asarray(offsets, dtype=np.uint32), len(barcodes), len(tags)


class Counter(object):
    """Reusable (barcodes, tags, cutsite) set-up for counting many images."""

    def __init__(self, barcodes, tags, cutsite="TGCAG"):
        self.bar, self.tag, self.offsets, self.barnum, self.ntags = _setup(barcodes, tags, cutsite)

    def count(self, data, maxreads=5e9, first_line=0, reads_before=0):
        """data: bytes or uint8 ndarray.  Returns (int64 matrix, [reads, barcut, tag])."""
        arr = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        counts = np.zeros((self.barnum, self.ntags), dtype=np.int64)
        totals = np.zeros(3, dtype=np.uint64)
        rc = lib().orc_count(arr.ctypes.data, arr.size, self.bar._h, self.tag._h,
                             self.offsets.ctypes.data, self.ntags, float(maxreads),
                             first_line, reads_before, counts.ctypes.data, totals.ctypes.data)
        if rc == -2:
            raise IndexError("index out of range")
        if rc == -3:
            raise TypeError("'int' object is not subscriptable")
        return counts, [int(x) for x in totals]


def find_tags_bytes(data, barcodes, tags, cutsite="TGCAG", maxreads=5e9):
    c, _ = Counter(barcodes, tags, cutsite).count(data, maxreads)
    return c.tolist()


def count_lines(data):
    arr = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    return int(lib().orc_count_lines(arr.ctypes.data, arr.size))


# ---- sharded counting of large images (parity checks at 10^7 reads and more) ----------------------

_SHARD = {}


def _lines_worker(span):
    a, b = span
    return count_lines(_SHARD["img"][a:b])


def _count_worker(job):
    a, b, first_line = job
    # sequence lines (index % 4 == 1) that precede line `first_line`
    m, tot = _SHARD["counter"].count(_SHARD["img"][a:b], first_line=first_line, reads_before=(first_line + 2) // 4)
    nz = np.flatnonzero(m)                      # sparse: a 384 x 500,000 matrix is 1.5 GB, its non-zeros are few
    return m.shape, nz, m.ravel()[nz], tot


def count_sharded(img, counter, procs=None):
    """Exact counts of a FASTQ image (uint8 ndarray whose lines end in '\\n') with ``counter``,
    sharded over forked host processes: the image is cut after line feeds, the lines before every
    cut are counted first (orc_count_lines), and every shard is then counted with its true first
    line index -- the same result as one sequential pass.  Returns (int64 matrix, totals)."""
    import multiprocessing as mp
    import os
    procs = procs or os.cpu_count() or 1
    n = img.size
    cuts = [0]
    for i in range(1, procs):
        pos = max(cuts[-1], n * i // procs)
        nl = np.flatnonzero(img[pos:pos + (1 << 20)] == 10)
        cuts.append(pos + int(nl[0]) + 1 if nl.size else n)
    cuts.append(n)
    spans = [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    if len(spans) <= 1:
        return counter.count(img)
    lib()
    _SHARD["img"] = img
    _SHARD["counter"] = counter
    try:
        with mp.get_context("fork").Pool(len(spans)) as pool:
            nlines = pool.map(_lines_worker, spans, chunksize=1)
            firsts = np.concatenate([[0], np.cumsum(nlines)[:-1]]).tolist()
            res = pool.map(_count_worker, [(a, b, int(f)) for (a, b), f in zip(spans, firsts)], chunksize=1)
    finally:
        _SHARD.clear()
    total = np.zeros(res[0][0], dtype=np.int64)
    tot = [0, 0, 0]
    for _shape, nz, vals, t in res:
        total.ravel()[nz] += vals
        tot = [x + y for x, y in zip(tot, t)]
    return total, tot
