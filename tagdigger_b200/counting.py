"""Host side of the counting path: the drop-in for ``find_tags_fastq``.

Mirrors /root/reference/tagdigger_fun.py:192-277 (same signature, defaults,
asserts and return shape); the per-read loop runs on the GPU through
``_native.Engine`` (ctypes over include/tagdigger_b200.h).  Nothing here counts
reads on the CPU.
"""

import gzip
import os

import numpy as np

from . import _native
from . import matchset

_engines = {}


def get_engine(device=None):
    """The process-wide Engine of a CUDA device (created on first use).  With
    ``device=None`` the device is LOCAL_RANK (torchrun) or 0."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    eng = _engines.get(device)
    if eng is None:
        eng = _native.Engine(device=device)
        _engines[device] = eng
    return eng


def load_plan(eng, p, row_of=None, nrows=None, set_tags=True):
    """Upload a matchset.CountPlan.  ``row_of[b]`` maps the barcode index the
    reference's trie would return to a matrix row (identity by default)."""
    if set_tags:
        eng.set_tags(p.tags.patterns, p.tags.index, any_base=p.tags.any_base)
    if nrows is not None:
        eng.set_matrix(nrows, p.ntags)
    rows = p.bar.index if row_of is None else [row_of[b] for b in p.bar.index]
    eng.begin_file(p.bar.patterns, rows, p.bar_tag_off, any_base=p.bar.any_base)


def _is_gz(name):
    return name[-2:].lower() == "gz"            # tagdigger_fun.py:240


def _run_file(eng, fqfile, limit):
    """Stream one file; map library errors to the exceptions the reference's
    open()/gzip.open() would raise."""
    with open(fqfile, "rb"):                    # FileNotFoundError / PermissionError / IsADirectoryError
        pass
    try:
        return eng.count_file(fqfile, _is_gz(fqfile), limit)
    except _native.TdgError as e:
        if e.code == _native.TDG_ERR_GZIP:
            raise gzip.BadGzipFile(e.message)
        if e.code == _native.TDG_ERR_IO:
            raise OSError(e.message)
        raise


def find_tags_fastq(fqfile, barcodes, tags, cutsite="TGCAG", maxreads=5e9, tassel_tagcount=False,
                    device=None, totals=None):
    """Count reads of ``fqfile`` per (barcode, tag): the GPU replacement of
    tagdigger_fun.find_tags_fastq (tagdigger_fun.py:192-277).  Returns a
    ``len(barcodes) x len(tags)`` list of lists of int.

    ``totals`` (optional list) receives [reads, reads with barcode and cut
    site, reads with tag] -- the running totals the reference prints."""
    if tassel_tagcount:
        raise NotImplementedError("tassel_tagcount=True (count= weights, tagdigger_fun.py:251-253) "
                                  "is not implemented on the GPU path yet")
    p = matchset.plan(barcodes, tags, cutsite)
    limit = _native.limit_from_maxreads(maxreads)
    if p.barnum == 0 or p.ntags == 0:
        # the reference's trie builder indexes an empty list (tagdigger_fun.py:76)
        raise IndexError("list index out of range")
    eng = get_engine(device)
    load_plan(eng, p, nrows=p.barnum)
    tot = _run_file(eng, fqfile, limit)
    counts = eng.read_matrix()
    print("Reads: {0} With barcode and cut site: {1} With tag: {2}".format(tot[0], tot[1], tot[2]))
    if totals is not None:
        totals[:] = tot[:3]
    return counts.tolist()


def find_tags_bytes(data, barcodes, tags, cutsite="TGCAG", maxreads=5e9, device=None, totals=None,
                    pieces=None):
    """Same as find_tags_fastq on an in-memory FASTQ image (bytes), streamed
    through tdg_submit.  ``pieces`` optionally lists split points to submit the
    image in several calls (exercises the carry-over of partial lines)."""
    p = matchset.plan(barcodes, tags, cutsite)
    limit = _native.limit_from_maxreads(maxreads)
    eng = get_engine(device)
    load_plan(eng, p, nrows=p.barnum)
    arr = np.frombuffer(data, dtype=np.uint8)
    cuts = [0] + sorted(pieces or []) + [arr.size]
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b > a:
            eng.submit(arr[a:b], limit)
    eng.end_file(limit)
    tot = eng.file_totals()
    if totals is not None:
        totals[:] = tot
    return eng.read_matrix().tolist()
