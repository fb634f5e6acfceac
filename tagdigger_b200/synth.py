"""Deterministic synthetic GBS inputs of the BASELINE.json shapes (host side).

Everything here is numpy-vectorised so that a few million reads take seconds.
The read mix follows SURVEY.md section 8(d): barcode + cut site + known tag +
random tail; barcode + cut site + unknown sequence; no valid barcode; near
misses (one substitution inside barcode/cut site/tag); plus reads with one
``N``, lower-case reads, reads too short to hold a tag, and quality lines that
begin with ``@`` or ``+``.  The generator also returns the ground truth it
built each read from, which large-size tests use as a size-independent check.
"""

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_HDR_ALPHABET = np.frombuffer(b"ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789:", dtype=np.uint8)


def _random_bases(rng, shape):
    return _ACGT[rng.integers(0, 4, size=shape, dtype=np.uint8)]


def make_barcodes(n, rng, lo=4, hi=9, cutsite="TGCAG"):
    """``n`` distinct ACGT barcodes of length lo..hi such that the set
    {barcode + cutsite} is prefix-free (a requirement of any real key file:
    the reference's index build fails or silently drops otherwise)."""
    chosen = []
    pats = []
    while len(chosen) < n:
        L = int(rng.integers(lo, hi + 1))
        bc = _random_bases(rng, L).tobytes().decode()
        pat = bc + cutsite
        if any(p.startswith(pat) or pat.startswith(p) for p in pats):
            continue
        chosen.append(bc)
        pats.append(pat)
    return chosen


def make_marker_pairs(npairs, rng, length=64, cutsite="TGCAG", min_snp_pos=None,
                      lengths=None):
    """Biallelic marker pairs: two tags that start with the cut site and differ
    at one position.  Returns (marker names, [(allele0, allele1)], sequences
    [(tag0, tag1)]).  ``lengths`` (array of per-pair tag lengths incl. cut
    site) overrides ``length`` for variable-length sets."""
    cl = len(cutsite)
    if lengths is None:
        lengths = np.full(npairs, length, dtype=np.int64)
    maxlen = int(lengths.max())
    body = _random_bases(rng, (npairs, maxlen))
    if cl:
        body[:, :cl] = np.frombuffer(cutsite.encode(), dtype=np.uint8)
    lo = cl if min_snp_pos is None else min_snp_pos
    snp = rng.integers(lo, lengths, size=npairs)
    alt = body.copy()
    rows = np.arange(npairs)
    # substitute a different base at the SNP position
    code = np.searchsorted(_ACGT, body[rows, snp])
    alt[rows, snp] = _ACGT[(code + rng.integers(1, 4, size=npairs)) % 4]
    names, alleles, seqs = [], [], []
    seen = set()
    for i in range(npairs):
        L = int(lengths[i])
        a = body[i, :L].tobytes().decode()
        b = alt[i, :L].tobytes().decode()
        if a in seen or b in seen:
            continue
        seen.add(a)
        seen.add(b)
        names.append("M%06d" % i)
        alleles.append((a[snp[i]], b[snp[i]]))
        seqs.append((a, b))
    return names, alleles, seqs


def merged_csv(names, alleles, seqs, snp_positions=None):
    """The marker set as a 'Merged' tag CSV (Marker name, Tag sequence with
    ``[A/C]``), as text."""
    out = ["Marker name,Tag sequence"]
    for name, (a, b) in zip(names, seqs):
        pos = next(i for i in range(len(a)) if a[i] != b[i])
        out.append("%s,%s[%s/%s]%s" % (name, a[:pos], a[pos], b[pos], a[pos + 1:]))
    return "\n".join(out) + "\n"


def key_csv(fastq_name, barcodes, samples):
    out = ["File,Barcode,Sample"]
    for bc, s in zip(barcodes, samples):
        out.append("%s,%s,%s" % (fastq_name, bc, s))
    return "\n".join(out) + "\n"


def make_fastq(nreads, barcodes, tags, rng, cutsite="TGCAG", readlen=100,
               tags_include_cutsite=True, p_hit=0.60, p_unknown=0.20,
               p_nobarcode=0.15, p_nearmiss=0.05, p_n=0.02, p_lower=0.001,
               p_short=0.001, p_qual_at=0.01, zipf=1.0, newline=b"\n",
               adapter_tail=None):
    """Build a FASTQ image.

    ``tags`` are full tag sequences (including the cut site when
    ``tags_include_cutsite``).  Returns ``(bytes, truth)`` where ``truth`` is a
    dict with the per-read construction: ``kind`` (0 hit, 1 unknown, 2 no
    barcode, 3 near miss), ``barcode``, ``tag``, ``mutated`` (an N or a
    substitution landed inside barcode+cutsite+tag, or the read was cut short),
    and ``expected`` -- the (len(barcodes) x len(tags)) int64 matrix of reads
    that must be counted by construction.

    ``adapter_tail``: optional bytes appended after a (shorter) insert to model
    adapter read-through (two-enzyme libraries); the insert is then the tag
    followed by 0-40 random bases.
    """
    nb, nt = len(barcodes), len(tags)
    cl = len(cutsite)
    bl = np.array([len(b) for b in barcodes], dtype=np.int64)
    tl = np.array([len(t) for t in tags], dtype=np.int64)
    maxbl, maxtl = int(bl.max()), int(tl.max())
    bmat = np.full((nb, max(maxbl, 1)), ord("A"), dtype=np.uint8)
    for i, b in enumerate(barcodes):
        bmat[i, :len(b)] = np.frombuffer(b.encode(), dtype=np.uint8)
    tmat = np.full((nt, maxtl), ord("A"), dtype=np.uint8)
    for i, t in enumerate(tags):
        tmat[i, :len(t)] = np.frombuffer(t.encode(), dtype=np.uint8)
    cs = np.frombuffer(cutsite.encode(), dtype=np.uint8) if cl else np.zeros(0, np.uint8)

    kind = rng.choice(4, size=nreads, p=[p_hit, p_unknown, p_nobarcode, p_nearmiss])
    bidx = rng.integers(0, nb, size=nreads)
    # Zipf-ish tag popularity so that hot cells exist
    w = 1.0 / np.arange(1, nt + 1) ** zipf
    w /= w.sum()
    tidx = rng.choice(nt, size=nreads, p=w)

    seq = _random_bases(rng, (nreads, readlen))
    j = np.arange(readlen)[None, :]
    b_len = bl[bidx][:, None]
    t_len = tl[tidx][:, None]
    has_bar = (kind != 2)[:, None]
    has_tag = ((kind == 0) | (kind == 3))[:, None]
    # barcode
    m = has_bar & (j < b_len)
    seq = np.where(m, bmat[bidx[:, None], np.minimum(j, bmat.shape[1] - 1)], seq)
    # cut site
    if cl:
        rel = j - b_len
        m = has_bar & (rel >= 0) & (rel < cl)
        seq = np.where(m, cs[np.clip(rel, 0, cl - 1)], seq)
    # tag (tags that include the cut site start right after the barcode)
    tstart = b_len if tags_include_cutsite else b_len + cl
    rel = j - tstart
    m = has_tag & (rel >= 0) & (rel < t_len)
    seq = np.where(m, tmat[tidx[:, None], np.clip(rel, 0, maxtl - 1)], seq)
    span = (tstart + t_len)[:, 0]            # end of barcode+cutsite+tag

    if adapter_tail is not None:
        at = np.frombuffer(adapter_tail, dtype=np.uint8)
        ins_end = span + rng.integers(0, 41, size=nreads)
        rel = j - ins_end[:, None]
        m = has_tag & (rel >= 0) & (rel < len(at))
        seq = np.where(m, at[np.clip(rel, 0, len(at) - 1)], seq)

    mutated = np.zeros(nreads, dtype=bool)
    rows = np.arange(nreads)
    # near miss: one substitution inside barcode+cutsite+tag
    nm = kind == 3
    pos = (rng.random(nreads) * np.minimum(span, readlen)).astype(np.int64)
    code = np.searchsorted(_ACGT, seq[rows, pos])
    sub = _ACGT[(code + rng.integers(1, 4, size=nreads)) % 4]
    seq[rows[nm], pos[nm]] = sub[nm]
    mutated |= nm
    # one N at a uniform position
    hasn = rng.random(nreads) < p_n
    npos = rng.integers(0, readlen, size=nreads)
    seq[rows[hasn], npos[hasn]] = ord("N")
    mutated |= hasn & (npos < span)
    # lower case (does not change the outcome)
    lower = rng.random(nreads) < p_lower
    seq[lower] |= 0x20
    # short reads
    short = rng.random(nreads) < p_short
    seqlen = np.full(nreads, readlen, dtype=np.int64)
    seqlen[short] = rng.choice([3, 11, 19], size=int(short.sum()))
    mutated |= short & (seqlen < span)
    mutated |= span > readlen

    good = (kind == 0) & ~mutated
    expected = np.zeros((nb, nt), dtype=np.int64)
    np.add.at(expected, (bidx[good], tidx[good]), 1)

    qual = rng.integers(35, 75, size=(nreads, readlen), dtype=np.uint8)
    qa = rng.random(nreads) < p_qual_at
    qual[qa, 0] = np.where(rng.random(int(qa.sum())) < 0.5, ord("@"), ord("+"))
    hl = rng.integers(30, 61, size=nreads)
    hdr = _HDR_ALPHABET[rng.integers(0, len(_HDR_ALPHABET), size=(nreads, 60))]
    hdr[:, 0] = ord("@")

    nl = np.frombuffer(newline, dtype=np.uint8)
    k = len(nl)
    reclen = hl + seqlen + 1 + seqlen + 4 * k
    offs = np.zeros(nreads + 1, dtype=np.int64)
    np.cumsum(reclen, out=offs[1:])
    out = np.empty(int(offs[-1]), dtype=np.uint8)
    key = hl * 1024 + seqlen
    for kv in np.unique(key):
        sel = np.nonzero(key == kv)[0]
        h, s = int(kv) // 1024, int(kv) % 1024
        block = np.concatenate(
            [hdr[sel, :h], np.broadcast_to(nl, (len(sel), k)),
             seq[sel, :s], np.broadcast_to(nl, (len(sel), k)),
             np.full((len(sel), 1), ord("+"), np.uint8), np.broadcast_to(nl, (len(sel), k)),
             qual[sel, :s], np.broadcast_to(nl, (len(sel), k))], axis=1)
        # scatter fixed-width rows to their record offsets, in slabs to bound memory
        width = block.shape[1]
        step = max(1, (1 << 24) // width)
        for a in range(0, len(sel), step):
            dst = offs[sel[a:a + step]][:, None] + np.arange(width)[None, :]
            out[dst] = block[a:a + step]
    truth = dict(kind=kind, barcode=bidx, tag=tidx, mutated=mutated,
                 expected=expected, nreads=nreads)
    return out.tobytes(), truth


# ---- the table shapes of BASELINE.json's configs (bench + parity tests) --------------------------

SHAPES = {
    # name: barcodes, marker pairs, read length, tag lengths incl. cut site (None = 64), blank barcode,
    #       cut site as given to the counter, concrete site written into tags/reads, read-mix overrides
    "C2": dict(nbar=96, npairs=20000, readlen=100),
    "C3": dict(nbar=1, npairs=20000, readlen=100, blank=True, mix=dict(p_nobar=0.05, p_unknown=0.30)),
    "C4": dict(nbar=384, npairs=250000, readlen=100, lengths=(20, 64)),
    "C4-150": dict(nbar=384, npairs=250000, readlen=150, lengths=(20, 64)),
    "C4-stacks": dict(nbar=384, npairs=250000, readlen=150, lengths=(80, 140)),
    "C5": dict(nbar=96, npairs=20000, readlen=100, lengths=(30, 64)),
    "ApeKI": dict(nbar=96, npairs=20000, readlen=100, cutsite="CWGC", site="CAGC"),
    "short": dict(nbar=96, npairs=20000, readlen=50, lengths=(30, 30)),
}


def shape_tables(name, seed=7, npairs=None):
    """Barcodes and tags of a named shape: ``(barcodes, tags, cutsite, site, readlen, mix)``.
    Variable-length random tags overlap now and then; those markers are dropped the way
    ``tagdigger_script`` would (``sanitizeTags``), so the tag set is one the reference accepts."""
    import contextlib
    import io
    from . import hostio
    sh = SHAPES[name]
    rng = np.random.default_rng(seed)
    cutsite = sh.get("cutsite", "TGCAG")
    site = sh.get("site", cutsite)
    bcs = [""] if sh.get("blank") else make_barcodes(sh["nbar"], rng, cutsite=site)
    n = npairs or sh["npairs"]
    lens = None
    if sh.get("lengths"):
        lens = rng.integers(sh["lengths"][0], sh["lengths"][1] + 1, size=n)
    mnames, _, seqs = make_marker_pairs(n, rng, cutsite=site, lengths=lens)
    tags = [s for p in seqs for s in p]
    if sh.get("lengths") and sh["lengths"][0] != sh["lengths"][1]:
        names = ["%s_%d" % (m, k) for m in mnames for k in (0, 1)]
        with contextlib.redirect_stdout(io.StringIO()):
            tags = hostio.sanitizeTags([names, tags])[1]
    return bcs, tags, cutsite, site, sh["readlen"], dict(sh.get("mix", {}))
