"""CPU restatement of TagDigger's per-read counting path.  TEST INFRASTRUCTURE ONLY.

This module is the *oracle*: an independent, plain-Python restatement of the
algorithm in the reference's ``tagdigger_fun.py`` (file:line citations are into
``/root/reference/``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package ``tagdigger_b200`` never does.

Parity pinning: the reference ships no tests, so this
restatement is pinned against outputs of the reference itself, generated in the
build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference/tagdigger_fun.py``) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every vector; when ``/root/reference`` is
present the same tests also run live differential comparisons.

Data structure note: the reference stores its 4-ary index trie as nested Python
lists and walks it recursively (one frame per base).  Here the same trie is kept
in flat integer arrays and walked iteratively; the construction rules (first
pattern through a node wins when it ends there, a later pattern ending at a
node that the first pattern continues through is an ``AssertionError``) are the
reference's, restated.
"""

import csv
import gzip
import io

BASES = "ACGT"

# tagdigger_fun.py:19-20
ENZYMES = {"ApeKI": "CWGC", "EcoT22I": "TGCAT", "NcoI": "CATGG",
           "NsiI": "TGCAT", "PstI": "TGCAG", "SbfI": "TGCAGG", "None": ""}

# tagdigger_fun.py:136-190 -- expansion order of the IUPAC codes, each code
# exhausted (left to right, one occurrence at a time) before the next.
_IUPAC_ORDER = (("R", "AG"), ("Y", "CT"), ("K", "GT"), ("M", "AC"), ("S", "CG"),
                ("W", "AT"), ("B", "CGT"), ("D", "AGT"), ("H", "ACT"),
                ("V", "ACG"), ("N", "ACGT"))


def expand_cut_site(cutsite):
    """All concrete cut sites for an IUPAC cut site (tagdigger_fun.py:136-190)."""
    sites = [cutsite]
    for code, letters in _IUPAC_ORDER:
        while code in sites[0]:
            grown = []
            for letter in letters:
                grown.extend(s.replace(code, letter, 1) for s in sites)
            sites = grown
    return sites


def barcode_patterns(barcodes, cutsite):
    """barcode+cutsite, upper-cased (tagdigger_fun.py:60-69)."""
    if not all(set(b.upper()) <= set(BASES) for b in barcodes):
        raise AssertionError("Non-ACGT barcode.")
    if not set(cutsite.upper()) <= set(BASES):
        raise AssertionError("Invalid cut site.")
    return [(b + cutsite).upper() for b in barcodes]


class Trie(object):
    """Flat 4-ary trie equivalent to the nested lists of tagdigger_fun.py:71-113.

    ``kids[4*n + c]`` is the child of node ``n`` under base ``c`` (0 = absent);
    ``leaf[n]`` is the stored index if node ``n`` terminates a pattern, else -1.
    ``any_base`` marks the special tree of tagdigger_fun.py:109-110 (a single
    empty pattern): any A/C/G/T first character matches index 0.
    ``root_is_leaf`` marks the degenerate tree the reference builds when the
    first of several patterns is empty; looking anything up in it raises, as
    the nested-list version does.
    """

    __slots__ = ("kids", "leaf", "any_base", "root_is_leaf")

    def __init__(self):
        self.kids = [0, 0, 0, 0]
        self.leaf = [-1]
        self.any_base = False
        self.root_is_leaf = False


def build_trie(sequences, numseq):
    """tagdigger_fun.py:98-113 (+ :71-96 for the per-node rule)."""
    trie = Trie()
    if numseq == 1 and sequences == [""]:
        trie.any_base = True
        return trie
    if len(sequences) == 0:
        # tree_one_level indexes num_seq[0] (tagdigger_fun.py:76)
        raise IndexError("list index out of range")
    idx_of = []
    running = 0
    for _ in sequences:               # index = position mod numseq (:102-108)
        idx_of.append(running)
        running += 1
        if running == numseq:
            running = 0
    # explicit DFS in A,C,G,T pre-order, like tree_recursive (:88-96)
    stack = [(0, 0, list(range(len(sequences))))]
    while stack:
        node, depth, members = stack.pop()
        first = members[0]
        if len(sequences[first]) == depth:          # :76-77 first one ends here
            trie.leaf[node] = idx_of[first]
            if node == 0:
                trie.root_is_leaf = True
            continue
        buckets = ([], [], [], [])
        for m in members:                           # :81-85
            s = sequences[m]
            if len(s) <= depth:
                raise AssertionError(
                    "Problematic sequence: {}.  Likely due to overlapping tags."
                    .format(idx_of[m]))
            buckets[BASES.find(s[depth])].append(m)  # -1 -> last bucket, as :83-85
        pending = []
        for c in range(4):
            if buckets[c]:
                child = len(trie.leaf)
                trie.leaf.append(-1)
                trie.kids.extend((0, 0, 0, 0))
                trie.kids[4 * node + c] = child
                pending.append((child, depth + 1, buckets[c]))
        stack.extend(reversed(pending))             # visit A first
    return trie


def lookup(sequence, trie):
    """Index of the stored pattern that prefixes ``sequence``, else -1
    (tagdigger_fun.py:115-134)."""
    if trie.any_base:
        if len(sequence) == 0 or sequence[0] not in BASES:
            return -1
        return 0
    node = 0
    kids = trie.kids
    leaf = trie.leaf
    for ch in sequence:
        c = BASES.find(ch)
        if c < 0:
            return -1
        if trie.root_is_leaf:
            # the reference indexes into ['', idx] here (:128) and blows up
            if c == 1:
                raise TypeError("'int' object is not subscriptable")
            raise IndexError("index out of range")
        node = kids[4 * node + c]
        if node == 0:
            return -1
        if leaf[node] >= 0:
            return leaf[node]
    return -1


def prepare(barcodes, tags, cutsite):
    """Setup half of find_tags_fastq (tagdigger_fun.py:198-233).

    Returns (barcode trie, tag trie, per-barcode tag offset, barnum, ntags)."""
    if not all(set(b.upper()) <= set(BASES) for b in barcodes):
        raise AssertionError("Non-ACGT barcode.")
    cutsite = cutsite.upper()
    if not set(cutsite) <= set("ACGTNRYKMSWBDHV"):
        raise AssertionError("Invalid cut site.")
    tags = [t.upper() for t in tags]
    if not all(set(t) <= set(BASES) for t in tags):
        raise AssertionError("Non-ACGT tag.")
    cutlen = len(cutsite)
    offsets = [len(b) + cutlen for b in barcodes]
    barnum = len(barcodes)
    sites = expand_cut_site(cutsite)
    pats = []
    for site in sites:
        pats.extend(barcode_patterns(barcodes, site))
    bartrie = build_trie(pats, barnum)
    if set(t[:cutlen] for t in tags) <= set(sites):          # :222
        if len(sites) == 1:
            tags = [t[cutlen:] for t in tags]                 # :226
        else:
            offsets = [o - cutlen for o in offsets]           # :231
    tagtrie = build_trie(tags, len(tags))
    return bartrie, tagtrie, offsets, barnum, len(tags)


def open_text(path):
    """gz iff the last two characters of the name are 'gz' (:240-243)."""
    if path[-2:].lower() == "gz":
        return gzip.open(path, "rt")
    return open(path, "r")


def count_lines(lines, bartrie, tagtrie, offsets, barnum, ntags,
                maxreads=5e9, tassel_tagcount=False, totals=None):
    """Loop half of find_tags_fastq (tagdigger_fun.py:245-277) over any
    iterable of text lines."""
    counts = [[0] * ntags for _ in range(barnum)]
    nreads = nbar = ntag = 0
    weight = 1
    lineno = 0
    for line in lines:
        phase = lineno & 3
        if phase == 0 and tassel_tagcount:
            weight = int(line[line.find("count=") + 6:].strip())       # :253
        if phase == 1:
            nreads += 1
            seq = line.strip().upper()                                  # :256
            b = lookup(seq, bartrie)
            if b > -1:
                nbar += 1
                t = lookup(seq[offsets[b]:], tagtrie)
                if t > -1:
                    ntag += 1
                    counts[b][t] += weight if tassel_tagcount else 1
            if nreads >= maxreads:                                      # :272
                break
        lineno += 1
    if totals is not None:
        totals[:] = [nreads, nbar, ntag]
    return counts


def find_tags_fastq(fqfile, barcodes, tags, cutsite="TGCAG", maxreads=5e9,
                    tassel_tagcount=False, totals=None):
    """Whole of tagdigger_fun.py:192-277."""
    bartrie, tagtrie, offsets, barnum, ntags = prepare(barcodes, tags, cutsite)
    with open_text(fqfile) as con:
        return count_lines(con, bartrie, tagtrie, offsets, barnum, ntags,
                           maxreads, tassel_tagcount, totals)


def find_tags_text(text, barcodes, tags, cutsite="TGCAG", maxreads=5e9,
                   tassel_tagcount=False, totals=None):
    """Same, on an in-memory ``bytes`` FASTQ image, decoded and split exactly
    as the reference's text-mode file iteration would (universal newlines)."""
    bartrie, tagtrie, offsets, barnum, ntags = prepare(barcodes, tags, cutsite)
    con = io.TextIOWrapper(io.BytesIO(text), encoding="utf-8", newline=None)
    return count_lines(con, bartrie, tagtrie, offsets, barnum, ntags,
                       maxreads, tassel_tagcount, totals)


def sanitize_tags(names, seqs):
    """tagdigger_fun.py:1030-1058 on copies; returns (names, seqs, removed names)."""
    names = list(names)
    seqs = list(seqs)
    removed = []
    ordered = sorted(seqs)
    for i in range(len(ordered) - 1):
        short = ordered[i]
        if ordered[i + 1].startswith(short) and short in seqs:
            tagname = names[seqs.index(short)]
            marker = tagname[:tagname.find("_")]          # note: find may be -1
            doomed = [j for j in range(len(seqs)) if names[j].startswith(marker)]
            for j in sorted(doomed, reverse=True):
                removed.append(names.pop(j))
                seqs.pop(j)
    return names, seqs, removed


def combine_read_counts(countsdict, bckeys):
    """tagdigger_fun.py:1061-1098."""
    files = sorted(bckeys.keys())
    everyone = set()
    for f in files:
        everyone.update(bckeys[f][1])
    ntags = len(countsdict[files[0]][0])
    total = [[0] * ntags for _ in range(len(everyone))]
    order = [""] * len(everyone)
    nxt = 0
    for f in files:
        for s, sample in enumerate(bckeys[f][1]):
            if sample in order:
                row = order.index(sample)
                total[row] = [a + b for a, b in zip(countsdict[f][s], total[row])]
            else:
                order[nxt] = sample
                total[nxt] = countsdict[f][s]
                nxt += 1
    return [order, total]


def counts_csv_bytes(counts, samnames, tagnames):
    """Bytes that writeCounts (tagdigger_fun.py:1100-1111) puts in the file."""
    buf = io.StringIO(newline="")
    w = csv.writer(buf)
    w.writerow([""] + list(tagnames))
    for name, row in zip(samnames, counts):
        w.writerow([name] + list(row))
    return buf.getvalue().encode("utf-8")


def extract_markers(tagnames):
    """tagdigger_fun.py:1113-1142 (first-appearance marker order)."""
    if len(tagnames) != len(set(tagnames)):
        raise Exception("Non-unique tag names found.")
    markers = []
    where = {}
    alleles = []
    for i, t in enumerate(tagnames):
        m = t[:t.find("_")]
        if m not in where:
            where[m] = len(markers)
            markers.append(m)
            alleles.append([[], []])
        slot = alleles[where[m]]
        slot[0].append(t[t.rfind("_") + 1:])
        slot[1].append(i)
    return [markers, alleles]


def diploid_geno_csv_bytes(counts, samnames, tagnames):
    """Bytes that writeDiploidGeno (tagdigger_fun.py:1144-1180) writes."""
    markers, alleles = extract_markers(tagnames)
    if not all(set(a[0]) <= {"0", "1"} for a in alleles):
        raise Exception("All allele names must be '0' or '1'.")
    buf = io.StringIO(newline="")
    w = csv.writer(buf)
    w.writerow([""] + markers)
    for name, row in zip(samnames, counts):
        cells = []
        for a in alleles:
            c0 = row[a[1][a[0].index("0")]]
            c1 = row[a[1][a[0].index("1")]]
            if c0 > 0 and c1 == 0:
                cells.append("0")
            elif c0 > 0 and c1 > 0:
                cells.append("1")
            elif c0 == 0 and c1 > 0:
                cells.append("2")
            else:
                cells.append("")
        w.writerow([name] + cells)
    return buf.getvalue().encode("utf-8")


# ---------------------------------------------------------------------------
# Trim decision (barcode splitter), tagdigger_fun.py:1203-1283
# ---------------------------------------------------------------------------

def reverse_complement(seq):
    """tagdigger_fun.py:1203-1206 (only A/C/G/T are complemented)."""
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    return "".join(comp.get(ch, ch) for ch in reversed(seq))


def adapter_tables(adapter, barcodes):
    """tagdigger_fun.py:1208-1249: per barcode, (trie over reversed adapter
    prefixes, slice index per stored prefix)."""
    keep0 = adapter[0][0].find("^")
    full0 = adapter[0][0][:keep0] + adapter[0][1]
    rev0 = full0[::-1]
    slices0 = [rev0[i:] for i in range(len(rev0) - keep0)]
    index0 = [keep0 - len(s) for s in slices0]
    tables = []
    for bc in barcodes:
        keep1 = adapter[1][0].find("^")
        full1 = adapter[1][0][:keep1] + adapter[1][1].replace(
            "[barcode]", reverse_complement(bc))
        rev1 = full1[::-1]
        slices1 = [rev1[i:] for i in range(len(rev1) - keep1)]
        index1 = [keep1 - len(s) for s in slices1]
        both = slices0 + slices1
        try:
            tables.append((build_trie(both, len(both)), index0 + index1))
        except AssertionError:
            # :1237-1248 overlap fallback: sort, drop anything that extends its
            # sorted predecessor, and index *every* slice with the rare-cutter
            # remnant length.
            both = sorted(both)
            kept = [both[k] for k in range(len(both))
                    if not (k > 0 and both[k].startswith(both[k - 1]))]
            tables.append((build_trie(kept, len(kept)),
                           [keep1 - len(s) for s in kept]))
    return tables


def find_adapter_seq(sequence, table, fullsite0, fullsite1, searchstart):
    """tagdigger_fun.py:1251-1283: slice index, or 999 for 'no 3' trim'."""
    hit0 = sequence.find(fullsite0, searchstart)
    hit1 = sequence.find(fullsite1, searchstart)
    if hit0 == -1 and hit1 == -1:
        which = lookup(sequence[::-1], table[0])
        return 999 if which == -1 else table[1][which]
    if hit1 == -1:
        return hit0 + len(fullsite0)
    if hit0 == -1:
        return hit1 + len(fullsite1)
    if hit0 < hit1:
        return hit0 + len(fullsite0)
    return hit1 + len(fullsite1)


def split_records(lines, barcodes, cutsite, adapter, maxreads=500000000, every=False):
    """Loop of barcodeSplitter (tagdigger_fun.py:1328-1363) as a generator of
    (barcode index, 4 output lines, slice2 or 999) for every read that matches
    a barcode; used to check trim decisions read by read.  every=True: reads without
    a barcode are reported too, as (-1, None, 999)."""
    if not set(cutsite) <= set(BASES):
        raise AssertionError("Only ACGT cut sites allowed.")
    pats = barcode_patterns(barcodes, cutsite)
    bartrie = build_trie(pats, len(pats))
    tables = adapter_tables(adapter, barcodes)
    full0 = adapter[0][0].replace("^", "")
    full1 = adapter[1][0].replace("^", "")
    cutlen = len(cutsite)
    nreads = 0
    c1 = seq = c2 = ""
    for lineno, line in enumerate(lines):
        phase = lineno & 3
        if phase == 0:
            c1 = line.strip()
        elif phase == 1:
            seq = line.strip().upper()
        elif phase == 2:
            c2 = line.strip()
        else:
            nreads += 1
            qual = line.strip()
            b = lookup(seq, bartrie)
            if b > -1:
                s1 = len(barcodes[b])
                s2 = find_adapter_seq(seq, tables[b], full0, full1, s1 + cutlen)
                cut = len(seq) if s2 == 999 else s2
                head = c1 + barcodes[b]
                yield b, [head, seq[s1:cut], "+" if c2 == "+" else head,
                          qual[s1:cut]], s2
            elif every:
                yield -1, None, 999
            if nreads >= maxreads:
                break
