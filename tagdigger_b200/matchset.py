"""Match-set compiler: which patterns can ever match, and with which index.

The reference looks reads up in a 4-ary trie built by ``build_sequence_tree``
(/root/reference/tagdigger_fun.py:71-113).  The GPU path uses hash tables of
packed keys instead, which gives the identical answer provided the table holds
exactly the patterns that are *reachable* in that trie.  This module derives
that set -- and the exceptions the trie builder would raise -- directly from the
pattern list, without building a trie:

* the index of pattern ``i`` is ``i mod numseq`` (:102-108);
* walking down from the root, the first pattern (in list order) among those
  that pass through a node decides: if it ends at that node, the node is a leaf
  with that pattern's index and every other pattern through it -- equal or
  longer -- is silently unreachable (:76-77); otherwise, if any other pattern
  ends there, the builder raises ``AssertionError("Problematic sequence: ...")``
  naming the first such pattern in list order (:81-82);
* nodes are visited depth first in A, C, G, T order (:88-96), so when several
  nodes would raise, the lexicographically first one does;
* ``['']`` with ``numseq == 1`` is the special tree that matches any A/C/G/T
  without consuming it (:109-110); an empty list raises ``IndexError`` (:76).

In a sorted list of the distinct patterns, everything that extends a pattern
``v`` sits contiguously right behind it, so one linear scan finds each
"outermost" pattern (no proper prefix of it is in the set) together with the
group of patterns extending it; by the rules above only the outermost pattern's
node is ever inspected.
"""


class EffectiveSet(object):
    """patterns[i] is reachable and returns index[i]; ``any_base`` marks the
    single-empty-pattern tree."""

    __slots__ = ("patterns", "index", "any_base")

    def __init__(self, patterns, index, any_base=False):
        self.patterns = patterns
        self.index = index
        self.any_base = any_base


class DegenerateTree(Exception):
    """The first of several patterns is empty: the reference builds a tree whose
    root is a leaf and then fails with IndexError/TypeError on the first read it
    looks up (tagdigger_fun.py:128).  Raised at set-up time here."""


def effective_set(sequences, numseq):
    """Effective (reachable, prefix-free) part of ``sequences`` under the
    reference's trie rules, in first-occurrence order.  Raises what
    ``build_sequence_tree(sequences, numseq)`` raises."""
    sequences = list(sequences)
    if numseq == 1 and sequences == [""]:
        return EffectiveSet([""], [0], any_base=True)
    if len(sequences) == 0:
        raise IndexError("list index out of range")
    first = {}
    for i, s in enumerate(sequences):
        if s not in first:
            first[s] = i
    if "" in first:
        if first[""] == 0:
            raise DegenerateTree("the first pattern is empty")
        raise AssertionError("Problematic sequence: {}.  Likely due to overlapping tags."
                             .format(first[""] % numseq))
    ordered = sorted(first)
    keep = []
    n = len(ordered)
    i = 0
    while i < n:
        v = ordered[i]
        j = i + 1
        earliest = None
        while j < n and ordered[j].startswith(v):
            p = first[ordered[j]]
            if earliest is None or p < earliest:
                earliest = p
            j += 1
        if earliest is not None and earliest < first[v]:
            raise AssertionError("Problematic sequence: {}.  Likely due to overlapping tags."
                                 .format(first[v] % numseq))
        keep.append(first[v])
        i = j
    keep.sort()
    return EffectiveSet([sequences[p] for p in keep], [p % numseq for p in keep])


_IUPAC = (("R", "AG"), ("Y", "CT"), ("K", "GT"), ("M", "AC"), ("S", "CG"), ("W", "AT"),
          ("B", "CGT"), ("D", "AGT"), ("H", "ACT"), ("V", "ACG"), ("N", "ACGT"))


def enumerate_cut_sites(cutsite):
    """Concrete sites of an IUPAC cut site, in the reference's order
    (tagdigger_fun.py:136-190): codes are resolved in the fixed order
    R, Y, K, M, S, W, B, D, H, V, N; each occurrence (leftmost first) multiplies
    the list, letter-major."""
    sites = [cutsite]
    for code, letters in _IUPAC:
        while code in sites[0]:
            sites = [s.replace(code, letter, 1) for letter in letters for s in sites]
    return sites


class CountPlan(object):
    """Everything the device needs for one (barcodes, tags, cutsite) set-up."""

    __slots__ = ("barnum", "ntags", "bar", "bar_tag_off", "tags")


class TagPlan(object):
    """The tag half of a CountPlan: depends on (tags, cutsite) only, so one key file's worth of
    FASTQ files can share it (the reference rebuilds its tag trie for every file)."""

    __slots__ = ("ntags", "tags", "strip", "shift", "sites", "cutlen")


def plan_tags(tags, cutsite):
    """tagdigger_fun.py:200-204, :207, :221-233 -- everything that does not involve barcodes."""
    cutsite = cutsite.upper()
    assert set(cutsite) <= set("ACGTNRYKMSWBDHV"), "Invalid cut site."
    tags = [t.upper() for t in tags]
    assert all([set(t) <= set("ACGT") for t in tags]), "Non-ACGT tag."
    cutlen = len(cutsite)
    sites = enumerate_cut_sites(cutsite)
    tp = TagPlan()
    tp.ntags = len(tags)
    tp.sites = sites
    tp.cutlen = cutlen
    tp.strip = tp.shift = False
    if set(t[:cutlen] for t in tags) <= set(sites):
        if len(sites) == 1:
            tags = [t[cutlen:] for t in tags]       # one site: it is stripped from every tag
            tp.strip = True
        else:
            tp.shift = True                         # several sites: tags stay whole, comparison starts AT the site
    tp.tags = effective_set(tags, len(tags))
    return tp


def plan(barcodes, tags, cutsite, tagplan=None):
    """Set-up half of find_tags_fastq (tagdigger_fun.py:198-233): same asserts in the same
    order, same pattern lists, same offsets; the two tries become effective sets.
    ``tagplan`` (from :func:`plan_tags` for the same tags and cut site) skips the tag half."""
    assert all([set(b.upper()) <= set("ACGT") for b in barcodes]), "Non-ACGT barcode."
    assert set(cutsite.upper()) <= set("ACGTNRYKMSWBDHV"), "Invalid cut site."
    if tagplan is None:
        assert all([set(t.upper()) <= set("ACGT") for t in tags]), "Non-ACGT tag."
    cutsite = cutsite.upper()
    cutlen = len(cutsite)
    offsets = [len(b) + cutlen for b in barcodes]
    barnum = len(barcodes)
    sites = enumerate_cut_sites(cutsite)
    barcut = [(b + site).upper() for site in sites for b in barcodes]
    p = CountPlan()
    p.barnum = barnum
    p.bar = effective_set(barcut, barnum)           # barcode trie first, as in the reference (:219)
    tp = tagplan if tagplan is not None else plan_tags(tags, cutsite)
    p.ntags = tp.ntags
    if tp.shift:
        offsets = [o - cutlen for o in offsets]
    p.bar_tag_off = [offsets[r] for r in p.bar.index]
    p.tags = tp.tags
    return p
