"""Host side of the trim decision: which adapter prefixes can be found at the end
of a read, and with which slice index.

The reference builds, per barcode, a trie over the REVERSED prefixes of
``site remnant + adapter`` for the common and the rare cutter
(build_adapter_tree, /root/reference/tagdigger_fun.py:1208-1249) and looks the
reversed read up in it (findAdapterSeq, :1251-1283).  Only the patterns the trie
builder leaves reachable matter; :func:`trim_tables` derives them with the same
rules as the counting path (matchset.effective_set) -- including the reference's
fallback when prefixes of the two strings collide (:1237-1248: sort, drop whatever
extends its sorted predecessor, index EVERY remaining prefix with the rare cutter's
remnant length) and the messages it prints on that path.  The decision itself
runs on the GPU (csrc/tdg_trim.cuh) through ``Engine.set_trim`` /
``Engine.trim_batch``.
"""

from . import hostio, matchset


def _prefixes(site, adapter):
    """(forward string, remnant length, reversed prefixes longest first, their slice indices)."""
    keep = site.find("^")
    full = site[:keep] + adapter
    rev = full[::-1]
    slices = [rev[i:] for i in range(len(rev) - keep)]
    return full, keep, slices, [keep - len(s) for s in slices]


def trim_tables(adapter, barcodes):
    """Per barcode: the rare-cutter string and the list of reachable candidates
    ``(which string, prefix length, slice index)``.  Returns
    ``(site0, site1, a0, [a1 per barcode], [candidates per barcode])``."""
    a0, keep0, slices0, index0 = _prefixes(adapter[0][0], adapter[0][1])
    a1_all, cands = [], []
    for bc in barcodes:
        a1, keep1, slices1, index1 = _prefixes(adapter[1][0],
                                               adapter[1][1].replace("[barcode]", hostio.reverseComplement(bc)))
        both = slices0 + slices1
        indices = index0 + index1
        try:
            eff = matchset.effective_set(both, len(both))
            mine = []
            for pat, pos in zip(eff.patterns, eff.index):
                mine.append((0 if pos < len(slices0) else 1, len(pat), indices[pos]))
        except AssertionError:
            print("Some overlap of adapter sequence for barcode {}.".format(bc))
            ordered = sorted(both)
            kept = []
            for k, pat in enumerate(ordered):
                if k > 0 and pat.startswith(ordered[k - 1]):
                    print("Won't search for {0} at end of sequence since {1} is already being searched for."
                          .format(pat[::-1], ordered[k - 1][::-1]))
                else:
                    kept.append(pat)
            eff = matchset.effective_set(kept, len(kept))
            mine = [(0 if a0.startswith(pat[::-1]) else 1, len(pat), keep1 - len(pat)) for pat in eff.patterns]
        a1_all.append(a1)
        cands.append(mine)
    return adapter[0][0].replace("^", ""), adapter[1][0].replace("^", ""), a0, a1_all, cands


def load_trim(eng, adapter, barcodes):
    """Upload the trim tables of (adapter set, barcodes) to an Engine."""
    site0, site1, a0, a1_all, cands = trim_tables(adapter, barcodes)
    eng.set_trim(site0, site1, a0, a1_all, cands)
    return site0, site1


def find_adapter_seqs(eng, sequences, barindices, searchstarts):
    """slice2 of every read (999 = no 3' trim): the batched, on-device form of
    findAdapterSeq(sequence, adaptertrees[barindex], fullsite0, fullsite1, searchstart)."""
    return eng.trim_batch(sequences, barindices, searchstarts)
