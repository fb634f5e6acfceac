"""Barcode splitter: one FASTQ file in, one trimmed FASTQ file per barcode out.

Drop-in for the reference's ``barcodeSplitter`` (/root/reference/tagdigger_fun.py:1286-1368):
same signature, same asserts, same messages, byte-identical output files.  The two decisions
that cost the reference its time -- which barcode a read carries (:1333) and where the genomic
part ends (findAdapterSeq, :1337-1339; ~90 % of its run time are Python trie walks) -- are
taken on the GPU for a block of reads at a time (``tdg_split_batch``: one warp per read, warp
ballots over positions and adapter prefixes).  Reading, slicing and writing text stay on the
host, in input order, exactly as the reference does them.
"""

import csv
import gzip
import hashlib

from . import hostio, matchset, trimming
from .counting import get_engine

BLOCK_READS = 200000          # reads decided per GPU call


def _byte_to_char_index(text, raw, index):
    """Slice index in characters for an index the device computed in bytes
    (differs only for sequence lines with non-ASCII characters)."""
    if index < 0 or index == 999 or len(raw) == len(text):
        return index
    return len(raw[:index].decode("utf-8", errors="ignore"))


def barcodeSplitter(inputFile, barcodes, outputFiles, cutsite="TGCAG", adapter=hostio.adapters["PstI-MspI-Hall"],
                    maxreads=500000000, device=None):
    assert set(cutsite) <= hostio.BASES, "Only ACGT cut sites allowed."
    assert all([set(bc) <= hostio.BASES for bc in barcodes]), "Found non-ACGT barcodes."
    assert len(adapter) == 2
    assert all([set(a[0]) <= set("ACGT^") for a in adapter])
    assert set(adapter[0][1]) <= hostio.BASES
    assert set(adapter[1][1]) <= set("[barcode]ACGT")

    print("Building indices for rapid searching...")
    barlen = [len(bc) for bc in barcodes]
    barcut = hostio.combine_barcode_and_cutsite(barcodes, cutsite)
    patterns = matchset.effective_set(barcut, len(barcut))          # raises like build_sequence_tree
    eng = get_engine(device)
    tables = trimming.trim_tables(adapter, barcodes)                 # prints the reference's overlap messages
    eng.begin_file(patterns.patterns, patterns.index, [len(p) for p in patterns.patterns], any_base=patterns.any_base)
    eng.set_trim(tables[0], tables[1], tables[2], tables[3], tables[4])
    cutlen = len(cutsite)
    print("Done with indexing setup.")
    print(inputFile)

    if inputFile[-2:].lower() == "gz":
        fqcon = gzip.open(inputFile, "rt")
    else:
        fqcon = open(inputFile, "r")
    outcons = [open(name, mode="w") for name in outputFiles]
    counts = [0, 0, 0]                 # reads, with barcode and cut site, clipped on 3' end

    def flush(block):
        """Decide one block on the GPU, write it out in input order."""
        seqs = [rec[1] for rec in block]
        raws = [s.encode("utf-8") for s in seqs]
        bars, cuts = eng.split_batch(raws, barlen, cutlen)
        for (comment1, sequence, comment2, quality), raw, b, cut in zip(block, raws, bars.tolist(), cuts.tolist()):
            counts[0] += 1
            if b > -1:
                counts[1] += 1
                slice1 = barlen[b]
                if cut == 999:
                    slice2 = len(sequence)
                else:
                    slice2 = _byte_to_char_index(sequence, raw, cut)
                    counts[2] += 1
                out = outcons[b]
                head = comment1 + barcodes[b] + "\n"
                out.write(head)
                out.write(sequence[slice1:slice2] + "\n")
                out.write("+\n" if comment2 == "+" else head)
                out.write(quality[slice1:slice2] + "\n")
            if counts[0] % 1000000 == 0:
                print(inputFile)
            if counts[0] % 50000 == 0:
                print("Reads: {0} With barcode and cut site: {1} Clipped on 3' end: {2}".format(*counts))

    try:
        block = []
        comment1 = sequence = comment2 = ""
        nreads = 0
        for lineindex, line in enumerate(fqcon):
            phase = lineindex % 4
            if phase == 0:
                comment1 = line.strip()
            elif phase == 1:
                sequence = line.strip().upper()
            elif phase == 2:
                comment2 = line.strip()
            else:
                block.append((comment1, sequence, comment2, line.strip()))
                nreads += 1
                if len(block) >= BLOCK_READS:
                    flush(block)
                    block = []
                if nreads >= maxreads:
                    break
        if block:
            flush(block)
    finally:
        fqcon.close()
        for o in outcons:
            o.close()
    return None


def writeMD5sums(filelist, outfile):
    """CSV of file names and MD5 checksums (tagdigger_fun.py:1370-1386)."""
    width = max([len(f) for f in filelist])
    with open(outfile, mode="w", newline="") as con:
        w = csv.writer(con)
        w.writerow(["File name", "MD5 sum"])
        for f in filelist:
            md5 = hashlib.md5()
            with open(f, "rb") as fq:
                for chunk in iter(lambda: fq.read(50 * 1048576), b""):
                    md5.update(chunk)
            w.writerow([f, md5.hexdigest()])
            print("{:>{width}} {}".format(f, md5.hexdigest(), width=width))
    return None
