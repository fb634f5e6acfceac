"""Host side of TagDigger's user-facing API: key file, tag files, marker list,
tag hygiene, merge and output.  Pure Python; none of this touches reads.

These functions feed and drain the GPU counting path and fix which rows and
columns the count matrix has, so their OUTCOMES (return values, printed
messages, files written) follow the reference exactly -- every quirk listed in
SURVEY.md appendix A included.  They are written from that behavioural
description (citations are to /root/reference/tagdigger_fun.py) and checked
against outputs recorded from the reference (tests/golden/readers.json,
small_functions.json, script.json).  Membership tests that the reference does
with linear list scans use sets here; the outcomes are the same.
"""

import bisect
import csv
import gzip
import re

BASES = frozenset("ACGT")

# restriction enzyme -> what remains of its site after the barcode (tagdigger_fun.py:19-20)
enzymes = {"ApeKI": "CWGC", "EcoT22I": "TGCAT", "NcoI": "CATGG", "NsiI": "TGCAT",
           "PstI": "TGCAG", "SbfI": "TGCAGG", "None": ""}

_P7_HALL = "CTCAGGCATCACTCGATTCCTCCGTCGTATGCCGTCTTCTGCTTG"
_P7_CLARK = "CTCAGGCATCACTCGATTCCTATCTCGTATGCCGTCTTCTGCTTG"
_P7_POLAND = "AGATCGGAAGAGCGGTTCAGCAGGAATGCCGAGACCGATCTCGTATGCCGTCTTCTGCTTG"
_P5_BARCODED = "[barcode]AGATCGGAAGAGCGTCGTGTAGGGAAAGAGTGTAGATCTCGGTGGTCGCCGTATCATT"
# adapter sets (tagdigger_fun.py:27-47): (site with ^ at the end of genomic sequence, adapter after it)
adapters = {
    "PstI-MspI-Hall": [("CCG^G", _P7_HALL), ("CTGCA^G", _P5_BARCODED)],
    "NsiI-MspI-Hall": [("CCG^G", _P7_HALL), ("ATGCA^T", _P5_BARCODED)],
    "PstI-MspI-Clark": [("CCG^G", _P7_CLARK), ("CTGCA^G", _P5_BARCODED)],
    "NsiI-MspI-Clark": [("CCG^G", _P7_CLARK), ("ATGCA^T", _P5_BARCODED)],
    "PstI-MspI-Poland": [("CCG^G", _P7_POLAND), ("CTGCA^G", _P5_BARCODED)],
}

_COMPLEMENT = {ord("A"): "T", ord("C"): "G", ord("G"): "C", ord("T"): "A"}


def reverseComplement(sequence):
    """tagdigger_fun.py:1203-1206 (characters other than ACGT pass through)."""
    return sequence.translate(_COMPLEMENT)[::-1]


def combine_barcode_and_cutsite(barcodes, cutsite):
    """tagdigger_fun.py:60-69."""
    assert all([set(b.upper()) <= BASES for b in barcodes]), "Non-ACGT barcode."
    assert set(cutsite.upper()) <= BASES, "Invalid cut site."
    return [(b + cutsite).upper() for b in barcodes]


def _open_text(name, gz_rule):
    """Text-mode handle; `gz_rule(name)` decides whether to inflate."""
    if gz_rule(name):
        return gzip.open(name, "rt")
    return open(name, "r")


def isFastq(filename):
    """0 = not FASTQ / cannot open, 1 = plain FASTQ, 2 = gzipped FASTQ, judged by
    the first three lines (tagdigger_fun.py:279-307).  Only failures of open()
    are swallowed; an empty file or a corrupt gzip stream raises, as in the
    reference."""
    gz = filename[-2:].lower() == "gz"
    try:
        con = gzip.open(filename, "rt") if gz else open(filename, "r")
    except IOError:
        return 0
    verdict = 2 if gz else 1
    try:
        if con.readline()[0] != "@":
            verdict = 0
        if not set(con.readline().strip()) <= set("ACGTNacgtn"):
            verdict = 0
        if con.readline()[0] != "+":
            verdict = 0
    finally:
        con.close()
    return verdict


def readBarcodeKeyfile(filename, forSplitter=False):
    """{file: [[barcodes], [samples]]} from the key CSV (tagdigger_fun.py:309-374).
    Problems are printed and give None."""
    wanted = ("Input File", "Barcode", "Output File") if forSplitter else ("File", "Barcode", "Sample")
    try:
        table = {}
        seen = {}                     # file -> set of barcodes
        with open(filename, "r", newline="") as con:
            reader = csv.reader(con)
            cols = None
            nrow = 0                  # rows counted so far (all-blank rows are not counted)
            for row in reader:
                if cols is None:
                    cols = [row.index(w) for w in wanted]      # ValueError: header problem
                    nrow = 1
                    continue
                fname = row[cols[0]].strip()
                barcode = row[cols[1]].strip().upper()
                sample = row[cols[2]].strip()
                if not (fname or barcode or sample):
                    continue
                if not fname:
                    raise Exception("Blank cell found where file name should be in row {}.".format(nrow + 1))
                if not sample:
                    raise Exception("Blank cell found where sample name should be in row {}.".format(nrow + 1))
                if not set(barcode) <= BASES:
                    raise Exception("{0} in row {1} is not a valid barcode.".format(barcode, nrow + 1))
                if fname not in table:
                    table[fname] = [[], []]
                    seen[fname] = set()
                if barcode in seen[fname]:
                    raise Exception("Each barcode can only be present once for each file.")
                seen[fname].add(barcode)
                table[fname][0].append(barcode)
                table[fname][1].append(sample)
                nrow += 1
        if forSplitter:
            outputs = [s for pair in table.values() for s in pair[1]]
            if len(set(outputs)) < len(outputs):
                raise Exception("All output files must have unique names for barcode splitter.")
        return table
    except IOError:
        print("Could not read file {}.".format(filename))
    except ValueError:
        print("File header needed containing '{}', '{}', and '{}'.".format(*wanted))
    except Exception as err:  # noqa: BLE001 - the reference reports every failure this way
        print(err.args[0])
    return None


def compareTags(taglist, trim=True):
    """[(position, [base in each tag])] for the positions where the tags of one
    locus differ (tagdigger_fun.py:376-393).  Unequal lengths: cut to the
    shortest (`trim`) or pad with N, which never counts as a difference."""
    assert type(taglist) is list, "taglist must be list."
    assert all([set(t) <= BASES for t in taglist]), "taglist must be a list of ACGT strings."
    lengths = set(len(t) for t in taglist)
    if len(lengths) > 1:
        if trim:
            taglist = [t[:min(lengths)] for t in taglist]
        else:
            taglist = [t.ljust(max(lengths), "N") for t in taglist]
    out = []
    for i in range(len(taglist[0])):
        column = [t[i] for t in taglist]
        if len(set(column) - {"N"}) > 1:
            out.append((i, column))
    return out


def _reported(io_message):
    """Decorator: the reader's failures are printed and turned into None --
    unreadable files with `io_message`, everything else with the exception's
    first argument (tagdigger_fun.py:465-472 and its siblings)."""
    def wrap(fn):
        def run(filename, *args, **kwargs):
            try:
                return fn(filename, *args, **kwargs)
            except IOError:
                print(io_message.format(filename))
            except Exception as err:  # noqa: BLE001
                print(err.args[0])
            return None
        run.__name__ = fn.__name__
        run.__doc__ = fn.__doc__
        return run
    return wrap


def _uneak_name(line, lineno, filename):
    if line[:3] != ">TP":
        raise Exception("Line {0} of {1} does not start with '>TP'.".format(lineno, filename))
    return line[1:line.rfind("_")]


def _uneak_length(line):
    return int(line[line.rfind("_") + 1:].strip())


@_reported("File {} not readable.")
def readTags_UNEAK_FASTA(filename, toKeep=None):
    """Tag pairs from a TASSEL-UNEAK FASTA (tagdigger_fun.py:395-473): groups of
    four lines; each sequence is cut to the length after the last underscore of
    its header; both tags of a pair are cut to the shorter one; names become
    TPn_query_<base>_<0|1> with 0/1 in alphabetical order of the first
    differing base."""
    keep = None if toKeep is None else set(toKeep)
    names, seqs, known = [], [], set()
    name1 = name2 = seq1 = None
    len1 = len2 = 0
    with open(filename, "r") as con:
        for n, line in enumerate(con):
            phase = n % 4
            if phase == 0:
                name1 = _uneak_name(line, n + 1, filename)
                len1 = _uneak_length(line)
            elif phase == 2:
                name2 = _uneak_name(line, n + 1, filename)
                if name1[:name1.find("_")] != name2[:name2.find("_")]:
                    raise Exception("Tag name in line {0} does not match tag name in line {1}.".format(n + 1, n - 1))
                len2 = _uneak_length(line)
            else:
                seq = line.strip().upper()[:len1 if phase == 1 else len2]
                if not set(seq) <= BASES:
                    raise Exception("Line {0} is not ACGT sequence.".format(n + 1))
                if seq in known:
                    raise Exception("Non-unique sequence found: line {0}.".format(n + 1))
                if phase == 1:
                    seq1 = seq
                    continue
                marker = name1[:name1.find("_")]
                if keep is not None and marker not in keep:
                    continue
                shorter = min(len1, len2)
                if len1 != len2 and seq1[:shorter] == seq[:shorter]:
                    print("{} skipped because tags cannot be distinguished.".format(marker))
                    continue
                first = compareTags([seq1, seq])[0][1]          # IndexError when the tags are identical
                order = ("_0", "_1") if first[0] < first[1] else ("_1", "_0")
                names.append(name1 + "_" + first[0] + order[0])
                names.append(name2 + "_" + first[1] + order[1])
                for t in (seq1[:shorter], seq[:shorter]):
                    seqs.append(t)
                    known.add(t)
    return [names, seqs]


def _header_columns(row, wanted, message):
    if not set(wanted) <= set(row):
        raise Exception(message)
    return [row.index(w) for w in wanted]


def _marker_name(cell, shown=None):
    name = cell.strip()
    if "_" in name:
        raise Exception("Marker {}: marker names cannot contain underscores.".format(name if shown is None else shown))
    return name


@_reported("File {} not readable.")
def readTags_Rows(filename, toKeep=None):
    """One tag per row: Marker name, Allele name, Tag sequence (tagdigger_fun.py:475-514)."""
    keep = None if toKeep is None else set(toKeep)
    names, seqs, known = [], [], set()
    with open(filename, "r") as con:
        cols = None
        for n, row in enumerate(csv.reader(con)):
            if cols is None:
                cols = _header_columns(row, ("Marker name", "Allele name", "Tag sequence"),
                                       "Need 'Marker name', 'Allele name', and 'Tag sequence' in header row.")
                continue
            marker = _marker_name(row[cols[0]])
            if keep is not None and marker not in keep:
                continue
            allele = row[cols[1]].strip()
            tag = row[cols[2]].upper().strip()
            if not set(tag) <= BASES:
                raise Exception("Tag sequence not formatted as ACGT in row {}.".format(n + 1))
            if tag in known:
                raise Exception("Non-unique sequence found: line {0}.".format(n + 1))
            names.append(marker + "_" + allele)
            seqs.append(tag)
            known.add(tag)
    return [names, seqs]


@_reported("File {} not readable.")
def readTags_Columns(filename, toKeep=None):
    """Two alleles per row: Marker name, Tag sequence 0, Tag sequence 1
    (tagdigger_fun.py:516-561); allele names are the differing bases."""
    keep = None if toKeep is None else set(toKeep)
    names, seqs, known = [], [], set()
    with open(filename, "r") as con:
        cols = None
        for n, row in enumerate(csv.reader(con)):
            if cols is None:
                cols = _header_columns(row, ("Marker name", "Tag sequence 0", "Tag sequence 1"),
                                       "Need 'Marker name', 'Tag sequence 0', and 'Tag sequence 1' in header row.")
                continue
            marker = _marker_name(row[cols[0]])
            if keep is not None and marker not in keep:
                continue
            pair = [row[cols[1]].upper().strip(), row[cols[2]].upper().strip()]
            if not set(pair[0] + pair[1]) <= BASES:
                raise Exception("Tag sequence not formatted as ACGT in row {}.".format(n + 1))
            if pair[0] in known or pair[1] in known:
                raise Exception("Non-unique sequence found: line {0}.".format(n + 1))
            seqs.extend(pair)
            known.update(pair)
            diff = compareTags(pair)
            for k in (0, 1):
                names.append("{}_{}_{}".format(marker, "".join(d[1][k] for d in diff), k))
    return [names, seqs]


@_reported("File {} not readable.")
def readTags_Merged(filename, toKeep=None, allowDuplicates=False):
    """One marker per row with the variants in brackets, ACG[T/C]A...
    (tagdigger_fun.py:563-618).  A marker that repeats an earlier sequence is
    skipped with a message instead of failing the whole file."""
    keep = None if toKeep is None else set(toKeep)
    names, seqs, known = [], [], set()
    with open(filename, "r") as con:
        cols = None
        for n, row in enumerate(csv.reader(con)):
            if cols is None:
                cols = _header_columns(row, ("Marker name", "Tag sequence"),
                                       "Need 'Marker name' and 'Tag sequence' in header row.")
                continue
            cell = row[cols[1]]
            if not set("[/]") < set(cell):
                raise Exception("Characters '[/]' not found in row {}.".format(n + 1))
            marker = _marker_name(row[cols[0]], shown=row[cols[0]])
            if keep is not None and marker not in keep:
                continue
            lo, hi = cell.find("["), cell.find("]")
            variants = [v.strip().upper() for v in cell[lo + 1:hi].split("/")]
            tags = [(cell[:lo] + v + cell[hi + 1:]).upper().strip().replace("-", "") for v in variants]
            if not allowDuplicates and any(t in known for t in tags):
                print("Non-unique sequence found: line {0}.".format(n + 1))
                print("Marker {} skipped.".format(marker))
                continue
            seqs.extend(tags)
            known.update(tags)
            if not all(set(t) <= BASES for t in tags):
                raise Exception("Tag sequence not formatted correctly in row {}.".format(n + 1))
            names.extend("{}_{}_{}".format(marker, v, i) for i, v in enumerate(variants))
    return [names, seqs]


def _stacks_rows(path, keep, locus_col):
    """Rows of one Stacks catalog TSV (comment rows and unwanted loci dropped)."""
    with _open_text(path, lambda p: p.endswith(".gz")) as con:
        for row in csv.reader(con, delimiter="\t"):
            if row[0].startswith("#"):
                continue
            if keep is None or row[locus_col] in keep:
                yield row


def readTags_Stacks(tagsfile, snpsfile, allelesfile, toKeep=None, binaryOnly=False, version=1):
    """Tags from a Stacks catalog: haplotype letters substituted into the
    consensus at the SNP columns (tagdigger_fun.py:620-719)."""
    v1 = version == 1
    locus_col = 2 if v1 else 1
    keep = None if toKeep is None else set(toKeep)
    try:
        consensus = {}
        for row in _stacks_rows(tagsfile, keep, locus_col):
            consensus[row[locus_col]] = row[9 if v1 else 5]
        haplotypes = [(row[locus_col], row[3 if v1 else 2]) for row in _stacks_rows(allelesfile, keep, locus_col)]
        snp_cols = {}
        for row in _stacks_rows(snpsfile, keep, locus_col):
            snp_cols.setdefault(row[locus_col], []).append(int(row[3 if v1 else 2]))
        names, seqs = [], []
        for locus, hap in haplotypes:
            seq = consensus[locus]
            if hap:
                cols = snp_cols[locus]
                pieces = [seq[:cols[0]]]
                for i, letter in enumerate(hap):
                    pieces.append(letter)
                    pieces.append(seq[cols[i] + 1:] if i + 1 == len(hap) else seq[cols[i] + 1:cols[i + 1]])
                seq = "".join(pieces)
            seq = seq.upper()
            if set(seq) <= BASES:
                names.append(locus + "_" + hap)
                seqs.append(seq)
            else:
                print("{}_{} skipped for having non-ACGT nucleotides.".format(locus, hap))
        if binaryOnly:
            kept_names, kept_seqs = [], []
            for alleles, where in extractMarkers(names)[1]:
                if len(alleles) != 2:
                    continue
                order = ("_0", "_1") if alleles[0] < alleles[1] else ("_1", "_0")
                for k in (0, 1):
                    kept_names.append(names[where[k]] + order[k])
                    kept_seqs.append(seqs[where[k]])
            names, seqs = kept_names, kept_seqs
        return [names, seqs]
    except IOError:
        print("Files not readable.")
    except (IndexError, ValueError):
        print("Files in wrong format.")
    except KeyError:
        print("Locus names not matching properly.")
    except Exception as err:  # noqa: BLE001
        print(err.args[0])
    return None


_UNALIGNED = frozenset(4 + x for x in (0, 1, 2, 8, 16, 32, 64, 128))
_BOTTOM = frozenset(16 + x for x in (0, 1, 2, 8, 32, 64, 128))


def _sam_markers(filename):
    """{marker: [tag sequences]} of a SAM file under TASSEL-GBSv2 naming
    (tagdigger_fun.py:736-789)."""
    by_marker = {}
    width = 0
    with open(filename, "r") as con:
        for line in con:
            if line[0:3] == "@SQ":
                width = max(width, len(line.split()[2][3:]))
                continue
            if line[0] == "@":
                continue
            f = line.split()
            flags = int(f[1])
            if flags in _UNALIGNED:
                continue
            chrom = f[2].replace("_", "*")
            pos = int(f[3])
            seq = f[9]
            strand = "top"
            if flags in _BOTTOM:
                strand = "bot"
                seq = reverseComplement(seq)
                cigar = f[5]
                gone = sum(int(x[:-1]) for x in re.findall(r"\d+D", cigar))
                added = sum(int(x[:-1]) for x in re.findall(r"\d+I", cigar))
                pos = pos + len(seq) - added + gone - 1          # position of the cut-site end
            marker = "{}-{:0>{width}}-{}".format(chrom, pos, strand, width=width)
            have = by_marker.get(marker)
            if have is None:
                by_marker[marker] = [seq]
                continue
            # keep the shorter of two versions that extend one another
            have = [h for h in have if not h.startswith(seq)]
            if not any(seq.startswith(h) for h in have):
                have.append(seq)
            by_marker[marker] = have
    return by_marker


def readTags_TASSELSAM(filename, toKeep=None, binaryOnly=False, noMonomorphic=False,
                       writeMarkerKey=False, keyfilename=None):
    """Tags of a TASSEL-GBSv2 SAM file, markers named chrom-position-strand
    (tagdigger_fun.py:721-854); `toKeep` holds TASSEL SNP names (S01_1026)."""
    assert (not writeMarkerKey) or keyfilename is not None, "keyfilename needed."
    keep = None if toKeep is None else set(toKeep)
    names, seqs, key_rows = [], [], []
    try:
        by_marker = _sam_markers(filename)
        for marker in sorted(by_marker):
            tags = by_marker[marker]
            if (binaryOnly and len(tags) != 2) or (noMonomorphic and len(tags) == 1):
                continue
            diff = compareTags(tags, trim=False)
            if keep is not None or writeMarkerKey:
                chrom, pos, strand = marker.split("-")[:3]
                chrom = chrom.upper()
                if chrom.startswith("CHROMOSOME"):
                    chrom = chrom[10:]
                if chrom.startswith("CHR"):
                    chrom = chrom[3:]
                step = 1 if strand == "top" else -1
                snps = ["S{}_{}".format(chrom, int(pos) + step * d[0]) for d in diff]
                if keep is not None and not any(s in keep for s in snps):
                    continue
                if writeMarkerKey:
                    key_rows.extend((s, marker) for s in snps)
            alleles = ["".join(d[1][i] for d in diff) for i in range(len(tags))]
            tagnames = [marker + "_" + a for a in alleles]
            if binaryOnly and alleles[0] != alleles[1]:
                lo = 0 if alleles[0] < alleles[1] else 1
                tagnames[lo] += "_0"
                tagnames[1 - lo] += "_1"
            names.extend(tagnames)
            seqs.extend(tags)
        if not names:
            raise Exception("No markers output; is list of markers to keep in right format (e.g. S03_350622)?")
    except IOError:
        print("Could not read file {}.".format(filename))
        return None
    except Exception as err:  # noqa: BLE001
        print(err.args[0])
        return None
    if writeMarkerKey:
        try:
            with open(keyfilename, "w", newline="") as out:
                w = csv.writer(out)
                w.writerow(["TASSEL-GBSv2 marker name", "TagDigger marker name"])
                w.writerows(key_rows)
        except IOError:
            print("Could not write file {}.".format(keyfilename))
            return None
    return [names, seqs]


def _pyrad_locus(sequences, locus, binary_only):
    """Names and sequences of one pyRAD locus (tagdigger_fun.py:864-885)."""
    n = min(len(s) for s in sequences)
    seq = [s[:n] for s in sequences]
    while any(s[-1] == "-" for s in seq):            # trailing gap columns
        seq = [s[:-1] for s in seq]
        n -= 1
    seq = sorted(set(s for s in seq if "N" not in s))
    if not ((seq and not binary_only) or len(seq) == 2):
        return [], []
    variable = [i for i in range(n) if len(set(s[i] for s in seq)) > 1]
    names = ["{}_{}_{}".format(locus, "".join(s[i] for i in variable), k) for k, s in enumerate(seq)]
    return names, [s.replace("-", "") for s in seq]


def readTags_pyRAD(filename, toKeep=None, binaryOnly=False):
    """Tags from a pyRAD .alleles file (tagdigger_fun.py:856-919)."""
    keep = None if toKeep is None else set(toKeep)
    names, seqs = [], []
    pending = set()
    lineno = 0
    try:
        with open(filename, "r") as con:
            for line in con:
                if line[0] == ">":
                    s = line.split()[1]
                    if not set(s) <= set("ACGT-N"):
                        raise Exception("Character other than ACGTN- detected in sequence.")
                    pending.add(s)
                elif line[0] == "/":
                    locus = line.split()[-1][1:-1]
                    for ch in "|*-":
                        locus = locus.replace(ch, "")
                    if keep is None or locus in keep:
                        got = _pyrad_locus(pending, locus, binaryOnly)
                        names.extend(got[0])
                        seqs.extend(got[1])
                    pending = set()
                else:
                    raise Exception("File not in pyRAD format.")
                lineno += 1
        return [names, seqs]
    except IOError:
        print("File {} not readable.".format(filename))
    except Exception as err:  # noqa: BLE001
        print("Line {}:".format(lineno))
        print(err.args[0])
    return None


def readMarkerNames(filename):
    """One marker name per line; commas and surrounding whitespace dropped, blank
    lines ignored (tagdigger_fun.py:921-934)."""
    try:
        with open(filename, "r") as con:
            lines = con.readlines()
    except IOError:
        print("File {} not readable.".format(filename))
        return None
    cleaned = (ln.replace(",", "").strip() for ln in lines)
    return [c for c in cleaned if c != ""]


def sanitizeTags(taglist):
    """Drop the markers whose tag is a prefix of (or equal to) another tag, in
    place (tagdigger_fun.py:1030-1058).  The marker of the SHORTER tag goes --
    together with every tag whose name merely starts with that marker's name --
    and everything removed is printed."""
    assert len(taglist) == 2, "'taglist' should have two elements."
    assert len(taglist[0]) == len(taglist[1]), "List of tag names should be the same as list of tag sequences."
    names, seqs = taglist
    print("\nSanitizing tags...")
    ordered = sorted(seqs)
    for shorter, longer in zip(ordered, ordered[1:]):
        if not longer.startswith(shorter) or shorter not in seqs:
            continue
        tagname = names[seqs.index(shorter)]
        marker = tagname[:tagname.find("_")]
        print("Removing " + marker + " for overlap with another marker.")
        for i in reversed([k for k, nm in enumerate(names) if nm.startswith(marker)]):
            print(names.pop(i))
            print(seqs.pop(i))
    return taglist


def combineReadCounts(countsdict, bckeys):
    """[sample names, count rows] over all files: files in sorted order, samples
    in key-file order, equal sample names summed (tagdigger_fun.py:1061-1098)."""
    files = sorted(bckeys.keys())
    ncols = len(countsdict[files[0]][0])
    nsamples = len(set(s for f in files for s in bckeys[f][1]))
    samples = [""] * nsamples
    totals = [[0] * ncols for _ in range(nsamples)]
    row_of = {}
    for f in files:
        for k, sample in enumerate(bckeys[f][1]):
            r = row_of.get(sample)
            if r is None:
                r = row_of[sample] = len(row_of)
                samples[r] = sample
                totals[r] = countsdict[f][k]
            else:
                mine = countsdict[f][k]
                totals[r] = [mine[i] + totals[r][i] for i in range(ncols)]
    return [samples, totals]


def writeCounts(filename, counts, samnames, tagnames):
    """The count CSV: empty corner cell, tag names across, one row per sample;
    csv default dialect, i.e. \\r\\n line ends (tagdigger_fun.py:1100-1111)."""
    assert len(samnames) == len(counts), "Length of samnames should be the same as length of counts."
    assert len(tagnames) == len(counts[0]), "Length of tagnames should be length of second dimension of counts."
    if _is_int_array(counts):
        # an int32 matrix straight from the device: multithreaded native formatter, same bytes
        from . import _native
        _native.write_counts_csv(filename, counts, samnames, tagnames)
        return
    with open(filename, "w", newline="") as out:
        w = csv.writer(out)
        w.writerow([""] + tagnames)
        for name, row in zip(samnames, counts):
            w.writerow([name] + row)


def _is_int_array(counts):
    """True for a 2-D numpy integer array (the script keeps the device matrix in that form:
    config 4's 192 M cells would be gigabytes of Python integers)."""
    try:
        import numpy as np
    except ImportError:
        return False
    return isinstance(counts, np.ndarray) and counts.ndim == 2 and counts.dtype.kind in "iu"


def extractMarkers(tagnames):
    """[marker names, [[allele names, tag indices] per marker]] in order of first
    appearance; marker = text before the first underscore, allele = text after
    the last (tagdigger_fun.py:1113-1142)."""
    if len(tagnames) != len(set(tagnames)):
        raise Exception("Non-unique tag names found.")
    markers, groups, where = [], [], {}
    for i, name in enumerate(tagnames):
        marker = name[:name.find("_")]
        g = where.get(marker)
        if g is None:
            g = where[marker] = len(markers)
            markers.append(marker)
            groups.append([[], []])
        groups[g][0].append(name[name.rfind("_") + 1:])
        groups[g][1].append(i)
    return [markers, groups]


def writeDiploidGeno(filename, counts, samnames, tagnames):
    """Numeric diploid genotypes per marker: 0 = only allele 0 seen, 1 = both,
    2 = only allele 1, empty = neither (tagdigger_fun.py:1144-1180)."""
    assert len(samnames) == len(counts), "Length of samnames should be the same as length of counts."
    assert len(tagnames) == len(counts[0]), "Length of tagnames should be length of second dimension of counts."
    markers, groups = extractMarkers(tagnames)
    try:
        if not all(set(g[0]) <= {"0", "1"} for g in groups):
            raise Exception("All allele names must be '0' or '1'.")
        if _is_int_array(counts):
            from . import _native
            columns = [(g[1][g[0].index("0")], g[1][g[0].index("1")]) for g in groups] if len(samnames) else []
            _native.write_geno_csv(filename, counts, samnames, markers if columns else [],
                                   [c[0] for c in columns], [c[1] for c in columns])
            return None
        columns = None
        rows = []
        for sample, row in zip(samnames, counts):
            if columns is None:        # ValueError when a marker lacks allele 0 or 1 (only once there are samples)
                columns = [(g[1][g[0].index("0")], g[1][g[0].index("1")]) for g in groups]
            calls = []
            for c0, c1 in columns:
                a, b = row[c0] > 0, row[c1] > 0
                calls.append("1" if a and b else "0" if a else "2" if b else "")
            rows.append([sample] + calls)
        with open(filename, "w", newline="") as out:
            w = csv.writer(out)
            w.writerow([""] + markers)
            w.writerows(rows)
    except IOError:
        print("Could not write file {}.".format(filename))
    except Exception as err:  # noqa: BLE001
        print(err.args[0])
    return None


# bisect is part of the reference module's namespace; keep the name importable
_ = bisect
