"""The reference's module name, bound to this package: ``import tagdigger_b200.tagdigger_fun
as tagdigger_fun`` gives a user of TagDigger's counting workflow the functions they call
today (/root/reference/tagdigger_fun.py), with ``find_tags_fastq`` running on the GPU.

Host-side functions live in :mod:`tagdigger_b200.hostio`, the counting entry points in
:mod:`tagdigger_b200.counting`, the pattern-set rules in :mod:`tagdigger_b200.matchset`,
the barcode splitter in :mod:`tagdigger_b200.splitter`.
"""

from .counting import count_files, find_tags_fastq                                   # noqa: F401
from .hostio import (adapters, combine_barcode_and_cutsite, combineReadCounts, compareTags,  # noqa: F401
                     enzymes, extractMarkers, isFastq, readBarcodeKeyfile, readMarkerNames,
                     readTags_Columns, readTags_Merged, readTags_pyRAD, readTags_Rows,
                     readTags_Stacks, readTags_TASSELSAM, readTags_UNEAK_FASTA, reverseComplement,
                     sanitizeTags, writeCounts, writeDiploidGeno)
from .matchset import enumerate_cut_sites                                            # noqa: F401
from .splitter import barcodeSplitter, writeMD5sums                                  # noqa: F401
