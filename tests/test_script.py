"""tagdigger_b200.tagdigger_script against whole-script runs recorded from the
reference (tests/golden/script.json): byte-identical count and genotype CSVs.
Argument errors and the host-side bookkeeping are checked without a GPU."""

import base64
import os

import pytest

from conftest import load_golden, materialize
from tagdigger_b200 import counting, hostio, tagdigger_script

GOLD = load_golden("script.json")
OK = [i for i, c in enumerate(GOLD["cases"]) if c["returncode"] == 0]
BAD = [i for i, c in enumerate(GOLD["cases"]) if c["returncode"] != 0]


@pytest.mark.parametrize("i", BAD)
def test_argument_errors(i, in_tmp):
    case = GOLD["cases"][i]
    materialize(GOLD["filesets"][case["fileset"]], in_tmp)
    with pytest.raises(Exception) as err:
        tagdigger_script.main(case["argv"])
    assert case["stderr_tail"][0] == "Exception: " + str(err.value)
    for name in case["outfiles"]:
        assert not os.path.exists(name)


def test_argparse_surface():
    """Every flag of tagdigger_script.py:10-35 is accepted with the same spelling."""
    p = tagdigger_script.build_parser()
    flags = set()
    for a in p._actions:
        flags.update(a.option_strings)
    want = {"-e", "--enzyme", "-c", "--cutsite", "-w", "--directory", "--UNEAKtags", "--MergedTags", "--ColumnTags",
            "--RowTags", "--StacksTags", "--StacksSnps", "--StacksAlleles", "--TASSELSAM", "--pyRADalleles", "-k",
            "--tokeep", "--binaryOnly", "--TASSELkeyFile", "-b", "--barcodefile", "-o", "--outputcounts", "-g",
            "--outputgen"}
    assert want <= flags
    assert sorted(hostio.enzymes) == ["ApeKI", "EcoT22I", "NcoI", "None", "NsiI", "PstI", "SbfI"]
    with pytest.raises(SystemExit):
        p.parse_args(["-e", "XbaI", "-b", "k", "-o", "o"])


def test_global_rows_equal_combine_order():
    bckeys = {"b.fq": [["AA", "CC"], ["s2", "s1"]], "a.fq": [["GG", "TT", "AC"], ["s1", "s3", "s1"]]}
    samples, rows = counting.global_rows(bckeys)
    counts = {"a.fq": [[1, 0], [0, 2], [4, 4]], "b.fq": [[8, 0], [0, 16]]}
    want = hostio.combineReadCounts(counts, bckeys)
    assert samples == want[0] == ["s1", "s3", "s2"]
    got = [[0, 0] for _ in samples]
    for f in counts:
        for k, r in enumerate(rows[f]):
            got[r] = [x + y for x, y in zip(got[r], counts[f][k])]
    assert got == want[1]


def test_assign_files_partitions(tmp_path):
    names = []
    for i, n in enumerate([5000, 100, 100, 4000, 1, 3000, 2500]):
        p = tmp_path / ("f%d.fq" % i)
        p.write_bytes(b"x" * n)
        names.append(str(p))
    for world in (1, 2, 3, 8):
        parts = [counting.assign_files(names, r, world) for r in range(world)]
        assert sorted(f for part in parts for f in part) == sorted(names)
        if world == 2:
            loads = [sum(os.path.getsize(f) for f in part) for part in parts]
            assert abs(loads[0] - loads[1]) <= 1500


@pytest.mark.gpu
@pytest.mark.parametrize("i", OK)
def test_script_outputs_byte_identical(i, in_tmp):
    case = GOLD["cases"][i]
    materialize(GOLD["filesets"][case["fileset"]], in_tmp)
    assert tagdigger_script.main(case["argv"]) == 0
    for name, b64 in case["outfiles"].items():
        with open(name, "rb") as fh:
            assert fh.read() == base64.b64decode(b64), name


@pytest.mark.gpu
def test_batched_equals_per_file(in_tmp):
    """count_files (global rows on the device) == combineReadCounts over find_tags_fastq per file."""
    materialize(GOLD["filesets"]["lanes"], in_tmp)
    tags = hostio.readTags_Merged("tags.csv")
    bckeys = hostio.readBarcodeKeyfile("key.csv")
    per_file = {f: counting.find_tags_fastq(f, bckeys[f][0], tags[1]) for f in bckeys}
    want = hostio.combineReadCounts(per_file, bckeys)
    assert counting.count_files(bckeys, tags[1]) == want


def test_range_bounds_cut_at_line_ends(tmp_path):
    """Byte ranges for sharding one file across ranks: whole lines, covering the file, for LF
    and CRLF; a file without any '\\n' (lone-CR line ends) goes to rank 0 whole."""
    import random
    r = random.Random(4)
    for newline in (b"\n", b"\r\n"):
        lines = [bytes(r.choice(b"ACGT@+I") for _ in range(r.randint(0, 200))) for _ in range(3000)]
        data = newline.join(lines) + newline
        p = tmp_path / "x.fq"
        p.write_bytes(data)
        for world in (2, 3, 8):
            b = counting.range_bounds(str(p), world)
            assert b[0] == 0 and b[-1] == len(data) and b == sorted(b) and len(b) == world + 1
            for cut in b[1:-1]:
                assert data[cut - 1:cut] == b"\n"
    p.write_bytes(b"\r".join(lines))                  # lone-CR file: no '\n' to cut at, rank 0 takes it all
    n = len(b"\r".join(lines))
    assert counting.range_bounds(str(p), 4) == [0, n, n, n, n]
    p.write_bytes(b"ACGT\n")
    assert counting.range_bounds(str(p), 3) == [0, 5, 5, 5]
    assert not counting.shardable(str(p), 2)
