"""BASELINE config 5 as a flow (tests/golden/flow.json, recorded from the reference's own two
scripts): a PstI-MspI library with adapter read-through is split by barcode and trimmed
(barcode_splitter_script, /root/reference/barcode_splitter_script.py:8-36 -> barcodeSplitter,
tagdigger_fun.py:1286-1368), then the per-sample files are counted with a blank Barcode column and a
marker-list filter (tagdigger_script -k, tagdigger_script.py:71-76,123-133).  The split files must
have the reference's bytes and the count / genotype CSVs must be byte-identical."""

import base64
import hashlib
import os

import pytest

from conftest import load_golden, materialize
from tagdigger_b200 import barcode_splitter_script, tagdigger_script

FLOW = load_golden("flow.json")["cases"]


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(FLOW)))
def test_split_then_count_equals_reference(i, in_tmp):
    case = FLOW[i]
    materialize(case["files"], in_tmp)
    assert barcode_splitter_script.main(case["split_argv"]) == 0
    for name, digest in case["split_sha256"].items():
        with open(name, "rb") as fh:
            assert hashlib.sha256(fh.read()).hexdigest() == digest, name
    assert tagdigger_script.main(case["count_argv"]) == 0
    for name, b64 in case["outfiles"].items():
        with open(name, "rb") as fh:
            assert fh.read() == base64.b64decode(b64), name


def test_flow_fixture_is_a_two_stage_run():
    case = FLOW[0]
    assert "-k" in case["count_argv"] and "-a" in case["split_argv"]
    key = case["files"]["count_key.csv"]["text"].splitlines()
    assert key[0] == "File,Barcode,Sample" and all(line.split(",")[1] == "" for line in key[1:])     # blank barcodes
    assert len(case["split_sha256"]) == len(key) - 1
    counts = base64.b64decode(case["outfiles"]["counts.csv"]).decode().split("\r\n")
    assert len(counts[0].split(",")) - 1 == 60              # -k kept every second marker: 30 markers x 2 alleles
    assert sum(int(x) for row in counts[1:] if row for x in row.split(",")[1:]) > 500
    assert os.path.basename(case["count_argv"][-3]) == "counts.csv"
