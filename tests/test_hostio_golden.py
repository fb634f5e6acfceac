"""The host API mirror (tagdigger_b200/hostio.py) against outputs recorded from the
reference (tests/golden/readers.json, small_functions.json): return values, printed
messages, exception types and files written.  CPU only."""

import base64
import contextlib
import io
import os

import pytest

from conftest import load_golden, materialize
from tagdigger_b200 import hostio, matchset

READERS = load_golden("readers.json")
SMALL = load_golden("small_functions.json")


def _jsonable(x):
    if isinstance(x, (list, tuple)):
        return [_jsonable(v) for v in x]
    if isinstance(x, dict):
        return {str(k): _jsonable(v) for k, v in x.items()}
    if isinstance(x, (set, frozenset)):
        return sorted(_jsonable(v) for v in x)
    return x


def _run(case, fn):
    out = io.StringIO()
    ret, exc = None, None
    with contextlib.redirect_stdout(out):
        try:
            ret = fn(*case["args"], **case["kwargs"])
        except BaseException as e:  # noqa: BLE001
            exc = [type(e).__name__, str(e)]
    return _jsonable(ret), exc, out.getvalue()


def _check(case, fn):
    ret, exc, stdout = _run(case, fn)
    assert (exc[0] if exc else None) == (case["exc"][0] if case["exc"] else None), (exc, case["exc"])
    if exc and exc[0] in ("AssertionError", "Exception"):
        assert exc[1] == case["exc"][1]
    assert ret == case["ret"]
    assert stdout == case["stdout"]
    for name, b64 in case["outfiles"].items():
        if b64 is None:
            assert not os.path.exists(name)
        else:
            with open(name, "rb") as fh:
                assert fh.read() == base64.b64decode(b64), name


@pytest.mark.parametrize("i", range(len(READERS)))
def test_readers_golden(i, in_tmp):
    case = READERS[i]
    materialize(case["files"], in_tmp)
    _check(case, getattr(hostio, case["func"]))


@pytest.mark.parametrize("i", range(len(SMALL)))
def test_small_functions_golden(i, in_tmp):
    case = SMALL[i]
    materialize(case["files"], in_tmp)
    fn = matchset.enumerate_cut_sites if case["func"] == "enumerate_cut_sites" else getattr(hostio, case["func"])
    _check(case, fn)
