"""Host text checks of tdg_count_file (tagdigger_b200/csrc/tdg_text.h), CPU only: the UTF-8
validator accepts exactly what Python's strict decoder accepts (the reference reads its files in
text mode, tagdigger_fun.py:240-243), and the line-limit finder cuts where Python's universal
newlines put the n-th line end (the maxreads stop of tagdigger_fun.py:272-273)."""

import io
import random

import pytest

from feed_check import line_limit, utf8_first_invalid


def _py_first_invalid(data):
    try:
        data.decode("utf-8")
        return -1
    except UnicodeDecodeError as e:
        return e.start


CASES = [
    b"", b"plain ascii\n", u"grüße € \U0001F9EC ퟿ ".encode("utf-8"),
    b"\x80", b"abc\xffdef", b"\xc0\xaf", b"\xc1\xbf", b"\xc2", b"\xc2A", b"\xe0\x80\x80", b"\xe0\x9f\xbf", b"\xe0\xa0\x80",
    b"\xed\xa0\x80", b"\xed\x9f\xbf", b"\xef\xbf\xbf", b"\xf0\x80\x80\x80", b"\xf0\x8f\xbf\xbf", b"\xf0\x90\x80\x80",
    b"\xf4\x8f\xbf\xbf", b"\xf4\x90\x80\x80", b"\xf5\x80\x80\x80", b"ok\xe2\x82", b"ok\xe2\x82\xac", b"ok\xe2\x82\xacA\xe2A",
    b"\xf1\x80\x80", b"\xf1\x80\x80A", b"A" * 100 + b"\xe2\x28\xa1", b"\xfe", b"\xf8\x88\x80\x80\x80",
]


@pytest.mark.parametrize("piece", [1, 2, 3, 7, 1 << 16])
def test_utf8_validator_equals_python(piece):
    for data in CASES:
        assert utf8_first_invalid(data, piece) == _py_first_invalid(data), (data, piece)


def test_utf8_validator_random():
    r = random.Random(5)
    alphabet = ["A", "\n", u"é", u"€", u"\U0001F600", ""]
    for trial in range(400):
        text = "".join(r.choice(alphabet) for _ in range(r.randint(0, 60))).encode("utf-8")
        data = bytearray(text)
        for _ in range(r.randint(0, 2)):
            if data:
                data[r.randrange(len(data))] = r.choice([0x80, 0xBF, 0xC0, 0xE0, 0xED, 0xF0, 0xF4, 0xFF, 0x41])
        data = bytes(data)
        for piece in (1, 5, 1 << 16):
            assert utf8_first_invalid(data, piece) == _py_first_invalid(data), (data, piece)


def _py_offset_of_line_end(data, lines):
    """Bytes Python's text layer has consumed after `lines` complete lines (universal newlines);
    len(data) + 1 when the data holds fewer terminated lines."""
    raw = io.TextIOWrapper(io.BytesIO(data), encoding="latin-1", newline="")   # keeps the terminators
    used = 0
    for _ in range(lines):
        line = raw.readline()
        if not line or not (line.endswith("\n") or line.endswith("\r")):
            return len(data) + 1
        used += len(line)
    return used


@pytest.mark.parametrize("piece", [1, 2, 5, 64, 1 << 16])
def test_line_limit_equals_python(piece):
    r = random.Random(9)
    for trial in range(150):
        parts = []
        for _ in range(r.randint(0, 12)):
            parts.append("".join(r.choice("ACGT@+I") for _ in range(r.randint(0, 9))))
            parts.append(r.choice(["\n", "\r\n", "\r", "\n", "\r\r", "\n\r"]))
        if r.random() < 0.5:
            parts.append("tail")
        data = "".join(parts).encode()
        total = len(io.TextIOWrapper(io.BytesIO(data), encoding="latin-1", newline=None).readlines())
        for lines in range(1, total + 2):
            want = _py_offset_of_line_end(data, lines)
            got = line_limit(data, lines, piece)
            # a '\r' that is the very last byte: Python ends the line at end of file, the feed leaves it
            # pending (nothing follows that could need it)
            if want == len(data) and data.endswith(b"\r"):
                assert got in (want, len(data) + 1), (data, lines, piece)
            else:
                assert got == want, (data, lines, piece)
