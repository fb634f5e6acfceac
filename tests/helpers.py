"""Input builders shared by the CPU and GPU tests."""

import random

import numpy as np


def rand_seq(r, n, alphabet="ACGT"):
    return "".join(r.choice(alphabet) for _ in range(n))


def small_setup(r, cutsite="TGCAG", nbar=6, ntag=12, tag_len=(10, 40)):
    """Prefix-free barcodes (with the cut site) and tags that start with it."""
    barcodes = []
    while len(barcodes) < nbar:
        b = rand_seq(r, r.randint(3, 8))
        pat = b + cutsite
        if any(p.startswith(pat) or pat.startswith(p) for p in (x + cutsite for x in barcodes)):
            continue
        barcodes.append(b)
    tags = []
    while len(tags) < ntag:
        t = cutsite + rand_seq(r, r.randint(*tag_len))
        if any(u.startswith(t) or t.startswith(u) for u in tags):
            continue
        tags.append(t)
    return barcodes, tags


def line_soup(r, barcodes, tags, cutsite, nlines, newline_kinds=("\n",), weird=0.15,
              final_newline=True, max_pad=4):
    """A text image whose lines are a mix of matching reads, near misses,
    FASTQ-looking headers/quality strings, empty lines and whitespace, joined
    by line ends drawn from ``newline_kinds``.  Returns bytes (ASCII)."""
    out = []
    for i in range(nlines):
        k = r.random()
        if k < 0.45:
            t = r.choice(tags)
            line = r.choice(barcodes) + t + rand_seq(r, r.randint(0, 30))
        elif k < 0.55:
            line = r.choice(barcodes) + cutsite + rand_seq(r, r.randint(0, 50))
        elif k < 0.65:
            line = "@" + rand_seq(r, r.randint(0, 40), "ABCDEFGHIJ0123456789:")
        elif k < 0.75:
            line = "+"
        elif k < 0.85:
            line = rand_seq(r, r.randint(0, 90), "IJKLMNOP@+#$%&")
        elif k < 0.9:
            line = ""
        else:
            line = rand_seq(r, r.randint(0, 120))
        if r.random() < weird:
            w = r.random()
            if w < 0.25:
                line = line.lower()
            elif w < 0.5:
                pad = "".join(r.choice(" \t\x0b\x0c\x1c\x1d\x1e\x1f") for _ in range(r.randint(1, max_pad)))
                line = pad + line + (pad if r.random() < 0.5 else "")
            elif w < 0.75 and line:
                j = r.randrange(len(line))
                line = line[:j] + r.choice("N. \tx") + line[j + 1:]
            else:
                line = line[:r.randint(0, len(line))]
        out.append(line)
        if i < nlines - 1 or final_newline:
            out.append(r.choice(newline_kinds))
    return "".join(out).encode("ascii")


def fastq_of(reads, newline=b"\n", r=None):
    """Minimal 4-line records around the given sequence lines (bytes)."""
    out = []
    for i, s in enumerate(reads):
        if isinstance(s, str):
            s = s.encode()
        q = b"I" * len(s)
        out.append(b"@r%d" % i + newline + s + newline + b"+" + newline + q + newline)
    return b"".join(out)
