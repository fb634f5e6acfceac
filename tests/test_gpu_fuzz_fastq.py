"""Damaged FASTQ through the kernel's fast path against the C oracle (scripts/gpu_fuzz.py holds the
generator; a long run of it -- 1,500+ images -- is part of the round's GPU checks)."""

import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu

_SPEC = importlib.util.spec_from_file_location(
    "gpu_fuzz", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "gpu_fuzz.py"))


@pytest.mark.parametrize("seed0", [0, 1000])
def test_damaged_fastq_images(seed0):
    mod = importlib.util.module_from_spec(_SPEC)
    _SPEC.loader.exec_module(mod)
    assert mod.run(14, seed0) == 0
