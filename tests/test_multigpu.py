"""CSV identity across GPU counts (needs >= 2 CUDA devices; skipped on a one-GPU box).

Runs scripts/multigpu_check.py at world = every visible GPU: the recorded whole-script cases
under torchrun (files dealt to the ranks, ONE NCCL all-reduce of the count matrix) must give the
reference's CSV bytes, and one plain FASTQ file cut into byte ranges across the ranks (LF and CRLF)
must give the CSV of the single-rank run and the C oracle's counts.  The transcript is kept under
gpurun_out/ so that a run on the 8-GPU box leaves an artifact."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


@pytest.mark.skipif(_device_count() < 2, reason="needs at least two GPUs")
def test_csv_identical_across_gpu_counts():
    n = min(_device_count(), 8)
    p = subprocess.run([sys.executable, os.path.join(REPO, "scripts", "multigpu_check.py"), str(n)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True, timeout=1500)
    out_dir = os.path.join(REPO, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "multigpu_check_world%d.txt" % n), "w") as fh:
            fh.write(p.stdout)
    assert p.returncode == 0, p.stdout[-4000:]
    assert "MISMATCH" not in p.stdout
