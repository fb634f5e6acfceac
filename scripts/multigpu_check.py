#!/usr/bin/env python3
"""Run on a box with >= 2 GPUs: the golden 'lanes' script case under torchrun
(files dealt to the ranks, one NCCL all-reduce) must give the byte-identical CSVs.

    python scripts/multigpu_check.py 2
"""
import base64
import json
import os
import subprocess
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, REPO)
from conftest import file_bytes  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    gold = json.load(open(os.path.join(REPO, "tests", "golden", "script.json")))
    bad = 0
    for case in gold["cases"]:
        if case["returncode"] != 0:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            for name, spec in gold["filesets"][case["fileset"]].items():
                with open(os.path.join(tmp, name), "wb") as fh:
                    fh.write(file_bytes(spec))
            env = dict(os.environ, PYTHONPATH=REPO)
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                   "--master-addr", "127.0.0.1", "--master-port", "29533", "-m", "tagdigger_b200.tagdigger_script"] + case["argv"]
            p = subprocess.run(cmd, cwd=tmp, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
            ok = p.returncode == 0
            for name, b64 in case["outfiles"].items():
                path = os.path.join(tmp, name)
                ok = ok and os.path.exists(path) and open(path, "rb").read() == base64.b64decode(b64)
            print("world=%d %s %s" % (n, " ".join(case["argv"]), "OK" if ok else "MISMATCH"))
            if not ok:
                bad += 1
                print(p.stdout[-3000:])
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
