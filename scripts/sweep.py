#!/usr/bin/env python3
"""Build tuning variants of the library (geometry macros of csrc/tdg_kernel.cuh) and time them.

    python scripts/sweep.py build            # here (no GPU): writes build/variants/*.so
    python scripts/sweep.py run [reads]      # on the GPU box: bench each variant, print kernel ms
"""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "gpurun_variants")
VARIANTS = {
    "w14_c15": ["-DTDG_WARPS=14", "-DTDG_CHUNKS=15"],      # the library's geometry
    "w13_c15": ["-DTDG_WARPS=13", "-DTDG_CHUNKS=15"],
    "w12_c15": ["-DTDG_WARPS=12", "-DTDG_CHUNKS=15"],      # 3 warps per scheduler: up to 168 registers
    "w15_c13": ["-DTDG_WARPS=15", "-DTDG_CHUNKS=13"],
    "w16_c13": ["-DTDG_WARPS=16", "-DTDG_CHUNKS=13"],
}


def build():
    sys.path.insert(0, REPO)
    from tagdigger_b200 import _native
    os.makedirs(OUT, exist_ok=True)
    for name, flags in VARIANTS.items():
        lib = os.path.join(OUT, "lib_%s.so" % name)
        cmd = ["nvcc"] + _native.NVCC_FLAGS + flags + ["-Xptxas", "-v", "-o", lib,
                                                       os.path.join(REPO, "tagdigger_b200", "csrc", "tdg_api.cu"), "-lz"]
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
        regs = [ln for ln in p.stdout.splitlines() if "Used" in ln or "spill" in ln]
        print(name, "rc", p.returncode, "|", " ; ".join(r.strip() for r in regs[2:4]))
        if p.returncode:
            print(p.stdout[-2000:])


def run():
    reads = sys.argv[2] if len(sys.argv) > 2 else "100000000"
    for name in VARIANTS:
        lib = os.path.join(OUT, "lib_%s.so" % name)
        if not os.path.exists(lib):
            continue
        env = dict(os.environ, TDG_LIB=lib)
        p = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--reads", reads, "--steps", "5", "--warmup", "3",
                            "--no-e2e", "--no-cpu"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, universal_newlines=True)
        try:
            j = json.loads(p.stdout.strip().splitlines()[-1])
            print("%-14s ms/step %8.3f  frac %.4f  check %s" % (name, j["ms_per_step"], j["roofline"]["frac"], j["check"]), flush=True)
        except (ValueError, IndexError):
            print(name, "FAILED", p.returncode, p.stderr[-500:], flush=True)


if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]]()
