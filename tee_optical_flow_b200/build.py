"""Builds libteeflow.so (the CUDA engine + C ABI) in-tree with nvcc for sm_100a.

    python -m tee_optical_flow_b200.build

-fmad=false: no FMA contraction, so every float op rounds exactly like the CPU reference (OpenCV's baseline
x86-64 build / oracle/tvl1_oracle.c compiled with -ffp-contract=off).  nvcc's defaults -prec-div=true and
-prec-sqrt=true keep division and square root IEEE-correct.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libteeflow.so"
# second build of the same translation unit with the TMA-staged inner iteration (-DTEEFLOW_TMA_INNER=1): bit-identical
# results, slower; kept building and parity-tested (tests/test_tma_variant_gpu.py, TEEFLOW_LIB selects it)
LIB_TMA = PKG / "libteeflow_tma.so"
VARIANTS = {None: (LIB, []), "tma": (LIB_TMA, ["-DTEEFLOW_TMA_INNER=1"])}
SOURCES = [CSRC / "teeflow.cu"]
# every header the translation unit can include: a stale library after an edit would go unnoticed by the tests
HEADERS = sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.inc")) + sorted(CSRC.glob("*.h")) + \
    [PKG.parent / "include" / "teeflow.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-shared",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set $NVCC)")


def needs_build(variant=None) -> bool:
    lib = VARIANTS[variant][0]
    if not lib.exists():
        return True
    t = lib.stat().st_mtime
    return any(p.exists() and p.stat().st_mtime > t for p in SOURCES + HEADERS + [Path(__file__)])


def build_library(force: bool = False, verbose: bool = False, variant=None) -> Path:
    lib, defines = VARIANTS[variant]
    if not force and not needs_build(variant):
        return lib
    # TEEFLOW_NVCC_EXTRA: extra flags for tuning experiments (e.g. "-DTEEFLOW_MIN_CTAS=3 -DTEEFLOW_PF_ROWS=6")
    extra = os.environ.get("TEEFLOW_NVCC_EXTRA", "").split()
    cmd = [find_nvcc(), *NVCC_FLAGS, *defines, *extra, "-o", str(lib), *map(str, SOURCES)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), file=sys.stderr)
    env = dict(os.environ)
    env.pop("CC", None)   # this image exports CC=/opt/gcc/bin/gcc, which nvcc's host pass does not need
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if verbose or res.returncode != 0:
        print(res.stdout, file=sys.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed ({res.returncode}):\n{res.stdout[-4000:]}")
    return lib


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
    if "--all" in sys.argv:
        print(build_library(force="--force" in sys.argv, verbose=True, variant="tma"))
