"""Runs the hypot self-test (teeflow_selftest_hypot) for the three operand modes and prints mismatches / rejects."""
import ctypes as C
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tee_optical_flow_b200.engine import TVL1Engine

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 30)
with TVL1Engine(device=0) as eng:
    for mode in (0, 1, 2):
        bad, rej = C.c_int64(-1), C.c_int64(-1)
        rc = eng._lib.teeflow_selftest_hypot(eng._h, mode, n, 99 + mode, C.byref(bad), C.byref(rej))
        print(f"mode {mode}: rc {rc} pairs {n} mismatches {bad.value} rejected {rej.value} ({rej.value / n:.2e})", flush=True)
