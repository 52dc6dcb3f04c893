"""The TMA-staged build of the inner iteration (libteeflow_tma.so, -DTEEFLOW_TMA_INNER=1: cp.async.bulk.tensor boxes
into a per-warp shared-memory ring, mbarrier-signalled) must give the same bits as the shipped register form: the
whole engine parity suite (oracle golden pairs, ragged sizes, 600x800, clips with refill, counters) is re-run against
it in a subprocess (TEEFLOW_LIB selects the library), and one clip is compared between the two libraries."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tma_lib():
    from tee_optical_flow_b200.build import build_library
    return build_library(variant="tma")


def test_engine_parity_suite_on_the_tma_build(tma_lib):
    env = dict(os.environ, TEEFLOW_LIB=str(tma_lib))
    res = subprocess.run([sys.executable, "-m", "pytest", "tests/test_engine_gpu.py", "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider"],
                         cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:]
    assert " passed" in res.stdout


def test_tma_build_equals_default_build_on_a_clip(tma_lib, tmp_path):
    """Same clip through both libraries: flow bits, fp16 flow and per-level iteration counters are equal."""
    script = (
        "import sys, numpy as np\n"
        "from tee_optical_flow_b200.engine import TVL1Engine\n"
        "from tee_optical_flow_b200.synth import make_clip\n"
        "fr = make_clip(seed=5, n_frames=9, H=200, W=333, peak_disp=4.0, period=8.0)\n"
        "with TVL1Engine(device=0) as eng:\n"
        "    f32, f16 = eng.calc_clip(fr, want_f32=True, want_f16=True)\n"
        "    c, info = eng.last_counters()\n"
        "np.savez(sys.argv[1], f32=f32, f16=f16, c=c)\n")
    outs = []
    for name, lib in (("default", None), ("tma", str(tma_lib))):
        env = dict(os.environ)
        env.pop("TEEFLOW_LIB", None)
        if lib:
            env["TEEFLOW_LIB"] = lib
        out = tmp_path / f"{name}.npz"
        res = subprocess.run([sys.executable, "-c", script, str(out)], cwd=ROOT, env=env, stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True, timeout=600)
        assert res.returncode == 0, res.stdout[-2000:]
        outs.append(np.load(out))
    a, b = outs
    assert np.array_equal(a["f32"], b["f32"]) and np.array_equal(a["f16"], b["f16"]) and np.array_equal(a["c"], b["c"])
