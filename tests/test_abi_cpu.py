"""CPU tests of the C-ABI surface and host logic (no compute calls without a GPU)."""
import ctypes as C
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def built_lib():
    from tee_optical_flow_b200.build import build_library
    return build_library()


def _declared_symbols():
    text = (ROOT / "include" / "teeflow.h").read_text()
    return sorted(set(re.findall(r"TEEFLOW_API\s+[\w\s\*]+?\b(teeflow_\w+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = _declared_symbols()
    for name in ["teeflow_create", "teeflow_destroy", "teeflow_set_param", "teeflow_calc_clip", "teeflow_calc_pairs",
                 "teeflow_calc_clip_host", "teeflow_calc_pair_host", "teeflow_get_counters", "teeflow_get_stats",
                 "teeflow_last_error"]:
        assert name in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = C.CDLL(str(built_lib))
    for name in _declared_symbols():
        assert hasattr(lib, name), f"libteeflow.so does not export {name}"


def test_ctypes_table_matches_header(built_lib):
    from tee_optical_flow_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    lib = _lib.load()
    assert lib.teeflow_abi_version() == 1


def test_default_params_are_opencv_defaults(built_lib):
    from tee_optical_flow_b200 import _lib
    p = _lib.TeeflowParams()
    _lib.load().teeflow_default_params(C.byref(p))
    assert (p.tau, p.lambda_, p.theta, p.epsilon, p.scale_step) == (0.25, 0.15, 0.3, 0.01, 0.8)
    assert (p.nscales, p.warps, p.inner_iterations, p.outer_iterations, p.median_filtering) == (5, 5, 30, 10, 5)


def test_no_gpu_is_a_loud_error(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tee_optical_flow_b200.engine import TVL1Engine
    from tee_optical_flow_b200.exceptions import EngineUnavailableError, OpticalFlowCalculationError
    with pytest.raises(EngineUnavailableError):
        TVL1Engine()
    assert issubclass(EngineUnavailableError, OpticalFlowCalculationError)


def test_bad_params_rejected_before_touching_the_gpu(built_lib):
    from tee_optical_flow_b200 import _lib
    lib = _lib.load()
    p = _lib.TeeflowParams()
    lib.teeflow_default_params(C.byref(p))
    p.median_filtering = 7
    h = C.c_void_p()
    assert lib.teeflow_create(C.byref(p), 0, C.byref(h)) == _lib.ERR_BAD_ARG
    assert b"median_filtering" in lib.teeflow_last_error(None)
    assert not h.value


def test_product_never_imports_the_oracle():
    """the product package must not reference oracle/ (a CPU fallback would void the parity claim)"""
    for f in (ROOT / "tee_optical_flow_b200").rglob("*"):
        if f.suffix in {".py", ".cu", ".cuh", ".h"}:
            text = f.read_text()
            for needle in ("import oracle", "from oracle", "libtvl1_oracle", "oracle.tvl1", "oracle/_ref"):
                assert needle not in text, (f, needle)


def test_config_superset_of_reference():
    from tee_optical_flow_b200.config import OpticalFlowCalculationConfig, default_optical_flow_config
    c = default_optical_flow_config()
    # the reference's fields and defaults (optical_flow/config.py:174-188)
    assert (c.lambda_value, c.moving_avg_window, c.moving_avg_threshold, c.min_mask_size) == (0.15, 4, 0.49, 500)
    assert (c.ecg_sampling_rate, c.art_sampling_rate, c.cvp_sampling_rate, c.pap_sampling_rate) == (500, 125, 125, 125)
    # OpenCV defaults for the knobs the reference leaves untouched
    p = c.tvl1_params()
    assert p == dict(tau=0.25, lambda_=0.15, theta=0.3, nscales=5, warps=5, epsilon=0.01, inner_iterations=30,
                     outer_iterations=10, scale_step=0.8, median_filtering=5, max_slots=0)
    assert OpticalFlowCalculationConfig(iterations=12).tvl1_params()["inner_iterations"] == 12


def test_synth_is_deterministic():
    from tee_optical_flow_b200.synth import make_clip, make_masks
    a = make_clip(seed=3, n_frames=3, H=48, W=64)
    b = make_clip(seed=3, n_frames=3, H=48, W=64)
    assert a.dtype == np.uint8 and np.array_equal(a, b)
    assert not np.array_equal(a, make_clip(seed=4, n_frames=3, H=48, W=64))
    m = make_masks(3, 3, 48, 64)
    assert set(m) == {"rv", "av", "bkgd"} and m["rv"].shape == (3, 48, 64, 2) and m["rv"].dtype == bool
    assert not np.any(m["bkgd"] & (m["rv"] | m["av"]))


def test_median_networks_are_verified_and_current():
    """the compare-exchange networks of the 5x5 median (sorted rows, shared middle six) select the median of every
    zero-one window (2^25 cases => every input, 0-1 principle) and median_networks.inc is what the generator emits"""
    import subprocess
    import sys
    res = subprocess.run([sys.executable, str(ROOT / "tools" / "gen_median_networks.py"), "--check"],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "wrong medians over 2^25 zero-one windows: 0" in res.stdout


def test_tma_variant_builds_exports_the_abi_and_carries_tma_instructions(built_lib):
    """libteeflow_tma.so (same source, -DTEEFLOW_TMA_INNER=1) builds, exports every declared symbol, and its SASS holds
    the TMA tensor loads (UTMALDG) and mbarrier waits the staged inner iteration uses; the shipped default has none."""
    import shutil
    import subprocess
    from tee_optical_flow_b200.build import build_library
    tma = build_library(variant="tma")
    lib = C.CDLL(str(tma))
    for name in _declared_symbols():
        assert hasattr(lib, name), f"libteeflow_tma.so does not export {name}"
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not shutil.which(cuobjdump):
        pytest.skip("cuobjdump not available")
    count = lambda path, pat: sum(pat in l for l in subprocess.run([cuobjdump, "-sass", str(path)], capture_output=True, text=True).stdout.splitlines())
    assert count(tma, "UTMALDG.4D") > 0 and count(tma, "SYNCS.PHASECHK") > 0
    assert count(built_lib, "UTMALDG") == 0
