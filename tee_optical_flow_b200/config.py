"""OpticalFlowCalculationConfig: superset of the reference dataclass (optical_flow/config.py:174-188).

The reference exposes one TV-L1 knob, ``lambda_value`` (passed to ``setLambda`` at
calculate_optical_flow.py:578) and leaves every other DualTVL1 parameter at OpenCV's default.  The extra
fields below are those defaults, so ``OpticalFlowCalculationConfig()`` reproduces the
reference, and BASELINE config 5 (7 scales, 10 warps) can be expressed.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional


@dataclass
class OpticalFlowCalculationConfig:
    """Configuration for optical flow calculation and processing."""
    lambda_value: float = 0.15
    moving_avg_window: int = 4
    moving_avg_threshold: float = 0.49
    min_mask_size: int = 500
    waveform_flatness_threshold: float = 0.05
    pap_max_mean: float = 100.0
    cvp_max_mean: float = 50.0
    cvp_min_mean: float = -10.0
    ecg_sampling_rate: int = 500
    art_sampling_rate: int = 125
    cvp_sampling_rate: int = 125
    pap_sampling_rate: int = 125
    # --- DualTVL1 parameters the reference leaves at OpenCV's defaults
    tau: float = 0.25
    theta: float = 0.3
    nscales: int = 5
    warps: int = 5
    epsilon: float = 0.01
    inner_iterations: int = 30
    outer_iterations: int = 10
    iterations: Optional[int] = None   # convenience alias: sets inner_iterations when given
    scale_step: float = 0.8
    median_filtering: int = 5
    # --- engine knobs (no reference counterpart)
    max_slots: int = 0                 # frame pairs solved concurrently on one GPU (0 = engine default)

    def tvl1_params(self) -> dict:
        return dict(tau=self.tau, lambda_=self.lambda_value, theta=self.theta, nscales=self.nscales,
                    warps=self.warps, epsilon=self.epsilon,
                    inner_iterations=self.iterations if self.iterations is not None else self.inner_iterations,
                    outer_iterations=self.outer_iterations, scale_step=self.scale_step,
                    median_filtering=self.median_filtering, max_slots=self.max_slots)


def default_optical_flow_config() -> OpticalFlowCalculationConfig:
    """Create default optical flow calculation configuration (config.py:191)."""
    return OpticalFlowCalculationConfig()


# ---- the downstream configs the waveform pipeline (waveforms.py) reads; field names and defaults of the reference
@dataclass
class CardiacCycleConfig:
    """optical_flow/config.py:12-29.  Only smooth_fraction / pad_len act on the angle-based detector of this path;
    the remaining fields belong to the ECG / arterial / area detectors (out of scope) and keep the schema intact."""
    smooth_fraction: float = 0.2
    pad_len: int = 20
    sys_thres: float = 0.9
    dia_thres: float = 0.5
    rr_sys_ratio: float = 0.333
    sys_extension: int = 2
    t_peak_thres: float = 0.5
    t_min_dist: int = 20
    rr_search_range: List[float] = field(default_factory=lambda: [0.2, 0.75])
    low_peak_thres: float = 0.9
    low_min_dist: int = 50
    high_peak_thres: float = 0.9
    high_min_dist: int = 50
    sys_upstroke_multiplier: int = 2
    sys_upstroke_offset: int = 5


@dataclass
class PeakDetectionConfig:
    """optical_flow/config.py:74-82."""
    peak_thres: float = 0.2
    min_dist: int = 5
    pick_peak_by_subset: bool = True
    show_all_peaks: bool = False
    smooth_fraction: float = 0.3
    pad_len: int = 20


@dataclass
class AnalysisConfig:
    """optical_flow/config.py:85-95."""
    percentile: int = 99
    perc_lo: int = 1
    perc_hi: int = 99
    av_filter_flag: bool = True
    av_savgol_window: int = 10
    av_savgol_poly: int = 4
    print_report: bool = False
    return_value: bool = True


def default_cardiac_cycle_config() -> CardiacCycleConfig:
    return CardiacCycleConfig()


def default_peak_detection_config() -> PeakDetectionConfig:
    return PeakDetectionConfig()
