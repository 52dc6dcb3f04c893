"""OpticalFlowCalculationConfig: superset of the reference dataclass (optical_flow/config.py:174-188).

The reference exposes one TV-L1 knob, ``lambda_value`` (passed to ``setLambda`` at
calculate_optical_flow.py:578) and leaves every other DualTVL1 parameter at OpenCV's default.  The extra
fields below are those defaults, so ``OpticalFlowCalculationConfig()`` reproduces the
reference, and BASELINE config 5 (7 scales, 10 warps) can be expressed.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional


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
