"""Exception types of the reference (optical_flow/exceptions.py:7-34), same names and hierarchy, so that code
written against the reference catches the same classes."""


class OpticalFlowError(Exception):
    """Base exception for optical flow processing errors."""


class DICOMReadError(OpticalFlowError):
    """Raised when DICOM file cannot be read."""


class WaveformLoadError(OpticalFlowError):
    """Raised when waveform file cannot be loaded."""


class WaveformValidationError(OpticalFlowError):
    """Raised when waveform validation fails."""


class OpticalFlowCalculationError(OpticalFlowError):
    """Raised when optical flow calculation fails."""


class ConfigurationError(OpticalFlowError):
    """Raised when configuration is invalid."""


class EngineUnavailableError(OpticalFlowCalculationError):
    """libteeflow.so is missing / not loadable, or no CUDA device: there is deliberately no CPU fallback."""


class SaliencyParityWarning(UserWarning):
    """The `no_saliency=False` input stage (StaticSaliencyFineGrained) runs a restatement whose parity with
    cv2.saliency is unpinned; emitted once per process_frames call that uses it."""
