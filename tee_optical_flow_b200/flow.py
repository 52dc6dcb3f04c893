"""The reference's producer API for the TV-L1 path, same names / arguments / error behaviour
(optical_flow/calculate_optical_flow.py), with the per-pair OpenCV call replaced by the batched B200 engine.

    calculate_optical_flow(saliency_1, saliency_2, mask_dict, OF_model, bkgd_comp, OF_algo)   :627-660
    process_frames(frames, ...)      array-level form of process_video (no DICOM / HDF5 libraries needed)
    process_video(dcm_path, save_path, segmentor_model, ...)                                  :478-625
    process_folder(dcm_folder, save_folder, segmentor_model, nchunks, chunk_index, ...)       :243-290

Out of scope here: SAM mask prediction, DICOM parsing and waveform loading are inputs to this
path; process_video imports pydicom / h5py lazily and raises a clear error when they are missing.
"""
from __future__ import annotations

import logging
import os
import traceback
import warnings
from typing import Any, Dict, List, Optional

import numpy as np

from .config import OpticalFlowCalculationConfig, default_optical_flow_config
from .engine import TVL1Engine
from .exceptions import ConfigurationError, DICOMReadError, OpticalFlowCalculationError, SaliencyParityWarning

logger = logging.getLogger(__name__)


# ------------------------------------------------------------------------------------------------ container content
def rgb2gray(rgb: np.ndarray) -> np.ndarray:
    """skimage.color.rgb2gray for the 'echo' dataset of the container (:399-401): float image in [0,1], luminance
    0.2125 R + 0.7154 G + 0.0721 B.  Host-side like the reference; the SOLVER's input stage (img2uint8(rgb2gray(.)),
    :588) never runs here -- it is teeflow_prepare_frames on the GPU, checked against oracle/frame_prep_ref.py."""
    a = np.asarray(rgb)
    a = a.astype(np.float64) / 255.0 if a.dtype == np.uint8 else a.astype(np.float64)
    return (a[..., 0] * 0.2125 + a[..., 1] * 0.7154) + a[..., 2] * 0.0721


def gray2rgb(nparr: np.ndarray) -> np.ndarray:
    """skimage.color.gray2rgb (:536): replicate a (N,H,W) clip into three identical channels (data movement only)."""
    return np.stack([nparr] * 3, axis=-1)


# ------------------------------------------------------------------------------------------------ per pair
def calculate_optical_flow(saliency_1: np.ndarray, saliency_2: np.ndarray, mask_dict: Dict[str, np.ndarray],
                           OF_model: Any, bkgd_comp: str = 'none', OF_algo: str = 'TVL1') -> Optional[np.ndarray]:
    """Same contract as the reference (:627-660).  `OF_model` is a TVL1Engine (or anything with .calc)."""
    if OF_algo == 'TVL1':
        if bkgd_comp not in ('WASE', 'none'):
            error_msg = f'bkgd_comp value must be [WASE, none], got {bkgd_comp}!'
            logger.error(error_msg)
            raise OpticalFlowCalculationError(error_msg)
        if isinstance(OF_model, TVL1Engine):
            # the weight map sum_n bkgd[n] is cached per mask object: the reference's loop passes the same
            # mask_dict for every pair of a clip (:594), so only the first pair uploads and reduces it
            OF_model.set_wase_masks(mask_dict['bkgd'] if bkgd_comp == 'WASE' else None, cache=True)
            try:
                return OF_model.calc(saliency_1, saliency_2, None)      # background already subtracted on the GPU
            finally:
                OF_model.set_wase_masks(None, cache=True)
        flow = OF_model.calc(saliency_1, saliency_2, None)
    elif OF_algo == 'deepflow':
        raise OpticalFlowCalculationError("OF_algo='deepflow' is outside this engine (TV-L1 path only)")
    else:
        error_msg = 'OF_algo only supports deepflow or TVL1'
        logger.error(error_msg)
        raise OpticalFlowCalculationError(error_msg)
    background = 0
    if bkgd_comp == 'WASE':
        masked_flow = flow * mask_dict['bkgd']
        background = np.mean(masked_flow[masked_flow != 0])
    return flow - background


# ------------------------------------------------------------------------------------------------ per clip
def process_frames(frames: np.ndarray, mask_dict: Optional[Dict[str, np.ndarray]] = None,
                   pixel_spacing: Optional[float] = None, frame_rate: Optional[float] = None,
                   mode: str = 'RVIO_2class', bkgd_comp: str = 'none', no_saliency: bool = True,
                   OF_algo: str = 'TVL1', save_mask_subset: Optional[List[str]] = None,
                   config: Optional[OpticalFlowCalculationConfig] = None, engine: Optional[TVL1Engine] = None,
                   frames_are_prepared: bool = False, patient_id: str = '', heart_rate: float = 0) -> Dict[str, Any]:
    """Array-level process_video: frames (N,H,W[,3]) -> the HDF5 layout of _save_optical_flow_to_hdf5 (:370-475)
    as an in-memory dict: datasets 'echo' (N,H,W) f16, 'flow' (N,H,W,2) f16, one bool dataset per mask label, and
    'attrs' (the attributes of the 'flow' dataset that OpticalFlowDataset reads, optical_flow_dataset.py:45-111).
    """
    if config is None:
        config = default_optical_flow_config()
    if mode == 'otsu':
        if bkgd_comp != 'none':
            raise ConfigurationError(f'bkgd_comp {bkgd_comp} is not supported in mode=otsu, can only support bkgd_comp=none')
        if save_mask_subset is not None:
            raise ConfigurationError('In mode=otsu, save_mask_subset must be None')
    elif mode not in ('A4C', 'RVIO_2class'):
        raise ConfigurationError(f'Input for mode must be [A4C, otsu, RVIO_2class], not {mode}.')
    if OF_algo != 'TVL1':
        raise OpticalFlowCalculationError('OF_algo only supports deepflow or TVL1' if OF_algo != 'deepflow'
                                          else "OF_algo='deepflow' is outside this engine (TV-L1 path only)")
    if bkgd_comp not in ('WASE', 'none'):
        raise OpticalFlowCalculationError(f'bkgd_comp value must be [WASE, none], got {bkgd_comp}!')
    frames = np.asarray(frames)
    if not no_saliency and not (frames.dtype == np.uint8 and frames.ndim == 4 and frames.shape[-1] == 3):
        raise OpticalFlowCalculationError('no_saliency=False needs (N,H,W,3) uint8 frames (computeSaliency input, :586)')
    if no_saliency and not frames_are_prepared and not (
            frames.dtype == np.uint8 and (frames.ndim == 3 or (frames.ndim == 4 and frames.shape[-1] == 3))):
        # no CPU fallback: DICOM pixel data reaches this path as uint8 (N,H,W[,3]); anything else is refused
        raise OpticalFlowCalculationError(
            f'frames must be uint8 (N,H,W) or (N,H,W,3), got {frames.dtype} {frames.shape}; '
            'pass frames_are_prepared=True for (N,H,W) uint8 images that are already the solver input')
    if frames.shape[0] < 2:
        raise OpticalFlowCalculationError('need at least two frames')
    mask_dict = mask_dict or {}
    if bkgd_comp == 'WASE' and 'bkgd' not in mask_dict:
        raise ConfigurationError("bkgd_comp='WASE' needs mask_dict['bkgd']")

    conversion_factor = 1.0 if (pixel_spacing is None or frame_rate is None) else pixel_spacing * frame_rate   # :538-541

    own = engine is None
    if own:
        engine = TVL1Engine(**config.tvl1_params())
    saliency_parity = 'not used'
    try:
        if not no_saliency:
            # saliency_obj.computeSaliency(frame) per frame (:586): float32 maps in [0,1] are the solver's images.
            # PARITY UNPINNED: cv2.saliency is in neither this image nor /root/reference; the GPU stage equals a
            # restatement from memory (oracle/saliency_ref.py), not a genuine cv2.saliency output.  Say so loudly.
            warnings.warn("no_saliency=False: the StaticSaliencyFineGrained stage is a restatement whose parity with "
                          "cv2.saliency is UNPINNED (see DESIGN.md); use no_saliency=True for the pinned input stage, "
                          "or pin it with tools/dump_golden.py on a machine with opencv-contrib",
                          SaliencyParityWarning, stacklevel=2)
            gray_u8 = engine.compute_saliency(frames)
            saliency_parity = 'unpinned'
        elif frames_are_prepared:
            gray_u8 = frames
        elif frames.dtype == np.uint8 and frames.ndim == 4 and frames.shape[-1] == 3:
            gray_u8 = engine.prepare_frames(frames)          # img2uint8(rgb2gray(.)) on the GPU (:588)
        elif frames.dtype == np.uint8 and frames.ndim == 3:
            gray_u8 = engine.prepare_frames(gray2rgb(frames))   # greyscale DICOM: gray2rgb first (:532-536)
        else:   # unreachable: refused before the engine was created
            raise OpticalFlowCalculationError('unsupported frame array')
        if gray_u8.ndim != 3 or gray_u8.dtype != (np.uint8 if no_saliency else np.float32):
            raise OpticalFlowCalculationError('prepared frames must be (N,H,W) uint8')
        engine.set_wase_masks(mask_dict['bkgd'] if bkgd_comp == 'WASE' else None)
        # pair loop (:584-597) + copy of the last flow (:599) + * conversion_factor (:600) + astype(float16) (:403)
        _, flow16 = engine.calc_clip(gray_u8, out_scale=conversion_factor, duplicate_last=True, want_f32=False,
                                     want_f16=True)
        counters, info = engine.last_counters()
    finally:
        engine.set_wase_masks(None)
        if own:
            engine.close()

    echo = (rgb2gray(frames) if frames.ndim == 4 else frames.astype(np.float64) / (255.0 if frames.dtype == np.uint8 else 1.0))
    saved = [k for k in mask_dict if save_mask_subset is None or k in save_mask_subset]
    out: Dict[str, Any] = {'echo': echo.astype(np.float16), 'flow': flow16}
    for k in saved:
        out[k] = np.asarray(mask_dict[k])
    out['attrs'] = {
        'frame_rate': frame_rate, 'nframes': int(frames.shape[0]), 'pixel_spacing': pixel_spacing, 'ID': patient_id,
        'HR': heart_rate, 'no_saliency': no_saliency, 'mode': mode,
        'units_converted': (pixel_spacing is not None and frame_rate is not None), 'waveforms_present': False,
        'labels': saved,
    }
    out['_engine_info'] = dict(info, saliency_parity=saliency_parity)
    out['_counters'] = counters
    return out


def save_hdf5(save_path: str, result: Dict[str, Any]) -> None:
    """Writes process_frames' dict as the reference's container (:399-472): datasets 'echo', 'flow', one per saved mask
    label (+ 'RWaveTime' when present), chunked + gzip-9, the attributes on 'flow'.  Uses the package's own HDF5 writer
    (hdf5.py): h5py / libhdf5 are not needed."""
    from .hdf5 import write_hdf5
    if os.path.exists(save_path):
        os.remove(save_path)
    data = {'echo': result['echo'], 'flow': result['flow']}
    for k in result['attrs']['labels']:
        data[k] = np.asarray(result[k])
    if 'RWaveTime' in result:
        data['RWaveTime'] = np.asarray(result['RWaveTime'], dtype=np.float64)
    write_hdf5(save_path, data, {'flow': dict(result['attrs'])})


def extract_dicom_metadata(ds: Any, verbose: bool = False) -> Dict[str, Any]:
    """_extract_dicom_metadata of the reference (:315-367): every item is read independently, so a missing tag only
    blanks its own entry.  pixel_spacing = PhysicalDeltaX of the first ultrasound region (0018,6011); frame_rate =
    CineRate, else round(1000 / FrameTime), else round(1000 / FrameTimeVector[1]); R_times = RWaveTimeVector."""
    md: Dict[str, Any] = {'pixel_spacing': None, 'frame_rate': None, 'R_times': None, 'R_wave_data_present': False}
    try:
        md['pixel_spacing'] = ds[0x0018, 0x6011][0]['PhysicalDeltaX'].value
    except (KeyError, AttributeError, IndexError, TypeError) as e:
        if verbose:
            logger.warning(f'No pixel spacing metadata: {e}. Flagging as no conversion factor.')
    try:
        if type(ds.RWaveTimeVector) != float and ds.RWaveTimeVector is not None:
            md['R_times'] = np.asarray(ds.RWaveTimeVector)
            md['R_wave_data_present'] = True
    except (AttributeError, KeyError, TypeError):
        pass
    try:
        md['frame_rate'] = ds.CineRate
    except (AttributeError, KeyError):
        try:
            md['frame_rate'] = np.round(1000 / float(ds.FrameTime))
        except (AttributeError, KeyError, ValueError, ZeroDivisionError):
            try:
                md['frame_rate'] = np.round(1000 / float(ds.FrameTimeVector[1]))
            except (AttributeError, KeyError, IndexError, ValueError, ZeroDivisionError, TypeError) as e:
                if verbose:
                    logger.warning(f'No frame rate information: {e}. Flagging as no conversion factor.')
    return md


def process_video(dcm_path: str, save_path: str, segmentor_model: Any, verbose: bool = True, mode: str = 'A4C',
                  bkgd_comp: str = 'none', flipLR: bool = False, no_saliency: bool = False, OF_algo: str = 'TVL1',
                  save_mask_subset: Optional[List[str]] = None, include_waveforms: bool = False,
                  waveform_folder: Optional[str] = None,
                  config: Optional[OpticalFlowCalculationConfig] = None, mask_fn=None, dicom_reader=None) -> None:
    """Signature of the reference (:478-483), same order of operations: read -> photometric conversion to RGB ->
    metadata -> gray2rgb -> flipLR -> masks -> flow -> container.  DICOM reading needs pydicom (or `dicom_reader`, a
    callable path -> dataset with `.pixel_array`, for tests); the masks come from `mask_fn(nparr, segmentor_model,
    mode, config)` (the reference's predict_movie / predict_movie_thres: SAM / Otsu, out of scope).  The container
    is written by the built-in HDF5 writer (hdf5.py), no h5py needed.  Physiologic waveform files
    (`include_waveforms`, waveform_loader.py) are outside this path and are refused rather than silently dropped."""
    if config is None:
        config = default_optical_flow_config()
    if mode == 'otsu':
        if bkgd_comp != 'none':
            raise ConfigurationError(f'bkgd_comp {bkgd_comp} is not supported in mode=otsu, can only support bkgd_comp=none')
        if save_mask_subset is not None:
            raise ConfigurationError('In mode=otsu, save_mask_subset must be None')
    if include_waveforms:
        raise ConfigurationError('include_waveforms=True is not supported: loading ECG/ART/CVP/PAP files '
                                 '(waveform_loader.py) is outside the TV-L1 path; R-wave times from the DICOM are kept')
    dcm = None
    if dicom_reader is None:
        try:
            import pydicom as dcm
        except ImportError as e:
            raise DICOMReadError(f'Failed to read DICOM file: {dcm_path} (pydicom is not installed)') from e
        dicom_reader = dcm.dcmread
    try:
        ds = dicom_reader(dcm_path)
        nparr = ds.pixel_array
    except Exception as e:
        raise DICOMReadError(f'Failed to read DICOM file: {dcm_path}') from e
    if dcm is not None:     # YBR -> RGB like the reference (:525-526)
        h = dcm.pixel_data_handlers
        if h.numpy_handler.should_change_PhotometricInterpretation_to_RGB(ds):
            nparr = h.convert_color_space(nparr, ds.PhotometricInterpretation, 'RGB')
    md = extract_dicom_metadata(ds, verbose)
    if nparr.ndim == 3 and nparr.shape[0] > 1:
        nparr = gray2rgb(nparr)                       # before the masks are predicted (:532-536)
    if flipLR:
        nparr = np.flip(nparr, axis=2)
    if mode not in ('A4C', 'RVIO_2class', 'otsu'):
        raise ConfigurationError(f'Input for mode must be [A4C, otsu, RVIO_2class], not {mode}.')
    if mask_fn is None:
        raise ConfigurationError('mask_fn is required: mask prediction (SAM / Otsu) is an input of this path')
    mask_dict = mask_fn(nparr, segmentor_model, mode, config)
    try:
        heart_rate = ds.HeartRate
    except (AttributeError, KeyError):
        heart_rate = 0
    result = process_frames(np.ascontiguousarray(nparr), mask_dict, md['pixel_spacing'], md['frame_rate'], mode,
                            bkgd_comp, no_saliency, OF_algo, save_mask_subset, config,
                            patient_id=str(getattr(ds, 'PatientID', '')), heart_rate=heart_rate)
    if md['R_wave_data_present']:
        result['RWaveTime'] = md['R_times']           # dataset 'RWaveTime' (:452-456)
    save_hdf5(save_path, result)


def chunk_bounds(total: int, nchunks: int, chunk_index: int):
    """The reference's chunking (:266-269): split = total // nchunks; chunk c owns [c*split, (c+1)*split); the
    remainder is dropped.  `--nchunks` maps onto GPU ranks: rank r == chunk r."""
    split = total // nchunks
    return chunk_index * split, (chunk_index + 1) * split


def process_folder(dcm_folder: str, save_folder: str, segmentor_model: Any, nchunks: int = 10, chunk_index: int = 0,
                   mode: str = 'RVIO_2class', bkgd_comp: str = 'none', flipLR: bool = False, verbose: bool = True,
                   recalculate: bool = False, no_saliency: bool = True, OF_algo: str = 'TVL1',
                   save_mask_subset: Optional[List[str]] = None, include_waveforms: bool = False,
                   waveform_folder: Optional[str] = None, pixel_spacing: Optional[float] = None,
                   frame_rate: Optional[float] = None, process_subset: bool = False,
                   file_subset_list: List[str] = [], mask_fn=None) -> None:
    """Signature and semantics of the reference (:243-290): chunked, skip-if-exists, per-file errors are logged
    and swallowed so that one bad clip does not kill the batch."""
    os.makedirs(save_folder, exist_ok=True)
    file_list = os.listdir(dcm_folder)
    if process_subset:
        if len(file_subset_list) == 0:
            print('ERROR! File subset list is empty!')
            return
        file_list = [f for f in file_list if f in file_subset_list]
    if include_waveforms and waveform_folder is None:
        print('ERROR if include_waveform is selected, must define waveform_folder!')
        return
    lo, hi = chunk_bounds(len(file_list), nchunks, chunk_index)
    for i in range(lo, hi):
        filename = file_list[i]
        save_path = os.path.join(save_folder, filename[:-3] + 'hdf5')
        if os.path.exists(save_path) and not recalculate:
            continue
        if filename[-3:] != 'dcm':
            logger.warning(f'File extension must be dcm, found {filename[-3:]}, skipping')
            continue
        try:
            process_video(os.path.join(dcm_folder, filename), save_path, segmentor_model, verbose=verbose, mode=mode,
                          bkgd_comp=bkgd_comp, flipLR=flipLR, no_saliency=no_saliency, OF_algo=OF_algo,
                          save_mask_subset=save_mask_subset, include_waveforms=include_waveforms,
                          waveform_folder=waveform_folder, mask_fn=mask_fn)
        except Exception as e:
            logger.error(f'Error processing {filename}: {e}')
            if verbose:
                traceback.print_exc()
