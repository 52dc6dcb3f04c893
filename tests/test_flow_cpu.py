"""CPU tests of the host-side API mirror (flow.py) -- chunking, metadata, error behaviour that does not need a GPU --
and of the frame-prep restatement the GPU kernels are checked against (oracle/frame_prep_ref.py)."""
import numpy as np
import pytest


def test_img2uint8_follows_reference_formula():
    """optical_flow_utils.py:30-31: img_as_ubyte((img - min) / max) -- divides by max, not by the range"""
    from oracle.frame_prep_ref import img2uint8, rgb2gray
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, (12, 14, 3), dtype=np.uint8)
    g = rgb2gray(rgb)
    want_gray = (rgb.astype(np.float64) / 255.0) @ np.array([0.2125, 0.7154, 0.0721])
    assert np.allclose(g, want_gray, rtol=0, atol=1e-15)
    out = img2uint8(g)
    want = np.clip(np.rint(((g - g.min()) / g.max()) * 255.0), 0, 255).astype(np.uint8)
    assert out.dtype == np.uint8 and np.array_equal(out, want)
    assert out.max() < 255 or g.min() == 0        # the (sic) formula never reaches 255 unless min == 0


def test_prepare_frames_gray_and_rgb():
    from oracle.frame_prep_ref import prepare_frames
    rng = np.random.default_rng(1)
    gray = rng.integers(0, 256, (3, 8, 9), dtype=np.uint8)
    a = prepare_frames(gray)
    b = prepare_frames(np.stack([gray] * 3, axis=-1))
    assert a.shape == (3, 8, 9) and a.dtype == np.uint8 and np.array_equal(a, b)


def test_chunk_bounds_reexported():
    from tee_optical_flow_b200.flow import chunk_bounds
    assert chunk_bounds(23, 4, 3) == (15, 20)


def test_process_frames_validates_before_touching_the_gpu():
    from tee_optical_flow_b200.exceptions import ConfigurationError, OpticalFlowCalculationError
    from tee_optical_flow_b200.flow import process_frames
    fr = np.zeros((3, 16, 16), np.uint8)
    with pytest.raises(ConfigurationError):
        process_frames(fr, mode='otsu', bkgd_comp='WASE', frames_are_prepared=True)
    with pytest.raises(ConfigurationError):
        process_frames(fr, mode='otsu', save_mask_subset=['rv'], frames_are_prepared=True)
    with pytest.raises(ConfigurationError):
        process_frames(fr, mode='nope', frames_are_prepared=True)
    with pytest.raises(OpticalFlowCalculationError):
        process_frames(fr, OF_algo='farneback', frames_are_prepared=True)
    with pytest.raises(OpticalFlowCalculationError):
        process_frames(fr, bkgd_comp='median', frames_are_prepared=True)
    with pytest.raises(ConfigurationError):
        process_frames(fr, bkgd_comp='WASE', frames_are_prepared=True)      # no 'bkgd' mask


def test_process_frames_has_no_host_frame_prep_fallback():
    """unusual dtypes are refused instead of being prepared on the host (north_star: no CPU fallback)"""
    from tee_optical_flow_b200.exceptions import OpticalFlowCalculationError
    from tee_optical_flow_b200.flow import process_frames
    with pytest.raises(OpticalFlowCalculationError, match='uint8'):
        process_frames(np.zeros((3, 16, 16, 3), np.float32))
    with pytest.raises(OpticalFlowCalculationError, match='uint8'):
        process_frames(np.zeros((3, 16, 16), np.uint16))


class _Tag:
    def __init__(self, v): self.value = v


class _FakeDicom:
    """attribute / item access of a pydicom dataset, just enough for extract_dicom_metadata"""
    def __init__(self, region_dx=None, **attrs):
        self._region = region_dx
        for k, v in attrs.items():
            setattr(self, k, v)

    def __getitem__(self, key):
        if key == (0x0018, 0x6011) and self._region is not None:
            return [{'PhysicalDeltaX': _Tag(self._region)}]
        raise KeyError(key)


def test_dicom_metadata_fallback_chain_matches_reference():
    """calculate_optical_flow.py:315-367: CineRate -> round(1000/FrameTime) -> round(1000/FrameTimeVector[1]); every
    item independent of the others"""
    from tee_optical_flow_b200.flow import extract_dicom_metadata
    md = extract_dicom_metadata(_FakeDicom(region_dx=0.031, CineRate=47))
    assert md['pixel_spacing'] == 0.031 and md['frame_rate'] == 47 and md['R_wave_data_present'] is False
    md = extract_dicom_metadata(_FakeDicom(region_dx=0.031, FrameTime='21.5'))
    assert md['pixel_spacing'] == 0.031 and md['frame_rate'] == np.round(1000 / 21.5)
    md = extract_dicom_metadata(_FakeDicom(FrameTimeVector=[0, 33.3, 33.3], RWaveTimeVector=[10.0, 800.0]))
    assert md['pixel_spacing'] is None and md['frame_rate'] == np.round(1000 / 33.3)
    assert md['R_wave_data_present'] and md['R_times'].tolist() == [10.0, 800.0]
    md = extract_dicom_metadata(_FakeDicom(region_dx=0.05, FrameTime='0'))      # ZeroDivisionError -> next fallback
    assert md['pixel_spacing'] == 0.05 and md['frame_rate'] is None
    md = extract_dicom_metadata(_FakeDicom(RWaveTimeVector=3.5))                # a bare float is not a vector (:344)
    assert md['R_wave_data_present'] is False


def test_process_video_refuses_waveform_files_and_needs_masks(tmp_path):
    from tee_optical_flow_b200.exceptions import ConfigurationError, DICOMReadError
    from tee_optical_flow_b200.flow import process_video
    with pytest.raises(ConfigurationError, match='include_waveforms'):
        process_video('x.dcm', str(tmp_path / 'o.hdf5'), None, include_waveforms=True, waveform_folder='w')
    def bad_reader(path):
        raise IOError('nope')
    with pytest.raises(DICOMReadError):
        process_video('x.dcm', str(tmp_path / 'o.hdf5'), None, dicom_reader=bad_reader, mask_fn=lambda *a: {})
    class DS(_FakeDicom):
        pixel_array = np.zeros((4, 8, 8), np.uint8)
    with pytest.raises(ConfigurationError, match='mask_fn'):
        process_video('x.dcm', str(tmp_path / 'o.hdf5'), None, dicom_reader=lambda p: DS(), mode='RVIO_2class')


def test_process_folder_swallows_per_file_errors(tmp_path, caplog):
    """a failed clip must not kill the batch (calculate_optical_flow.py:281-284)"""
    from tee_optical_flow_b200.flow import process_folder
    d = tmp_path / "dcm"; d.mkdir()
    (d / "a.dcm").write_bytes(b"not a dicom")
    (d / "b.txt").write_text("x")
    process_folder(str(d), str(tmp_path / "out"), None, nchunks=1, chunk_index=0, verbose=False, mask_fn=lambda *a: {})
    assert (tmp_path / "out").is_dir() and not list((tmp_path / "out").iterdir())


def test_downstream_restatements_selfcheck():
    """the restated third-party helpers of the parity harness behave like their originals on known cases"""
    from oracle import downstream_ref as R
    x = np.sin(np.linspace(0, 6 * np.pi, 60)) + 0.01 * np.cos(np.linspace(0, 90, 60))
    pk = R.peak_indexes(x, thres=0.5, min_dist=5)
    assert pk.tolist() == [5, 25, 44] or all(abs(a - b) <= 1 for a, b in zip(pk.tolist(), [5, 25, 44]))
    assert R.peak_indexes(np.ones(10)).size == 0
    assert R.find_start_stop(np.array([0, 1, 2, 5, 6, 9])) == [[0, 2], [5, 6], [9, 9]]
    y = R.spectral_smooth(x, 0.3, 20)
    assert y.shape == x.shape and np.abs(y - np.sin(np.linspace(0, 6 * np.pi, 60))).max() < 0.1
    with pytest.raises(ValueError):
        R.spectral_smooth(np.ones(10), 0.3, 20)


def test_dataset_mirror_follows_reference_contract():
    """optical_flow_dataset.py:45-111,172-229 on the in-memory HDF5 layout"""
    from tee_optical_flow_b200.dataset import OpticalFlowDataset
    rng = np.random.default_rng(2)
    N, H, W = 7, 6, 8
    flow = rng.standard_normal((N, H, W, 2)).astype(np.float16)
    rv = rng.random((N, H, W, 1)) > 0.5
    res = {'flow': flow, 'echo': np.zeros((N, H, W), np.float16), 'rv': np.repeat(rv, 2, axis=-1),
           'attrs': {'nframes': N, 'mode': 'RVIO_2class', 'waveforms_present': False, 'units_converted': True,
                     'frame_rate': 40.0, 'pixel_spacing': 0.05, 'ID': 'x', 'labels': ['rv']}}
    ds = OpticalFlowDataset(res)
    assert ds.nframes == N - 2 and ds.vel_array.dtype == np.float32
    assert np.array_equal(ds.vel_array, flow.astype(np.float32))
    assert np.array_equal(ds.accel_array, np.gradient(flow.astype(np.float32), 1 / 40.0, axis=0))
    assert np.array_equal(ds.get_masked_arr('velocity', 'rv'), flow.astype(np.float32) * res['rv'])
    assert np.array_equal(ds.get_masked_arr('PWR', 'rv'), ds.vel_array * ds.accel_array * res['rv'])
    assert ds.get_masked_arr('velocity', 'nope') is None and ds.get_masked_arr('speed', 'rv') is None
    res['attrs']['units_converted'] = False
    assert OpticalFlowDataset(res).frame_rate == 1
