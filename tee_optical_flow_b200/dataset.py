"""OpticalFlowDataset mirror: the consumer-side contract of the producer's output
(optical_flow/optical_flow_dataset.py:29-229), built from the in-memory HDF5 layout that
``flow.process_frames`` returns, or from an HDF5 file through the package's own reader (hdf5.py).

Same attribute names and semantics as the reference: ``vel_array`` = flow.astype(float32) (:57), ``nframes`` =
attrs['nframes'] - 2 (:58), ``accel_array`` = np.gradient(vel, 1/frame_rate, axis=0) (:100), ``pwr_array`` = vel *
accel (:101), ``get_masked_arr(param, label)`` = param array * mask (:189-197).  Host-side numpy, like the
reference; the heavy per-frame reductions on these arrays live in ``analysis.py`` (GPU).
"""
from __future__ import annotations

from typing import Any, Dict

import numpy as np


class OpticalFlowDataset:
    accepted_params = ['velocity', 'acceleration', 'PWR']

    def __init__(self, source, keep_file_open: bool = False):
        self.GRAPH_CALCULATED = False
        self.CARDIACCYCLE_CALCULATED = False
        if isinstance(source, dict):
            self._from_result(source)
        else:
            self._from_hdf5(str(source))

    # ------------------------------------------------------------------ constructors
    def _from_result(self, res: Dict[str, Any]):
        attrs = res['attrs']
        self.filename = str(attrs.get('ID', ''))
        self._init_common(np.asarray(res['flow']), np.asarray(res['echo']), attrs,
                          {k: np.asarray(res[k]) for k in attrs['labels']})

    def _from_hdf5(self, path: str):
        """the reference's constructor (optical_flow_dataset.py:45-111) on an HDF5 file, through the package's own
        reader (hdf5.py) -- h5py is not needed"""
        import os
        from .hdf5 import read_hdf5
        data, all_attrs = read_hdf5(path)
        attrs = all_attrs['flow']
        self.filename = os.path.basename(path)[:-4]
        masks = {str(k): data[str(k)] for k in attrs['labels']}
        self._init_common(data['flow'], data['echo'], attrs, masks)
        if 'RWaveTime' in data:
            self.RTimePresent = True
            self.RWaveTimes = data['RWaveTime']

    def _init_common(self, flow, echo, attrs, masks):
        self.echo_array = echo
        self.vel_array = flow.astype(np.float32)                      # (N, H, W, 2)
        self.nframes = int(attrs['nframes']) - 2
        self.mode = attrs['mode']
        self.RTimePresent = False
        self.waveforms_present = bool(attrs['waveforms_present'])
        self.units_converted_flag = bool(attrs['units_converted'])
        if self.units_converted_flag:
            # numpy scalars, as h5py hands attributes back: np.gradient's spacing type decides its working precision
            self.frame_rate = np.asarray(attrs['frame_rate'])[()]
            self.pixel_spacing = np.asarray(attrs['pixel_spacing'])[()]
            self.ID = attrs['ID']
        else:
            self.frame_rate = 1
            self.pixel_spacing = 1
        self.accel_array = np.gradient(self.vel_array, 1 / self.frame_rate, axis=0)
        self.pwr_array = self.vel_array * self.accel_array
        self.accepted_labels = list(attrs['labels'])
        self.mask_ds_dict = dict(masks)

    # ------------------------------------------------------------------ getters (reference names and error
    # behaviour: an invalid key prints an error and returns None, optical_flow_dataset.py:172-229)
    def _validate_label(self, label):
        return label in self.accepted_labels

    def _validate_param(self, param):
        return param in self.accepted_params

    def get_echo(self):
        return self.echo_array

    def get_mask(self, label):
        if self._validate_label(label):
            return self.mask_ds_dict[label]
        print(f'ERROR {label} not a valid key. Choose from {self.accepted_labels}')
        return None

    def _masked(self, arr, label):
        mask = self.get_mask(label)
        return None if mask is None else arr * mask

    def get_velocity(self, label):
        return self._masked(self.vel_array, label)

    def get_accel(self, label):
        return self._masked(self.accel_array, label)

    def get_pwr(self, label):
        return self._masked(self.pwr_array, label)

    def get_masked_arr(self, param, label):
        if param == 'velocity':
            return self.get_velocity(label)
        elif param == 'acceleration':
            return self.get_accel(label)
        elif param == 'PWR':
            return self.get_pwr(label)
        print(f'ERROR! {param} is not a valid optical flow parameter, choose from {self.accepted_params}')
        return None

    def stored_flow_f16(self) -> np.ndarray:
        """the fp16 flow as stored (input of the GPU analysis entry points)"""
        return self.vel_array.astype(np.float16)

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False
