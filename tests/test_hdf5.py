"""The package's HDF5 writer / reader (tee_optical_flow_b200/hdf5.py) on the reference's container layout
(calculate_optical_flow.py:399-472 written, optical_flow_dataset.py:45-114 read): byte-level structure checks against
the HDF5 file format specification, round trips, and -- when h5py is importable -- both directions against libhdf5."""
import struct
import zlib

import numpy as np
import pytest

from tee_optical_flow_b200 import hdf5 as H


def _container(n=5, h=12, w=16, seed=0, labels=('rv', 'av', 'bkgd')):
    rng = np.random.default_rng(seed)
    flow = rng.standard_normal((n, h, w, 2)).astype(np.float16)
    echo = rng.random((n, h, w)).astype(np.float16)
    data = {'echo': echo, 'flow': flow}
    for k in labels:
        data[k] = np.repeat(rng.random((n, h, w, 1)) > 0.5, 2, axis=-1)
    attrs = {'frame_rate': 47.0, 'nframes': n, 'pixel_spacing': 0.0312, 'ID': 'patient-äö-7', 'HR': 72,
             'no_saliency': True, 'mode': 'RVIO_2class', 'units_converted': True, 'waveforms_present': False,
             'labels': list(labels)}
    return data, attrs


def test_round_trip_reference_container(tmp_path):
    data, attrs = _container()
    data['RWaveTime'] = np.array([12.5, 840.0, 1660.25])
    p = tmp_path / "clip.hdf5"
    H.write_hdf5(str(p), data, {'flow': attrs})
    got, gattrs = H.read_hdf5(str(p))
    assert set(got) == set(data)
    for k, v in data.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k
    a = gattrs['flow']
    assert a['nframes'] == 5 and a['nframes'].dtype == np.int64 and a['HR'] == 72
    assert a['frame_rate'] == 47.0 and a['pixel_spacing'] == 0.0312
    assert a['ID'] == 'patient-äö-7' and a['mode'] == 'RVIO_2class'
    assert a['no_saliency'] is np.True_ or a['no_saliency'] == True
    assert bool(a['waveforms_present']) is False and bool(a['units_converted']) is True
    assert list(a['labels']) == ['rv', 'av', 'bkgd'] and a['labels'].dtype == object
    assert gattrs['echo'] == {}


def test_file_structure_follows_the_format_specification(tmp_path):
    data, attrs = _container(n=3, labels=('rv',))
    p = tmp_path / "s.hdf5"
    H.write_hdf5(str(p), data, {'flow': attrs})
    b = p.read_bytes()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0                 # signature, superblock version 0
    assert b[13] == 8 and b[14] == 8                                     # 8-byte offsets and lengths
    base, free, eof, drv = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and free == H.UNDEF and drv == H.UNDEF and eof == len(b) and eof % 8 == 0
    name_off, root_hdr, cache, _ = struct.unpack_from("<QQII", b, 56)
    btree, heap = struct.unpack_from("<QQ", b, 80)
    assert cache == 1 and b[root_hdr] == 1                               # cached symbol-table info, v1 object header
    mtype, msize = struct.unpack_from("<HH", b, root_hdr + 16)
    assert mtype == 0x0011 and struct.unpack_from("<QQ", b, root_hdr + 24) == (btree, heap)
    assert b[btree:btree + 4] == b"TREE" and b[btree + 4] == 0 and b[heap:heap + 4] == b"HEAP"
    snod = struct.unpack_from("<Q", b, btree + 24 + 8)[0]
    assert b[snod:snod + 4] == b"SNOD" and struct.unpack_from("<H", b, snod + 6)[0] == 3
    # link names are sorted and 8-byte aligned in the local heap
    r = H._Reader(b)
    links = r.links()
    assert list(links) == sorted(links) == ['echo', 'flow', 'rv'] and all(a % 8 == 0 for a in links.values())
    msgs = dict((t, body) for t, body in r.messages(links['flow']) if t != 0x000C)
    assert set(msgs) == {0x0001, 0x0003, 0x0005, 0x0008, 0x000B}
    assert msgs[0x0003][0] == 0x11 and struct.unpack_from("<I", msgs[0x0003], 4)[0] == 2      # IEEE float, 2 bytes
    assert struct.unpack_from("<HHBBBBI", msgs[0x0003], 8) == (0, 16, 10, 5, 0, 10, 15)       # binary16 field layout
    lay = msgs[0x0008]
    assert lay[0] == 3 and lay[1] == 2 and lay[2] == 5                   # v3, chunked, rank + 1
    assert struct.unpack_from("<5I", lay, 11) == (1, 12, 16, 2, 2)       # one frame per chunk, element size last
    pl = msgs[0x000B]
    assert pl[0] == 1 and pl[1] == 1 and struct.unpack_from("<H", pl, 8)[0] == 1 and struct.unpack_from("<I", pl, 16)[0] == 9
    # the chunk B-tree points at plain zlib streams of whole frames
    node = struct.unpack_from("<Q", lay, 3)[0]
    assert b[node:node + 4] == b"TREE" and b[node + 4] == 1 and struct.unpack_from("<H", b, node + 6)[0] == 3
    nbytes, mask = struct.unpack_from("<II", b, node + 24)
    offs = struct.unpack_from("<5Q", b, node + 32)
    addr = struct.unpack_from("<Q", b, node + 24 + 48)[0]
    assert offs == (0, 0, 0, 0, 0) and mask == 0
    assert zlib.decompress(b[addr:addr + nbytes]) == data['flow'][0].tobytes()
    # the mask is h5py's bool: ENUM {FALSE = 0, TRUE = 1} over a signed byte
    mm = dict(r.messages(links['rv']))
    dt = mm[0x0003]
    assert dt[0] == 0x18 and dt[1] == 2 and dt[8] == 0x10 and b"FALSE\x00" in dt and b"TRUE\x00" in dt and dt[36:38] == b"\x00\x01"


def test_many_chunks_use_a_two_level_index(tmp_path):
    rng = np.random.default_rng(1)
    arr = rng.integers(0, 1000, (150, 4, 3)).astype(np.int32)          # 150 chunks > 64 entries per node
    p = tmp_path / "m.hdf5"
    H.write_hdf5(str(p), {'x': arr, 'y': arr[:64].astype(np.float64)}, compression_level=1)
    b = p.read_bytes()
    r = H._Reader(b)
    lay = dict(r.messages(r.links()['x']))[0x0008]
    root = struct.unpack_from("<Q", lay, 3)[0]
    assert b[root + 5] == 1 and struct.unpack_from("<H", b, root + 6)[0] == 3      # level 1, three leaf nodes
    got, _ = H.read_hdf5(str(p))
    assert np.array_equal(got['x'], arr) and np.array_equal(got['y'], arr[:64].astype(np.float64))


def test_dataset_mirror_opens_the_written_file(tmp_path):
    """producer dict -> save_hdf5 -> OpticalFlowDataset(path): the consumer contract on a real file"""
    from tee_optical_flow_b200.dataset import OpticalFlowDataset
    from tee_optical_flow_b200.flow import save_hdf5
    data, attrs = _container(n=7)
    res = dict(data)
    res['attrs'] = attrs
    res['RWaveTime'] = np.array([5.0, 900.0])
    p = tmp_path / "clip.hdf5"
    save_hdf5(str(p), res)
    ds = OpticalFlowDataset(str(p))
    ref = OpticalFlowDataset(res)
    assert ds.nframes == ref.nframes == 5 and ds.mode == 'RVIO_2class' and ds.frame_rate == 47.0
    assert np.array_equal(ds.vel_array, ref.vel_array) and np.array_equal(ds.accel_array, ref.accel_array)
    assert np.array_equal(ds.get_masked_arr('velocity', 'rv'), ref.get_masked_arr('velocity', 'rv'))
    assert ds.RTimePresent and np.array_equal(ds.RWaveTimes, [5.0, 900.0])
    assert ds.filename == 'clip.'                                        # the reference's os.path.basename(path)[:-4]


def test_none_attributes_and_empty_label_list(tmp_path):
    data, attrs = _container(n=2, labels=())
    attrs.update(frame_rate=None, pixel_spacing=None, units_converted=False)
    p = tmp_path / "n.hdf5"
    H.write_hdf5(str(p), data, {'flow': attrs})
    _, a = H.read_hdf5(str(p))
    assert np.isnan(a['flow']['frame_rate']) and np.isnan(a['flow']['pixel_spacing']) and len(a['flow']['labels']) == 0


def test_reader_rejects_foreign_files(tmp_path):
    p = tmp_path / "x.hdf5"
    p.write_bytes(b"not hdf5 at all" * 10)
    with pytest.raises(ValueError):
        H.read_hdf5(str(p))


def test_cross_check_against_h5py_when_available(tmp_path):
    """both directions against libhdf5 (skipped in this image: h5py is not installed)"""
    h5py = pytest.importorskip("h5py")
    data, attrs = _container()
    p = tmp_path / "mine.hdf5"
    H.write_hdf5(str(p), data, {'flow': attrs})
    with h5py.File(str(p), 'r') as f:
        for k, v in data.items():
            assert f[k].dtype == v.dtype and np.array_equal(f[k][()], v)
            assert f[k].compression == 'gzip' and f[k].compression_opts == 9
        assert f['flow'].attrs['mode'] == 'RVIO_2class' and list(f['flow'].attrs['labels']) == attrs['labels']
        assert f['flow'].attrs['nframes'] == 5 and bool(f['flow'].attrs['no_saliency']) is True
    q = tmp_path / "theirs.hdf5"
    with h5py.File(str(q), 'w', libver='earliest') as f:
        for k, v in data.items():
            f.create_dataset(k, data=v, compression='gzip', compression_opts=9)
        for k, v in attrs.items():
            f['flow'].attrs[k] = v
    got, gattrs = H.read_hdf5(str(q))
    for k, v in data.items():
        assert np.array_equal(got[k], v)
    assert gattrs['flow']['mode'] == 'RVIO_2class' and list(gattrs['flow']['labels']) == attrs['labels']
