"""Minimal HDF5 writer + reader for the reference's container, without h5py / libhdf5 (neither exists in this image).

What the reference writes (optical_flow/calculate_optical_flow.py:399-472) and reads back
(optical_flow/optical_flow_dataset.py:45-114):

    datasets  'echo' (N,H,W) f16, 'flow' (N,H,W,2) f16, one (N,H,W,2) bool per mask label, optional 'RWaveTime';
              every one chunked + gzip level 9 (h5py `compression='gzip', compression_opts=9`)
    attrs     on 'flow': frame_rate, nframes, pixel_spacing, ID, HR, no_saliency, mode, units_converted,
              waveforms_present, labels (a list of str -> 1-D variable-length UTF-8 strings)

File structure written here (HDF5 File Format Specification, the classic "version 0" structures that every
libhdf5 reads): superblock v0; root group = object header v1 with a Symbol Table message -> local heap (link
names) + one v1 B-tree node (type 0) -> one symbol table node; per dataset an object header v1 with Dataspace (v1),
Datatype (v1), Fill Value (v2), Filter Pipeline (v1: deflate) and Data Layout (v3, chunked) messages plus the
Attribute (v1) messages; chunk index = v1 B-tree (type 1), one level, or two when a dataset has more chunks than a
node holds; variable-length strings live in one global heap collection.  Types follow h5py's conventions: numpy bool
-> ENUM {FALSE=0, TRUE=1} over int8, str -> variable-length UTF-8 string.  Chunks are whole frames (1, H, W[, C]);
they are deflated by a thread pool (zlib releases the GIL): gzip-9 is the producer's dominant host cost.

The reader understands what the writer emits (and the neighbouring common cases: contiguous layout, object header
continuation blocks, fixed-length strings, the shuffle filter), which is what tests/test_hdf5.py round-trips; when
h5py is importable the same tests cross-check both directions against it.
"""
from __future__ import annotations

import struct
import zlib
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
_CHUNK_K = 32                    # v1 B-tree for chunks: 2K entries per node (the library default for superblock v0)
_GROUP_INTERNAL_K = 16


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


# ================================================================================================ datatypes
def _dt_fixed(size: int, signed: bool) -> bytes:
    return struct.pack("<BBBBI", 0x10, 0x08 if signed else 0x00, 0, 0, size) + struct.pack("<HH", 0, size * 8)


def _dt_float(size: int) -> bytes:
    exp_bits, mant_bits, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[size]
    # class bits: little endian, mantissa normalisation 2 (implied MSB), sign bit location = size*8 - 1
    head = struct.pack("<BBBBI", 0x11, 0x20, size * 8 - 1, 0, size)
    return head + struct.pack("<HHBBBBI", 0, size * 8, mant_bits, exp_bits, 0, mant_bits, bias)


def _dt_bool_enum() -> bytes:
    base = _dt_fixed(1, True)
    names = _pad8(b"FALSE\x00") + _pad8(b"TRUE\x00")
    return struct.pack("<BBBBI", 0x18, 2, 0, 0, 1) + base + names + bytes([0, 1])


def _dt_vlen_str() -> bytes:
    base = struct.pack("<BBBBI", 0x13, 0x10, 0, 0, 1)          # 1-byte string, null terminated, UTF-8
    return struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0, 16) + base   # variable length: type string, null term, UTF-8


def _datatype_of(arr: np.ndarray) -> bytes:
    dt = arr.dtype
    if dt == np.bool_:
        return _dt_bool_enum()
    if dt.kind == "f":
        return _dt_float(dt.itemsize)
    if dt.kind in "iu":
        return _dt_fixed(dt.itemsize, dt.kind == "i")
    raise TypeError(f"no HDF5 type for numpy dtype {dt}")


def _dataspace(shape: Tuple[int, ...]) -> bytes:
    body = struct.pack("<BBBBI", 1, len(shape), 0, 0, 0)
    return body + b"".join(struct.pack("<Q", int(d)) for d in shape)


# ================================================================================================ writer
class _Writer:
    def __init__(self):
        self.buf = bytearray()

    def tell(self) -> int:
        return len(self.buf)

    def alloc(self, data: bytes, align: int = 8) -> int:
        self.buf += b"\x00" * (-len(self.buf) % align)
        addr = len(self.buf)
        self.buf += data
        return addr

    def patch(self, addr: int, data: bytes) -> None:
        self.buf[addr:addr + len(data)] = data


def _message(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHBBBB", mtype, len(data), flags, 0, 0, 0) + data


def _object_header(messages: List[bytes]) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\x00" * 4 + body


class _GlobalHeap:
    """one GCOL collection for the variable-length strings of the attributes"""

    def __init__(self):
        self.objects: List[bytes] = []

    def add(self, data: bytes) -> int:
        self.objects.append(data)
        return len(self.objects)                       # heap object index (1-based)

    def serialise(self) -> bytes:
        body = b""
        for i, obj in enumerate(self.objects, 1):
            body += struct.pack("<HHIQ", i, 1, 0, len(obj)) + _pad8(obj)
        size = max(4096, 16 + len(body) + 16)
        size += -size % 8
        free = size - 16 - len(body)
        body += struct.pack("<HHIQ", 0, 0, 0, free) + b"\x00" * (free - 16)     # object 0 = the free space
        return b"GCOL" + struct.pack("<BBBBQ", 1, 0, 0, 0, size) + body


def _attr_value(value: Any, gheap: _GlobalHeap, gheap_addr_slot: List[int]):
    """-> (datatype, dataspace, data, patch offsets of global-heap addresses inside data)"""
    patches: List[int] = []
    if isinstance(value, str):
        idx = gheap.add(value.encode("utf-8"))
        data = struct.pack("<IQI", len(value.encode("utf-8")), 0, idx)
        return _dt_vlen_str(), _dataspace(()), data, [4]
    if isinstance(value, (list, tuple)) and all(isinstance(v, str) for v in value) and len(value) > 0:
        data = b""
        for i, v in enumerate(value):
            raw = v.encode("utf-8")
            patches.append(len(data) + 4)
            data += struct.pack("<IQI", len(raw), 0, gheap.add(raw))
        return _dt_vlen_str(), _dataspace((len(value),)), data, patches
    if isinstance(value, (bool, np.bool_)):
        return _dt_bool_enum(), _dataspace(()), bytes([1 if value else 0]), []
    arr = np.asarray(value)
    if arr.dtype == object or arr.dtype.kind in "US":
        raise TypeError(f"attribute value {value!r} has no HDF5 mapping here")
    if arr.dtype.kind == "i" and arr.dtype.itemsize < 8:
        arr = arr.astype(np.int64) if isinstance(value, int) else arr
    return _datatype_of(arr), _dataspace(arr.shape), np.ascontiguousarray(arr).tobytes(), []   # (0-d stays scalar)


def _attribute_message(name: str, value: Any, gheap: _GlobalHeap, pending: List[Tuple[int, List[int]]], base_off: int) -> bytes:
    dt, ds, data, patches = _attr_value(value, gheap, [])
    nm = name.encode("utf-8") + b"\x00"
    head = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds))
    body = head + _pad8(nm) + _pad8(dt) + _pad8(ds)
    data_off = len(body)
    body += data
    msg = _message(0x000C, body)
    pending.append((base_off + 8 + data_off, patches))     # 8 = message header
    return msg


def _chunk_btree(w: _Writer, entries: List[Tuple[Tuple[int, ...], int, int]], rank1: int) -> int:
    """entries: (chunk offsets incl. the trailing 0, byte size, address) in index order -> address of the root node"""
    key_size = 8 + 8 * rank1
    node_size = 24 + (2 * _CHUNK_K + 1) * key_size + 2 * _CHUNK_K * 8

    def key(offs, nbytes):
        return struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", o) for o in offs)

    def node(level, items, last_key, left, right):
        body = b"TREE" + struct.pack("<BBHQQ", 1, level, len(items), left, right)
        for offs, nbytes, addr in items:
            body += key(offs, nbytes) + struct.pack("<Q", addr)
        body += last_key
        return body + b"\x00" * (node_size - len(body))

    def build(level, items, end_offs):
        if len(items) <= 2 * _CHUNK_K:
            return w.alloc(node(level, items, key(end_offs, 0), UNDEF, UNDEF))
        groups = [items[i:i + 2 * _CHUNK_K] for i in range(0, len(items), 2 * _CHUNK_K)]
        addrs = [w.alloc(b"\x00" * node_size) for _ in groups]
        ups = []
        for gi, g in enumerate(groups):
            nxt = groups[gi + 1][0] if gi + 1 < len(groups) else None
            last = key(nxt[0], 0) if nxt else key(end_offs, 0)
            w.patch(addrs[gi], node(level, g, last, addrs[gi - 1] if gi else UNDEF, addrs[gi + 1] if gi + 1 < len(groups) else UNDEF))
            ups.append((g[0][0], g[0][1], addrs[gi]))
        return build(level + 1, ups, end_offs)

    return build(0, entries, _end_key_offsets(entries, rank1))


def _end_key_offsets(entries, rank1):
    # the key after the last child: one chunk step past the last chunk along the slowest dimension
    last = list(entries[-1][0])
    last[0] += 1
    return tuple(last)


def _write_dataset(w: _Writer, arr: np.ndarray, attrs: Optional[Dict[str, Any]], gheap: _GlobalHeap,
                   pending: List[Tuple[int, List[int]]], level: int, pool: ThreadPoolExecutor) -> int:
    arr = np.ascontiguousarray(arr)
    if arr.ndim == 0:
        raise ValueError("scalar datasets are not part of the reference's container")
    # one chunk = one slice along the first axis (a frame); h5py would pick its own chunk shape, readers do not care
    chunk_shape = (1,) + arr.shape[1:]
    esz = 1 if arr.dtype == np.bool_ else arr.dtype.itemsize
    raw = arr.view(np.uint8) if arr.dtype == np.bool_ else arr
    comp = list(pool.map(lambda i: zlib.compress(raw[i].tobytes(), level), range(arr.shape[0])))
    rank1 = arr.ndim + 1
    entries = []
    for i, c in enumerate(comp):
        addr = w.alloc(c, align=8)
        entries.append(((i,) + (0,) * arr.ndim, len(c), addr))
    btree = _chunk_btree(w, entries, rank1) if entries else UNDEF
    layout = struct.pack("<BBB", 3, 2, rank1) + struct.pack("<Q", btree) + \
        b"".join(struct.pack("<I", d) for d in chunk_shape) + struct.pack("<I", esz)
    pipeline = struct.pack("<BBHI", 1, 1, 0, 0) + struct.pack("<HHHH", 1, 0, 1, 1) + struct.pack("<II", level, 0)
    fill = struct.pack("<BBBB", 2, 3, 2, 0)               # v2: allocate incrementally, write fill if set, undefined
    msgs = [_message(0x0001, _dataspace(arr.shape)), _message(0x0003, _datatype_of(arr), flags=1),
            _message(0x0005, fill), _message(0x000B, pipeline), _message(0x0008, layout)]
    hdr_addr_guess = w.tell() + (-w.tell() % 8)
    off = 16 + sum(len(m) for m in msgs)
    for k, v in (attrs or {}).items():
        m = _attribute_message(k, v, gheap, pending, hdr_addr_guess + off)
        msgs.append(m)
        off += len(m)
    addr = w.alloc(_object_header(msgs))
    assert addr == hdr_addr_guess
    return addr


def write_hdf5(path: str, datasets: Dict[str, np.ndarray], attrs: Optional[Dict[str, Dict[str, Any]]] = None,
               compression_level: int = 9, threads: int = 8) -> None:
    """datasets: name -> array (root-level datasets, like the reference's file); attrs: dataset name -> {attr: value}.
    None attribute values become NaN (h5py cannot store None; the reference only hits that when DICOM tags are missing)."""
    attrs = attrs or {}
    names = sorted(datasets)                               # symbol table nodes are ordered by name
    if not names:
        raise ValueError("no datasets")
    leaf_k = max(4, (len(names) + 1) // 2)
    w = _Writer()
    w.alloc(b"\x00" * 96)                                  # superblock, patched at the end
    root_hdr = w.alloc(b"\x00" * 40)                       # root object header: one Symbol Table message
    # local heap: "" at offset 0, then the link names
    heap_data = bytearray(b"\x00" * 8)
    name_off = {}
    for n in names:
        name_off[n] = len(heap_data)
        heap_data += _pad8(n.encode("utf-8") + b"\x00")
    heap_addr = w.alloc(b"\x00" * 32)
    heap_data_addr = w.alloc(bytes(heap_data))
    w.patch(heap_addr, b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap_data), 1, heap_data_addr))
    gheap = _GlobalHeap()
    pending: List[Tuple[int, List[int]]] = []
    obj_addr = {}
    with ThreadPoolExecutor(max_workers=max(1, threads)) as pool:
        for n in names:
            a = {k: (float("nan") if v is None else v) for k, v in attrs.get(n, {}).items()}
            obj_addr[n] = _write_dataset(w, np.asarray(datasets[n]), a, gheap, pending, compression_level, pool)
    if gheap.objects:
        gaddr = w.alloc(gheap.serialise())
        for data_addr, offs in pending:
            for o in offs:
                w.patch(data_addr + o, struct.pack("<Q", gaddr))
    # symbol table node + group B-tree
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for n in names:
        snod += struct.pack("<QQII", name_off[n], obj_addr[n], 0, 0) + b"\x00" * 16
    snod += b"\x00" * (8 + 2 * leaf_k * 40 - len(snod))
    snod_addr = w.alloc(snod)
    node_size = 24 + (2 * _GROUP_INTERNAL_K + 1) * 8 + 2 * _GROUP_INTERNAL_K * 8
    tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_addr, name_off[names[-1]])
    btree_addr = w.alloc(tree + b"\x00" * (node_size - len(tree)))
    w.patch(root_hdr, _object_header([_message(0x0011, struct.pack("<QQ", btree_addr, heap_addr))]))
    eof = w.tell() + (-w.tell() % 8)
    w.buf += b"\x00" * (eof - w.tell())
    sb = SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", leaf_k, _GROUP_INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", btree_addr, heap_addr)
    assert len(sb) == 96
    w.patch(0, sb)
    with open(path, "wb") as f:
        f.write(w.buf)


# ================================================================================================ reader
class _Reader:
    def __init__(self, data: bytes):
        self.d = data
        if data[:8] != SIGNATURE:
            raise ValueError("not an HDF5 file (signature)")
        if data[8] != 0 or data[13] != 8 or data[14] != 8:
            raise ValueError("only superblock version 0 with 8-byte offsets / lengths is supported")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", data, 16)
        self.root_btree, self.root_heap = struct.unpack_from("<QQ", data, 80)
        self.gheaps: Dict[int, Dict[int, bytes]] = {}

    # ---- groups
    def _heap_name(self, heap_addr: int, off: int) -> str:
        if self.d[heap_addr:heap_addr + 4] != b"HEAP":
            raise ValueError("bad local heap")
        data_addr = struct.unpack_from("<Q", self.d, heap_addr + 24)[0]
        end = self.d.index(b"\x00", data_addr + off)
        return self.d[data_addr + off:end].decode("utf-8")

    def _group_entries(self, btree: int, heap: int, out: Dict[str, int]) -> None:
        sig = self.d[btree:btree + 4]
        if sig == b"SNOD":
            n = struct.unpack_from("<H", self.d, btree + 6)[0]
            for i in range(n):
                name_off, obj = struct.unpack_from("<QQ", self.d, btree + 8 + 40 * i)
                out[self._heap_name(heap, name_off)] = obj
            return
        if sig != b"TREE":
            raise ValueError("bad group B-tree node")
        n = struct.unpack_from("<H", self.d, btree + 6)[0]
        pos = btree + 24 + 8
        for _ in range(n):
            child = struct.unpack_from("<Q", self.d, pos)[0]
            self._group_entries(child, heap, out)
            pos += 16

    def links(self) -> Dict[str, int]:
        out: Dict[str, int] = {}
        self._group_entries(self.root_btree, self.root_heap, out)
        return out

    # ---- object headers
    def messages(self, addr: int) -> List[Tuple[int, bytes]]:
        ver, _, nmsg, _, size = struct.unpack_from("<BBHII", self.d, addr)
        if ver != 1:
            raise ValueError("only version 1 object headers are supported")
        blocks = [(addr + 16, size)]
        out: List[Tuple[int, bytes]] = []
        while blocks and len(out) < nmsg:
            pos, left = blocks.pop(0)
            while left >= 8 and len(out) < nmsg:
                mtype, msize = struct.unpack_from("<HH", self.d, pos)
                body = self.d[pos + 8:pos + 8 + msize]
                if mtype == 0x0010:                       # continuation
                    blocks.append(struct.unpack_from("<QQ", body))
                out.append((mtype, body))
                pos += 8 + msize
                left -= 8 + msize
        return out

    # ---- datatypes
    def _dtype(self, b: bytes):
        """-> (kind, numpy dtype or None, size, consumed bytes); kind in {'num', 'bool', 'vlen_str', 'str'}"""
        cls, ver = b[0] & 0x0F, b[0] >> 4
        size = struct.unpack_from("<I", b, 4)[0]
        if cls == 0:
            signed = bool(b[1] & 0x08)
            return "num", np.dtype(("<i" if signed else "<u") + str(size)), size, 12
        if cls == 1:
            return "num", np.dtype("<f" + str(size)), size, 20
        if cls == 3:
            return "str", None, size, 8
        if cls == 8:
            nmemb = b[1] | (b[2] << 8)
            _, base_dt, bsize, used = self._dtype(b[8:])
            pos = 8 + used
            names = []
            for _ in range(nmemb):
                end = b.index(b"\x00", pos)
                names.append(b[pos:end].decode())
                pos += (end - pos + 1 + 7) // 8 * 8 if ver < 3 else end - pos + 1
            vals = list(np.frombuffer(b[pos:pos + nmemb * bsize], base_dt))
            if sorted(zip(names, vals)) == [("FALSE", 0), ("TRUE", 1)]:
                return "bool", np.dtype(np.bool_), size, pos + nmemb * bsize
            return "num", base_dt, size, pos + nmemb * bsize
        if cls == 9 and (b[1] & 0x0F) == 1:
            return "vlen_str", None, size, 16
        raise ValueError(f"unsupported datatype class {cls}")

    def _gheap_object(self, addr: int, index: int) -> bytes:
        if addr not in self.gheaps:
            if self.d[addr:addr + 4] != b"GCOL":
                raise ValueError("bad global heap collection")
            size = struct.unpack_from("<Q", self.d, addr + 8)[0]
            pos, end, objs = addr + 16, addr + size, {}
            while pos + 16 <= end:
                idx, _, _, osz = struct.unpack_from("<HHIQ", self.d, pos)
                if idx == 0:
                    break
                objs[idx] = self.d[pos + 16:pos + 16 + osz]
                pos += 16 + (osz + 7) // 8 * 8
            self.gheaps[addr] = objs
        return self.gheaps[addr][index]

    def _decode(self, kind, dt, size, shape, raw: bytes):
        n = int(np.prod(shape)) if shape else 1
        if kind == "vlen_str":
            vals = []
            for i in range(n):
                length, gaddr, idx = struct.unpack_from("<IQI", raw, 16 * i)
                vals.append(self._gheap_object(gaddr, idx)[:length].decode("utf-8"))
            return vals[0] if not shape else np.array(vals, dtype=object).reshape(shape)
        if kind == "str":
            vals = [raw[i * size:(i + 1) * size].split(b"\x00")[0] for i in range(n)]
            return vals[0] if not shape else np.array(vals).reshape(shape)
        arr = np.frombuffer(raw[:n * size], np.uint8 if kind == "bool" else dt)
        if kind == "bool":
            arr = arr.astype(np.bool_)
        return arr.reshape(shape)[()] if not shape else arr.reshape(shape).copy()

    @staticmethod
    def _shape(b: bytes) -> Tuple[int, ...]:
        ver, rank = b[0], b[1]
        off = 8 if ver == 1 else 4
        return tuple(struct.unpack_from("<Q", b, off + 8 * i)[0] for i in range(rank))

    # ---- chunk index
    def _chunks(self, node: int, rank1: int, out: List[Tuple[Tuple[int, ...], int, int, int]]) -> None:
        if self.d[node:node + 4] != b"TREE" or self.d[node + 4] != 1:
            raise ValueError("bad chunk B-tree node")
        level = self.d[node + 5]
        n = struct.unpack_from("<H", self.d, node + 6)[0]
        key = 8 + 8 * rank1
        pos = node + 24
        for _ in range(n):
            nbytes, mask = struct.unpack_from("<II", self.d, pos)
            offs = struct.unpack_from("<" + "Q" * rank1, self.d, pos + 8)
            child = struct.unpack_from("<Q", self.d, pos + key)[0]
            if level == 0:
                out.append((offs[:-1], nbytes, mask, child))
            else:
                self._chunks(child, rank1, out)
            pos += key + 8

    def dataset(self, addr: int):
        shape = kind = dt = size = None
        layout = None
        filters: List[Tuple[int, List[int]]] = []
        attrs: Dict[str, Any] = {}
        for mtype, body in self.messages(addr):
            if mtype == 0x0001:
                shape = self._shape(body)
            elif mtype == 0x0003:
                kind, dt, size, _ = self._dtype(body)
            elif mtype == 0x0008:
                layout = body
            elif mtype == 0x000B:
                nf = body[1]
                pos = 8
                for _ in range(nf):
                    fid, nlen, _, ncd = struct.unpack_from("<HHHH", body, pos)
                    pos += 8 + (nlen + 7) // 8 * 8
                    cd = list(struct.unpack_from("<" + "I" * ncd, body, pos))
                    pos += 4 * ncd + (4 if ncd % 2 else 0)
                    filters.append((fid, cd))
            elif mtype == 0x000C:
                ver, _, nlen, dlen, slen = struct.unpack_from("<BBHHH", body, 0)
                p8 = (lambda x: (x + 7) // 8 * 8) if ver == 1 else (lambda x: x)
                pos = 8
                name = body[pos:pos + nlen].split(b"\x00")[0].decode("utf-8")
                pos += p8(nlen)
                akind, adt, asize, _ = self._dtype(body[pos:pos + dlen])
                pos += p8(dlen)
                ashape = self._shape(body[pos:pos + slen])
                pos += p8(slen)
                attrs[name] = self._decode(akind, adt, asize, ashape, body[pos:])
        if shape is None or kind is None or layout is None:
            raise ValueError("object is not a dataset")
        lver, lclass = layout[0], layout[1]
        if lver != 3:
            raise ValueError("only version 3 data layout messages are supported")
        esz = size
        if lclass == 1:                                   # contiguous
            a, nbytes = struct.unpack_from("<QQ", layout, 2)
            return self._decode(kind, dt, size, shape, self.d[a:a + nbytes]), attrs
        if lclass != 2:
            raise ValueError("unsupported layout class")
        rank1 = layout[2]
        btree = struct.unpack_from("<Q", layout, 3)[0]
        cdims = struct.unpack_from("<" + "I" * rank1, layout, 11)[:-1]
        npdt = np.uint8 if kind == "bool" else dt
        out = np.zeros(shape, npdt)
        chunks: List[Tuple[Tuple[int, ...], int, int, int]] = []
        if btree != UNDEF:
            self._chunks(btree, rank1, chunks)
        for offs, nbytes, mask, caddr in chunks:
            raw = self.d[caddr:caddr + nbytes]
            for fi, (fid, cd) in reversed(list(enumerate(filters))):
                if mask & (1 << fi):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:                            # shuffle
                    n = len(raw) // esz
                    raw = np.frombuffer(raw, np.uint8).reshape(esz, n).T.tobytes()
                else:
                    raise ValueError(f"unsupported filter {fid}")
            block = np.frombuffer(raw, npdt).reshape(cdims)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
            out[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]
        return (out.astype(np.bool_) if kind == "bool" else out), attrs


def read_hdf5(path: str) -> Tuple[Dict[str, np.ndarray], Dict[str, Dict[str, Any]]]:
    """-> ({dataset name: array}, {dataset name: {attribute: value}}) of the root group."""
    with open(path, "rb") as f:
        r = _Reader(f.read())
    data, attrs = {}, {}
    for name, addr in r.links().items():
        data[name], attrs[name] = r.dataset(addr)
    return data, attrs
