"""Minimal pure-Python HDF5 reader / writer for the MLGWSC-1 strain and trigger files.

h5py is not available in this image, and the reference does all its I/O through it
(MLGWSC-1/inference.py:197-210 reads `file[det][str(int(start))]` datasets with `start_time` / `delta_t`
attributes -- written by pycbc's `TimeSeries.save(path, group=f"{det}/{int(start)}")`, generate_data.py:197-216,
i.e. chunked + shuffle + gzip float64 -- and :667-672 writes `time`, `stat`, `var`, `all_vals`).  This module
implements the subset of the HDF5 file format (version-0/1 superblock, version-1 object headers, symbol-table
groups, version-1 B-trees) those files use, with an h5py-like surface:

    with File(path, "r") as f:  f["H1"]["1238166018"][()], .attrs["start_time"], len(ds), ds.dtype, f.keys()
    with File(path, "w") as f:  f.create_dataset("time", data=arr); f.create_group("H1"); ds.attrs["delta_t"] = ..
    File(path, "a"): read what is there, append, rewrite on close

Reader: contiguous, compact and chunked layouts; deflate, shuffle and fletcher32 filters; little/big-endian
integers and IEEE floats, fixed-length and variable-length strings; scalar / simple attributes.
Writer: contiguous datasets of numeric numpy arrays (and fixed-length byte strings), scalar or 1-D numeric /
string attributes, nested groups.  Files it writes follow the layout libhdf5 itself produces for
`libver="earliest"` (checked structurally against the reference's own h5py-written
Signal_vs_Noise/results/Real_events/results_2_detectors_real_events.hdf in tests/test_hdf5io.py).
"""
from __future__ import annotations

import os
import struct
import zlib
from typing import Dict, Iterator, List, Optional, Tuple, Union

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"


# =================================================================================================
# reader
# =================================================================================================
class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        off = 0
        while True:                                   # the superblock may sit at 0, 512, 1024, ...
            if buf[off:off + 8] == SIGNATURE:
                break
            off = 512 if off == 0 else off * 2
            if off + 8 > len(buf):
                raise OSError("not an HDF5 file (signature not found)")
        ver = buf[off + 8]
        if ver not in (0, 1):
            raise OSError(f"HDF5 superblock version {ver} is not supported (only the classic 0/1 layout that "
                          "h5py's default libver='earliest' writes)")
        so, sl = buf[off + 13], buf[off + 14]
        if so != 8 or sl != 8:
            raise OSError("only 8-byte offsets/lengths are supported")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", buf, off + 16)
        p = off + 24 + (4 if ver == 1 else 0)
        self.base, _free, self.eof, _drv = struct.unpack_from("<QQQQ", buf, p)
        p += 32
        # root group symbol table entry
        _name_off, self.root_header, cache_type = struct.unpack_from("<QQI", buf, p)

    # ---- object headers --------------------------------------------------------------------------
    def messages(self, addr: int) -> List[Tuple[int, int, bytes]]:
        """[(type, flags, data)] of the version-1 object header at `addr` (continuations followed)."""
        b = self.b
        addr += self.base
        if b[addr:addr + 4] == b"OHDR":
            raise OSError("version-2 object headers (libver='latest') are not supported")
        ver, _r, nmsg, _ref, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise OSError(f"object header version {ver} not supported")
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg + 64:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, p)
                data = b[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x0010:                   # continuation
                    caddr, clen = struct.unpack_from("<QQ", data, 0)
                    blocks.append((caddr + self.base, clen))
                out.append((mtype, flags, data))
        return out

    # ---- groups ----------------------------------------------------------------------------------
    def group_links(self, header_addr: int) -> Dict[str, int]:
        links: Dict[str, int] = {}
        for mtype, _f, data in self.messages(header_addr):
            if mtype == 0x0011:
                btree, heap = struct.unpack_from("<QQ", data, 0)
                heap_data = self._local_heap(heap)
                self._walk_group_btree(btree, heap_data, links)
            elif mtype == 0x0006:
                raise OSError("link messages (new-style groups) are not supported")
        return links

    def _local_heap(self, addr: int) -> bytes:
        b = self.b
        addr += self.base
        if b[addr:addr + 4] != b"HEAP":
            raise OSError("bad local heap signature")
        size, _free, daddr = struct.unpack_from("<QQQ", b, addr + 8)
        return b[daddr + self.base:daddr + self.base + size]

    def _walk_group_btree(self, addr: int, heap: bytes, links: Dict[str, int]) -> None:
        b = self.b
        a = addr + self.base
        if b[a:a + 4] == b"SNOD":
            _v, _r, nsym = struct.unpack_from("<BBH", b, a + 4)
            p = a + 8
            for _ in range(nsym):
                name_off, hdr = struct.unpack_from("<QQ", b, p)
                end = heap.index(b"\x00", name_off)
                links[heap[name_off:end].decode("utf-8")] = hdr
                p += 40
            return
        if b[a:a + 4] != b"TREE":
            raise OSError("bad B-tree signature")
        ntype, _level, used = struct.unpack_from("<BBH", b, a + 4)
        if ntype != 0:
            raise OSError("expected a group B-tree node")
        p = a + 24 + 8                                 # skip key 0
        for _ in range(used):
            (child,) = struct.unpack_from("<Q", b, p)
            self._walk_group_btree(child, heap, links)
            p += 16                                    # child + next key

    # ---- datatype / dataspace / attributes -----------------------------------------------------------
    def _dtype(self, data: bytes, off: int = 0):
        """-> (numpy dtype or ('vlen_str',) / ('vlen', base), size, bytes consumed)"""
        cv, b0, b1, _b2, size = struct.unpack_from("<BBBBI", data, off)
        cls = cv & 0x0F
        if cls == 0:                                   # fixed point
            order = ">" if (b0 & 1) else "<"
            signed = bool(b0 & 0x08)
            return np.dtype(f"{order}{'i' if signed else 'u'}{size}"), size, 8 + 4
        if cls == 1:                                   # float
            order = ">" if (b0 & 1) else "<"
            if size not in (2, 4, 8):
                raise OSError(f"float of {size} bytes not supported")
            return np.dtype(f"{order}f{size}"), size, 8 + 12
        if cls == 3:                                   # fixed-length string
            return np.dtype(f"S{size}"), size, 8
        if cls == 9:                                   # variable length
            is_str = (b0 & 0x0F) == 1
            base, _bs, used = self._dtype(data, off + 8)
            return ("vlen_str",) if is_str else ("vlen", base), size, 8 + used
        if cls == 6:                                   # compound: not needed for this path
            raise OSError("compound datatypes are not supported")
        if cls == 8:                                   # enum (h5py bool): read as its base integer
            base, _bs, used = self._dtype(data, off + 8)
            return base, size, 8 + used
        raise OSError(f"datatype class {cls} not supported")

    @staticmethod
    def _dataspace(data: bytes) -> Tuple[int, ...]:
        ver = data[0]
        rank = data[1]
        if ver == 1:
            p = 8
        elif ver == 2:
            if data[3] == 2:                           # null dataspace
                return (0,)
            p = 4
        else:
            raise OSError(f"dataspace version {ver} not supported")
        return tuple(struct.unpack_from("<" + "Q" * rank, data, p)) if rank else ()

    def _global_heap_object(self, addr: int, index: int) -> bytes:
        b = self.b
        a = addr + self.base
        if b[a:a + 4] != b"GCOL":
            raise OSError("bad global heap signature")
        (csize,) = struct.unpack_from("<Q", b, a + 8)
        p, end = a + 16, a + csize
        while p + 16 <= end:
            idx, _ref, _res, osize = struct.unpack_from("<HHIQ", b, p)
            if idx == 0:
                break
            if idx == index:
                return b[p + 16:p + 16 + osize]
            p += 16 + ((osize + 7) & ~7)
        raise OSError("global heap object not found")

    def _decode(self, raw: bytes, dt, shape: Tuple[int, ...]):
        n = int(np.prod(shape)) if shape else 1
        if isinstance(dt, tuple):
            if dt[0] != "vlen_str":
                raise OSError("variable-length sequences are not supported")
            vals = []
            for i in range(n):
                _ln, gaddr, gidx = struct.unpack_from("<IQI", raw, i * 16)
                vals.append(self._global_heap_object(gaddr, gidx).decode("utf-8") if gaddr not in (0, UNDEF) else "")
            arr = np.array(vals, dtype=object).reshape(shape)
            return arr if shape else arr[()]
        arr = np.frombuffer(raw, dtype=dt, count=n).reshape(shape)
        return arr.astype(dt.newbyteorder("=")) if shape else arr.astype(dt.newbyteorder("="))[()]

    def attributes(self, header_addr: int) -> Dict[str, object]:
        out: Dict[str, object] = {}
        for mtype, _f, data in self.messages(header_addr):
            if mtype != 0x000C:
                continue
            ver = data[0]
            if ver == 1:
                nsz, tsz, ssz = struct.unpack_from("<HHH", data, 2)
                p = 8
                pad = lambda v: (v + 7) & ~7          # noqa: E731
            elif ver in (2, 3):
                nsz, tsz, ssz = struct.unpack_from("<HHH", data, 2)
                p = 8 + (1 if ver == 3 else 0)
                pad = lambda v: v                     # noqa: E731
            else:
                raise OSError(f"attribute message version {ver} not supported")
            name = data[p:p + nsz].split(b"\x00")[0].decode("utf-8")
            p += pad(nsz)
            dt, _size, _used = self._dtype(data, p)
            p += pad(tsz)
            shape = self._dataspace(data[p:p + ssz])
            p += pad(ssz)
            out[name] = self._decode(data[p:], dt, shape)
        return out

    # ---- datasets --------------------------------------------------------------------------------
    def dataset_info(self, header_addr: int) -> dict:
        info = {"filters": [], "layout": None}
        for mtype, _f, data in self.messages(header_addr):
            if mtype == 0x0001:
                info["shape"] = self._dataspace(data)
            elif mtype == 0x0003:
                info["dtype"], info["itemsize"], _ = self._dtype(data, 0)
            elif mtype == 0x0008:
                ver = data[0]
                if ver != 3:
                    raise OSError(f"data layout version {ver} not supported")
                cls = data[1]
                if cls == 0:
                    (sz,) = struct.unpack_from("<H", data, 2)
                    info["layout"] = ("compact", data[4:4 + sz])
                elif cls == 1:
                    a, sz = struct.unpack_from("<QQ", data, 2)
                    info["layout"] = ("contiguous", a, sz)
                elif cls == 2:
                    rank = data[2]
                    (a,) = struct.unpack_from("<Q", data, 3)
                    dims = struct.unpack_from("<" + "I" * rank, data, 11)
                    info["layout"] = ("chunked", a, dims)     # dims include the element size as the last entry
                else:
                    raise OSError(f"layout class {cls} not supported")
            elif mtype == 0x000B:
                ver, nf = data[0], data[1]
                p = 8 if ver == 1 else 2
                for _ in range(nf):
                    if ver == 1 or struct.unpack_from("<H", data, p)[0] >= 256:
                        fid, nlen, _flags, ncd = struct.unpack_from("<HHHH", data, p)
                        p += 8 + (((nlen + 7) & ~7) if ver == 1 else nlen)
                    else:                              # version 2, library filter: no name length / name
                        fid, _flags, ncd = struct.unpack_from("<HHH", data, p)
                        p += 6
                    cd = struct.unpack_from("<" + "I" * ncd, data, p)
                    p += 4 * ncd
                    if ver == 1 and ncd % 2:
                        p += 4
                    info["filters"].append((fid, cd))
        if "shape" not in info or "dtype" not in info or info["layout"] is None:
            raise OSError("object is not a dataset")
        return info

    def _unfilter(self, raw: bytes, filters, mask: int, itemsize: int) -> bytes:
        for i in reversed(range(len(filters))):
            if mask & (1 << i):
                continue
            fid, _cd = filters[i]
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                n = len(raw) // itemsize
                arr = np.frombuffer(raw, dtype=np.uint8, count=n * itemsize).reshape(itemsize, n)
                raw = arr.T.tobytes() + raw[n * itemsize:]
            elif fid == 3:
                raw = raw[:-4]
            else:
                raise OSError(f"HDF5 filter {fid} not supported")
        return raw

    def read_dataset(self, info: dict) -> np.ndarray:
        shape, dt = info["shape"], info["dtype"]
        lay = info["layout"]
        n = int(np.prod(shape)) if shape else 1
        if lay[0] == "compact":
            return self._decode(lay[1], dt, shape)
        if lay[0] == "contiguous":
            _, a, sz = lay
            if a == UNDEF:                             # never written: fill value (zeros)
                return np.zeros(shape, dtype=dt if not isinstance(dt, tuple) else object)
            return self._decode(self.b[a + self.base:a + self.base + n * info["itemsize"]], dt, shape)
        _, a, cdims = lay
        if isinstance(dt, tuple):
            raise OSError("chunked variable-length datasets are not supported")
        chunk = tuple(cdims[:-1])
        out = np.zeros(shape, dtype=dt.newbyteorder("="))
        if a != UNDEF:
            self._walk_chunk_btree(a, len(shape), chunk, info, out)
        return out

    def _walk_chunk_btree(self, addr: int, rank: int, chunk, info, out: np.ndarray) -> None:
        b = self.b
        a = addr + self.base
        if b[a:a + 4] != b"TREE":
            raise OSError("bad chunk B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", b, a + 4)
        if ntype != 1:
            raise OSError("expected a raw-data chunk B-tree node")
        keysize = 8 + 8 * (rank + 1)
        p = a + 24
        for _ in range(used):
            csize, mask = struct.unpack_from("<II", b, p)
            offs = struct.unpack_from("<" + "Q" * rank, b, p + 8)
            (child,) = struct.unpack_from("<Q", b, p + keysize)
            p += keysize + 8
            if level > 0:
                self._walk_chunk_btree(child, rank, chunk, info, out)
                continue
            raw = self._unfilter(b[child + self.base:child + self.base + csize], info["filters"], mask, info["itemsize"])
            arr = np.frombuffer(raw, dtype=info["dtype"], count=int(np.prod(chunk))).reshape(chunk)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, out.shape))
            out[sl] = arr[tuple(slice(0, s.stop - s.start) for s in sl)]


class AttributeDict(dict):
    """`.attrs` of a group / dataset (a plain dict; the writer serialises it on close)."""


class Dataset:
    def __init__(self, name: str, data: Optional[np.ndarray] = None, reader: Optional[_Reader] = None,
                 header: Optional[int] = None):
        self.name = name
        self._data = data
        self._reader, self._header = reader, header
        self._info = reader.dataset_info(header) if reader is not None else None
        self.attrs = AttributeDict(reader.attributes(header) if reader is not None else {})

    def _load(self) -> np.ndarray:
        if self._data is None:
            self._data = self._reader.read_dataset(self._info)
        return self._data

    @property
    def shape(self) -> Tuple[int, ...]:
        return tuple(self._info["shape"]) if self._data is None else tuple(np.shape(self._data))

    @property
    def dtype(self):
        if self._data is None:
            dt = self._info["dtype"]
            return np.dtype(object) if isinstance(dt, tuple) else dt.newbyteorder("=")
        return np.asarray(self._data).dtype

    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def size(self) -> int:
        return int(np.prod(self.shape)) if self.shape else 1

    def __len__(self) -> int:
        if not self.shape:
            raise TypeError("len() of a scalar dataset")
        return self.shape[0]

    def __getitem__(self, idx):
        return self._load()[idx]

    def __array__(self, dtype=None, copy=None):
        a = np.asarray(self._load())
        return a.astype(dtype) if dtype is not None else a


class Group:
    def __init__(self, name: str = "/", reader: Optional[_Reader] = None, header: Optional[int] = None):
        self.name = name
        self._children: Dict[str, Union["Group", Dataset]] = {}
        self._reader = reader
        self._pending: Dict[str, int] = reader.group_links(header) if reader is not None else {}
        self.attrs = AttributeDict(reader.attributes(header) if reader is not None else {})

    def _child_name(self, key: str) -> str:
        return (self.name.rstrip("/") + "/" + key) if self.name != "/" else "/" + key

    def _materialise(self, key: str):
        if key not in self._children and key in self._pending:
            hdr = self._pending[key]
            types = {m[0] for m in self._reader.messages(hdr)}
            if 0x0011 in types or (0x0008 not in types and 0x0003 not in types):
                self._children[key] = Group(self._child_name(key), self._reader, hdr)
            else:
                self._children[key] = Dataset(self._child_name(key), None, self._reader, hdr)
        return self._children[key]

    # -- mapping surface ---------------------------------------------------------------------------
    def keys(self):
        return sorted(set(self._children) | set(self._pending))

    def __iter__(self) -> Iterator[str]:
        return iter(self.keys())

    def __len__(self) -> int:
        return len(self.keys())

    def __contains__(self, key: str) -> bool:
        try:
            self[key]
            return True
        except KeyError:
            return False

    def values(self):
        return [self[k] for k in self.keys()]

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def __getitem__(self, key: str):
        node = self
        for part in [p for p in key.split("/") if p]:
            if not isinstance(node, Group) or (part not in node._children and part not in node._pending):
                raise KeyError(f"Unable to open object (object '{part}' doesn't exist)")
            node = node._materialise(part)
        return node

    # -- writing -----------------------------------------------------------------------------------
    def create_group(self, name: str) -> "Group":
        parts = [p for p in name.split("/") if p]
        node = self
        for i, part in enumerate(parts):
            exists = part in node._children or part in node._pending
            if exists:
                if i == len(parts) - 1:
                    raise ValueError(f"Unable to create group (name already exists): {name}")
                node = node._materialise(part)
            else:
                g = Group(node._child_name(part))
                node._children[part] = g
                node = g
        return node

    def require_group(self, name: str) -> "Group":
        try:
            g = self[name]
        except KeyError:
            return self.create_group(name)
        if not isinstance(g, Group):
            raise TypeError(f"{name} exists and is not a group")
        return g

    def create_dataset(self, name: str, data=None, shape=None, dtype=None, chunks=None, compression=None,
                       compression_opts=None, shuffle=False, **_ignored) -> Dataset:
        """Contiguous dataset, or chunked (+ shuffle / gzip) when `chunks` or `compression="gzip"` is given."""
        parts = [p for p in name.split("/") if p]
        parent = self.require_group("/".join(parts[:-1])) if len(parts) > 1 else self
        leaf = parts[-1]
        if leaf in parent._children or leaf in parent._pending:
            raise ValueError(f"Unable to create dataset (name already exists): {name}")
        if data is None:
            arr = np.zeros(shape if shape is not None else (), dtype=dtype or np.float32)
        else:
            arr = np.asarray(data)
            if dtype is not None:
                arr = arr.astype(dtype)
        if arr.dtype.kind == "U":
            arr = np.char.encode(arr, "utf-8")
        if arr.dtype.kind not in "fiuS" and arr.dtype != np.bool_:
            raise TypeError(f"hdf5io can only write numeric / byte-string arrays, got dtype {arr.dtype}")
        ds = Dataset(parent._child_name(leaf), np.ascontiguousarray(arr))
        if compression not in (None, "gzip"):
            raise ValueError("only gzip compression is supported")
        if (chunks or compression or shuffle) and arr.ndim >= 1 and arr.size:
            if chunks is None or chunks is True:       # ~128 KiB chunks along the first axis
                per_row = max(int(np.prod(arr.shape[1:])) * arr.dtype.itemsize, 1)
                chunks = (max(1, min(arr.shape[0], (1 << 17) // per_row)),) + tuple(arr.shape[1:])
            level = (4 if compression_opts is None else int(compression_opts)) if compression == "gzip" else None
            ds._storage = {"chunks": tuple(int(c) for c in chunks), "level": level, "shuffle": bool(shuffle)}
        parent._children[leaf] = ds
        return ds

    def _load_all(self) -> None:
        """Pull every object below this group into memory (before an 'a'-mode rewrite)."""
        for k in self.keys():
            c = self[k]
            if isinstance(c, Group):
                c._load_all()
            else:
                c._load()


# =================================================================================================
# writer
# =================================================================================================
def _dtype_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt == np.bool_:
        dt = np.dtype("u1")
    if dt.kind == "f":
        props = {2: (15, 10, 5, 0, 10, 15), 4: (31, 23, 8, 0, 23, 127), 8: (63, 52, 11, 0, 52, 1023)}[dt.itemsize]
        sign, eloc, esize, mloc, msize, bias = props
        return (struct.pack("<BBBBI", 0x11, 0x20, sign, 0, dt.itemsize)
                + struct.pack("<HHBBBBI", 0, dt.itemsize * 8, eloc, esize, mloc, msize, bias))
    if dt.kind in "iu":
        return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize) + \
            struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, max(dt.itemsize, 1))   # null-padded ASCII, as h5py writes numpy 'S'
    raise TypeError(f"cannot write dtype {dt}")


def _dataspace_message(shape: Tuple[int, ...]) -> bytes:
    if len(shape) == 0:
        return struct.pack("<BBBBI", 1, 0, 0, 0, 0)
    return struct.pack("<BBBBI", 1, len(shape), 1, 0, 0) + b"".join(struct.pack("<Q", s) for s in shape) * 2


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _message(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHBBBB", mtype, len(data), flags, 0, 0, 0) + data


def _attr_value(v) -> np.ndarray:
    if isinstance(v, str):
        return np.array(v.encode("utf-8"))
    if isinstance(v, bytes):
        return np.array(v)
    a = np.asarray(v)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "utf-8")
    if a.dtype == np.bool_:
        a = a.astype("u1")
    if a.dtype.kind not in "fiuS":
        raise TypeError(f"cannot write attribute of dtype {a.dtype}")
    if a.dtype.byteorder == ">":
        a = a.astype(a.dtype.newbyteorder("<"))
    return a


def _attribute_messages(attrs: Dict[str, object]) -> List[bytes]:
    out = []
    for name, v in attrs.items():
        a = _attr_value(v)
        nm = name.encode("utf-8") + b"\x00"
        dt = _dtype_message(a.dtype)
        sp = _dataspace_message(a.shape)
        body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + \
            np.ascontiguousarray(a).tobytes()
        out.append(_message(0x000C, body))
    return out


def _object_header(messages: List[bytes]) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\x00" * 4 + body


class _Writer:
    def __init__(self, leaf_k: int, internal_k: int = 16):
        self.buf = bytearray(96)                       # superblock placeholder
        self.leaf_k, self.internal_k = leaf_k, internal_k

    def alloc(self, data: bytes, align: int = 8) -> int:
        pad = -len(self.buf) % align
        self.buf += b"\x00" * pad
        addr = len(self.buf)
        self.buf += data
        return addr

    def write_dataset(self, ds: Dataset) -> int:
        arr = np.ascontiguousarray(ds._load())
        if arr.dtype == np.bool_:
            arr = arr.astype("u1")
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        raw = arr.tobytes()
        msgs = [_message(0x0001, _dataspace_message(arr.shape)),
                _message(0x0003, _dtype_message(arr.dtype), flags=1),
                _message(0x0005, struct.pack("<BBBB", 2, 2, 2, 1) + struct.pack("<I", 0), flags=1)]
        opts = getattr(ds, "_storage", None)
        if opts and len(raw) and arr.ndim >= 1:
            msgs += self._chunked_storage(arr, opts)
        else:
            msgs.append(_message(0x0008, struct.pack("<BBQQ", 3, 1, self.alloc(raw) if len(raw) else UNDEF, len(raw))))
        msgs += _attribute_messages(ds.attrs)
        return self.alloc(_object_header(msgs))

    def _chunked_storage(self, arr: np.ndarray, opts: dict) -> List[bytes]:
        """Chunked layout with (optional) shuffle + deflate, the storage pycbc's TimeSeries.save uses
        (`compression='gzip', compression_opts=9, shuffle=True`): filter-pipeline + layout messages."""
        rank, isz = arr.ndim, arr.dtype.itemsize
        chunk = tuple(int(min(c, s)) for c, s in zip(opts["chunks"], arr.shape))
        level, shuffle = opts.get("level"), opts.get("shuffle", False)
        grid = [range(0, s, c) for s, c in zip(arr.shape, chunk)]
        entries = []                                   # (offsets, address, stored size)
        import itertools
        for offs in itertools.product(*grid):
            block = np.zeros(chunk, dtype=arr.dtype)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, arr.shape))
            block[tuple(slice(0, x.stop - x.start) for x in sl)] = arr[sl]
            raw = block.tobytes()
            if shuffle and isz > 1:
                raw = np.frombuffer(raw, dtype=np.uint8).reshape(-1, isz).T.tobytes()
            if level is not None:
                raw = zlib.compress(raw, level)
            entries.append((offs, self.alloc(raw), len(raw)))
        K = 32                                         # default indexed-storage internal node K of a v0 superblock

        def key(size, offs):
            return struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", 0)

        def node(level_no, items, end_offs):
            body = bytearray(b"TREE" + struct.pack("<BBH", 1, level_no, len(items)) + struct.pack("<QQ", UNDEF, UNDEF))
            for offs, addr, size in items:
                body += key(size, offs) + struct.pack("<Q", addr)
            body += key(0, end_offs)
            body += b"\x00" * (24 + (2 * K + 1) * (8 + 8 * (rank + 1)) + 2 * K * 8 - len(body))
            return self.alloc(bytes(body))

        end = tuple(((s + c - 1) // c) * c for s, c in zip(arr.shape, chunk))
        level_no, items = 0, entries
        while True:
            groups = [items[i:i + 2 * K] for i in range(0, len(items), 2 * K)]
            nodes = []
            for gi, grp in enumerate(groups):
                nxt = groups[gi + 1][0][0] if gi + 1 < len(groups) else end
                nodes.append((grp[0][0], node(level_no, grp, nxt), grp[0][2]))
            if len(nodes) == 1:
                root = nodes[0][1]
                break
            items, level_no = nodes, level_no + 1
        filt = []
        if shuffle and isz > 1:
            filt.append(struct.pack("<HHHH", 2, 0, 1, 1) + struct.pack("<I", isz) + b"\x00" * 4)
        if level is not None:
            filt.append(struct.pack("<HHHH", 1, 0, 1, 1) + struct.pack("<I", level) + b"\x00" * 4)
        msgs = []
        if filt:
            msgs.append(_message(0x000B, struct.pack("<BBHI", 1, len(filt), 0, 0) + b"".join(filt)))
        lay = struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", root) + \
            b"".join(struct.pack("<I", c) for c in chunk) + struct.pack("<I", isz)
        msgs.append(_message(0x0008, lay))
        return msgs

    def write_group(self, g: Group) -> Tuple[int, int, int]:
        """-> (object header address, B-tree address, local heap address)"""
        names = sorted(g.keys(), key=lambda s: s.encode("utf-8"))
        child_hdr = {}
        for k in names:
            c = g[k]
            child_hdr[k] = self.write_group(c)[0] if isinstance(c, Group) else self.write_dataset(c)
        # local heap: offset 0 = empty string, then the names, then one free block
        heap = bytearray(8)
        name_off = {}
        for k in names:
            name_off[k] = len(heap)
            heap += _pad8(k.encode("utf-8") + b"\x00")
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 32) + b"\x00" * 16          # free block: next = H5HL_FREE_NULL, size 32
        heap_data_addr = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<BBBB", 0, 0, 0, 0) +
                               struct.pack("<QQQ", len(heap), free_off, heap_data_addr))
        # symbol nodes: up to 2*leaf_k entries each
        cap = 2 * self.leaf_k
        snods = []
        for i in range(0, max(len(names), 1), cap):
            part = names[i:i + cap]
            body = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(part)))
            for k in part:
                body += struct.pack("<QQII", name_off[k], child_hdr[k], 0, 0) + b"\x00" * 16
            body += b"\x00" * (40 * (cap - len(part)))
            snods.append((self.alloc(bytes(body)), name_off[part[-1]] if part else 0))
        if len(snods) > 2 * self.internal_k:
            raise OSError("too many entries for a single-level group B-tree")
        node = bytearray(b"TREE" + struct.pack("<BBH", 0, 0, len(snods) if names else 0) + struct.pack("<QQ", UNDEF, UNDEF))
        node += struct.pack("<Q", 0)
        for addr, last in (snods if names else []):
            node += struct.pack("<QQ", addr, last)
        node += b"\x00" * (24 + 8 + 16 * 2 * self.internal_k - len(node))
        btree_addr = self.alloc(bytes(node))
        msgs = [_message(0x0011, struct.pack("<QQ", btree_addr, heap_addr))] + _attribute_messages(g.attrs)
        return self.alloc(_object_header(msgs)), btree_addr, heap_addr

    def finish(self, root: Group) -> bytes:
        hdr, btree, heap = self.write_group(root)
        eof = len(self.buf)
        sb = SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + \
            struct.pack("<HHI", self.leaf_k, self.internal_k, 0) + \
            struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF) + \
            struct.pack("<QQII", 0, hdr, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def _max_fanout(g: Group) -> int:
    m = len(g.keys())
    for k in g.keys():
        c = g[k]
        if isinstance(c, Group):
            m = max(m, _max_fanout(c))
    return m


class File(Group):
    """h5py.File look-alike for modes 'r', 'w', 'a' (and 'r+' == 'a')."""

    def __init__(self, path: str, mode: str = "r"):
        self.filename = os.fspath(path)
        self.mode = "a" if mode == "r+" else mode
        self._closed = False
        if self.mode not in ("r", "w", "a", "w-", "x"):
            raise ValueError(f"unsupported mode {mode!r}")
        if self.mode in ("w-", "x") and os.path.exists(self.filename):
            raise FileExistsError(self.filename)
        if self.mode == "r" or (self.mode == "a" and os.path.exists(self.filename)):
            with open(self.filename, "rb") as fh:
                rd = _Reader(fh.read())
            super().__init__("/", rd, rd.root_header)
            if self.mode == "a":
                self._load_all()
        else:
            super().__init__("/")

    def flush(self) -> None:
        if self.mode == "r" or self._closed:
            return
        # every group gets one symbol node: choose the leaf K of the file accordingly (32 symbol nodes of 2K
        # entries fit one B-tree node, so this also covers groups with many segments)
        n = max(_max_fanout(self), 1)
        leaf_k = 4
        while 2 * leaf_k * 16 < n:
            leaf_k *= 2
        if leaf_k > 32767:
            raise OSError("group too large for this writer")
        data = _Writer(leaf_k).finish(self)
        tmp = self.filename + ".tmp~"
        with open(tmp, "wb") as fh:
            fh.write(data)
        os.replace(tmp, self.filename)

    def close(self) -> None:
        if not self._closed:
            self.flush()
            self._closed = True

    def __enter__(self) -> "File":
        return self

    def __exit__(self, *exc) -> None:
        self.close()


def open_file(path: str, mode: str = "r"):
    """h5py.File when h5py is importable (the reference's own I/O), else this module's File."""
    try:
        import h5py  # type: ignore
        return h5py.File(path, mode)
    except ImportError:
        return File(path, mode)
