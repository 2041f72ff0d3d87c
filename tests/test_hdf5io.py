"""hdf5io (pure-Python HDF5 subset) against (i) the one h5py-written file the reference
ships, (ii) its own writer, in the MLGWSC-1 strain layout `file[det][str(int(start))]` with `start_time` /
`delta_t` attributes and pycbc's chunked + shuffle + gzip storage (MLGWSC-1/generate_data.py:197-216), and
(iii) the structural layout of that shipped file (what libhdf5 itself writes for libver='earliest')."""
import os
import struct

import numpy as np
import pytest

from gw_whisper_b200 import hdf5io as H

REF_FILE = "/root/reference/Signal_vs_Noise/results/Real_events/results_2_detectors_real_events.hdf"
have_ref = os.path.exists(REF_FILE)


@pytest.mark.skipif(not have_ref, reason="reference tree not mounted")
def test_reads_the_references_h5py_written_file():
    with H.File(REF_FILE, "r") as f:
        assert sorted(f.keys()) == ["event_names", "model_output"]
        out = f["model_output"]
        assert out.shape == (70, 1) and out.dtype == np.float32
        v = out[()]
        assert np.all((v >= 0) & (v <= 1))                       # sigmoid outputs (evaluation_real_events.py:60)
        names = f["event_names"][()]                             # variable-length strings via the global heap
        assert names.shape == (70,) and names[0] == "GW190403_051519"
        assert all(n.startswith("GW") for n in names)


def _strain_file(path, rng, compressed):
    segs = {}
    with H.File(path, "w") as f:
        for det in ("H1", "L1"):
            g = f.create_group(det)
            for st, n in ((1238166018, 70001), (1238170000, 4096), (1238180000, 300000)):
                x = rng.standard_normal(n)
                kw = dict(compression="gzip", compression_opts=9, shuffle=True) if compressed else {}
                ds = g.create_dataset(str(st), data=x, **kw)
                ds.attrs["start_time"] = float(st)
                ds.attrs["delta_t"] = 1.0 / 2048
                segs[(det, str(st))] = x
    return segs


@pytest.mark.parametrize("compressed", [False, True])
def test_mlgwsc_strain_layout_round_trip(tmp_path, compressed):
    p = str(tmp_path / "strain.hdf")
    segs = _strain_file(p, np.random.default_rng(1), compressed)
    with H.File(p, "r") as f:
        assert list(f.keys()) == ["H1", "L1"]
        det_grp = next(iter(f.values()))                         # inference.py:534
        assert sorted(det_grp.keys()) == ["1238166018", "1238170000", "1238180000"]
        for (det, key), x in segs.items():
            ds = f[det][key]
            assert len(ds) == len(x) and ds.dtype == np.float64 and ds.ndim == 1
            assert np.array_equal(ds[()], x)                     # bit-exact through shuffle + deflate
            assert np.array_equal(ds[100:200], x[100:200])
            st = ds.attrs["start_time"]
            assert isinstance(st, np.float64) and st == float(key)   # what h5py returns (f64 times)
            assert ds.attrs["delta_t"] == 1.0 / 2048
        assert np.array_equal(f["H1/1238170000"][()], segs[("H1", "1238170000")])
        with pytest.raises(KeyError):
            f["V1"]
    if compressed:
        assert os.path.getsize(p) < 0.97 * sum(len(x) * 8 for x in segs.values())


def test_trigger_output_and_append_mode(tmp_path):
    """The four datasets inference.py:667-672 writes, then the debug-file pattern (:225-227): open 'a',
    require_group, create_dataset."""
    p = str(tmp_path / "out.hdf")
    t = np.array([10.5, 20.25]); s = np.array([0.9, 0.7]); v = np.array([0.2, 0.2])
    av = np.arange(9, dtype=np.float32)
    with H.File(p, "w") as f:
        f.create_dataset("time", data=t); f.create_dataset("stat", data=s)
        f.create_dataset("var", data=v); f.create_dataset("all_vals", data=av)
        f.create_dataset("empty", data=np.array([], dtype=np.float32))
    with H.File(p, "a") as f:
        f.require_group("H1").create_dataset("77", data=np.ones(5))
        f.require_group("H1").create_dataset("78", data=np.zeros((2, 3), dtype=np.int32))
        with pytest.raises(ValueError):
            f.create_dataset("time", data=t)
    with H.File(p) as f:
        assert np.array_equal(f["time"][()], t) and np.array_equal(f["stat"][()], s)
        assert np.array_equal(f["var"][()], v) and f["all_vals"].dtype == np.float32
        assert f["empty"].shape == (0,) and f["empty"][()].size == 0
        assert np.array_equal(f["H1"]["77"][()], np.ones(5))
        assert f["H1"]["78"].shape == (2, 3) and f["H1"]["78"].dtype == np.int32


def test_many_segments_in_one_group(tmp_path):
    p = str(tmp_path / "many.hdf")
    with H.File(p, "w") as f:
        g = f.create_group("H1")
        for i in range(700):
            g.create_dataset(str(1000000 + 37 * i), data=np.full(3, i, dtype=np.float32))
    with H.File(p) as f:
        keys = f["H1"].keys()
        assert len(keys) == 700
        assert f["H1"]["1000370"][()].tolist() == [10.0, 10.0, 10.0]


@pytest.mark.skipif(not have_ref, reason="reference tree not mounted")
def test_writer_layout_matches_libhdf5_structures(tmp_path):
    """Byte-level comparison of the structures both files must share: superblock fields, root symbol-table
    entry, object-header prefix, the message encodings of a float32 [70,1] dataset."""
    p = str(tmp_path / "cmp.hdf")
    with H.File(REF_FILE) as f:
        data = f["model_output"][()]
    with H.File(p, "w") as f:
        f.create_dataset("model_output", data=data)
    ref = open(REF_FILE, "rb").read()
    got = open(p, "rb").read()
    assert got[:16] == ref[:16]                                  # signature, versions, sizes of offsets/lengths
    assert got[18:24] == ref[18:24]                              # internal K, consistency flags
    assert got[24:40] == ref[24:40] and got[48:56] == ref[48:56]  # base address, free-space, driver-info addresses
    assert struct.unpack_from("<Q", got, 40)[0] == len(got)      # end-of-file address
    assert struct.unpack_from("<I", got, 72)[0] == struct.unpack_from("<I", ref, 72)[0] == 1   # cached group entry
    rd_ref, rd_got = H._Reader(ref), H._Reader(got)

    def msgs(rd, name):
        hdr = rd.group_links(rd.root_header)[name]
        return {t: (fl, d) for t, fl, d in rd.messages(hdr)}
    a, b = msgs(rd_ref, "model_output"), msgs(rd_got, "model_output")
    for mtype in (0x0001, 0x0003, 0x0005):                       # dataspace, datatype, fill value: identical bytes
        assert a[mtype] == b[mtype], hex(mtype)
    assert a[0x0008][1][:2] == b[0x0008][1][:2]                  # layout version 3, contiguous
    assert struct.unpack_from("<Q", a[0x0008][1], 10) == struct.unpack_from("<Q", b[0x0008][1], 10)   # byte size
    # local heap: same header layout, free-list terminator H5HL_FREE_NULL == 1
    for rd, buf in ((rd_ref, ref), (rd_got, got)):
        _, heap_addr = struct.unpack_from("<QQ", [d for t, _, d in rd.messages(rd.root_header) if t == 0x11][0], 0)
        assert buf[heap_addr:heap_addr + 8] == b"HEAP\x00\x00\x00\x00"
        size, free_off, daddr = struct.unpack_from("<QQQ", buf, heap_addr + 8)
        nxt, fsz = struct.unpack_from("<QQ", buf, daddr + free_off)
        assert nxt == 1 and free_off + fsz == size
