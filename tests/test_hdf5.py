"""The package's own HDF5 encoder/decoder (sim_juncs_b200/hdf5.py).

Anchor: tests/golden/libhdf5_written_testhdf5_7.4_GLNX86.mat is an HDF5 file written by the real library
(MATLAB 7.4 through libhdf5; copied from scipy/io/matlab/tests/data, BSD-licensed test data; the HDF5
superblock sits behind a 512-byte MATLAB header).  The reader must decode it, and the writer's encodings are
compared byte for byte with the structures inside it.
"""
import os
import struct
import types

import numpy as np
import pytest

from sim_juncs_b200 import hdf5
from sim_juncs_b200.output import FIELD_TYPE, LOC_TYPE, SRC_TYPE, field_samples_dict, write_field_samples_h5

GOLD = os.path.join(os.path.dirname(__file__), "golden", "libhdf5_written_testhdf5_7.4_GLNX86.mat")


def test_reader_decodes_a_file_written_by_libhdf5():
    f = hdf5.File(GOLD)
    assert f.keys() == ["testdouble"] and (f.leaf_k, f.internal_k, f.base) == (4, 16, 512)
    ds = f["testdouble"]
    assert ds.shape == (9, 1) and ds.dtype == np.dtype("<f8")
    # scipy/io/matlab/tests/test_mio.py: testdouble = 0 : pi/4 : 2 pi
    assert np.array_equal(ds.read().ravel(), np.arange(9) * (np.pi / 4))


def test_writer_encodings_equal_libhdf5_bytes():
    raw = open(GOLD, "rb").read()
    # float64 datatype message body at 1528 of the real file (message header at 1520, size 0x18)
    assert hdf5.encode_dtype(np.float64) == raw[1528:1548]
    # fill value message body the writer emits is the one libhdf5 wrote at 1512
    assert struct.pack("<BBBBI", 1, 2, 2, 1, 0) == raw[1512:1520]
    path = "/tmp/_sj_h5_bytes.h5"
    w = hdf5.H5Writer(path)
    w.create_dataset("testdouble", (np.arange(9) * (np.pi / 4)).reshape(9, 1))
    w.close()
    mine = open(path, "rb").read()
    os.remove(path)
    # superblock: signature, versions, sizes, K values equal; then addresses differ by layout only
    assert mine[:20] == raw[512:532]
    # local heap: same header layout, name at offset 8, free block (next = 1, size) closes the segment
    h_m, h_r = mine.index(b"HEAP"), raw.index(b"HEAP")
    assert mine[h_m:h_m + 8] == raw[h_r:h_r + 8]
    fl_m, fl_r = struct.unpack_from("<Q", mine, h_m + 16)[0], struct.unpack_from("<Q", raw, h_r + 16)[0]
    assert fl_m == fl_r == 24
    assert mine[h_m + 32:h_m + 32 + 24] == raw[h_r + 32:h_r + 32 + 24]            # "", "testdouble"
    assert struct.unpack_from("<Q", mine, h_m + 32 + 24)[0] == struct.unpack_from("<Q", raw, h_r + 32 + 24)[0] == 1
    # B-tree node: identical 24-byte header, key 0, and key 1 (heap offset of the only name); same node size
    t_m, t_r = mine.index(b"TREE"), raw.index(b"TREE")
    assert mine[t_m:t_m + 32] == raw[t_r:t_r + 32] and mine[t_m + 40:t_m + 48] == raw[t_r + 40:t_r + 48]
    s_m, s_r = mine.index(b"SNOD"), raw.index(b"SNOD")
    assert mine[s_m:s_m + 16] == raw[s_r:s_r + 16]                                 # version, count, name offset
    # the symbol node is allocated at its full size (8 + 2 K 40 bytes), as libhdf5 expects when reading it
    assert len(mine) >= s_m + 8 + 8 * 40 or mine.index(b"TREE") > s_m
    # group object header: one symbol-table message (type 0x11, size 16, flags of the real file)
    o_r = 1440
    assert raw[o_r + 16:o_r + 20] == struct.pack("<HH", 0x11, 16)
    root_m = struct.unpack_from("<Q", mine, 64)[0]
    assert mine[root_m + 16:root_m + 20] == struct.pack("<HH", 0x11, 16) and mine[root_m] == raw[o_r] == 1


def test_compound_datatype_message_layout():
    """Version-1 compound encoding (HDF5 File Format Specification, IV.A.2.d, class 6): per member the name padded
    to 8 bytes, byte offset, rank byte + 3 reserved, permutation, reserved, four dimension sizes, member type."""
    b = hdf5.encode_dtype(FIELD_TYPE)
    f64 = hdf5.encode_dtype(np.float64)
    assert b[:8] == bytes([0x16, 2, 0, 0]) + struct.pack("<I", 16)
    member = lambda name, off: name.ljust(8, b"\0") + struct.pack("<I", off) + bytes(28) + f64
    assert b[8:] == member(b"Re", 0) + member(b"Im", 8)
    assert len(hdf5.encode_dtype(LOC_TYPE)) == 8 + 3 * 60
    s = hdf5.encode_dtype(SRC_TYPE)
    assert s[:8] == bytes([0x16, 6, 0, 0]) + struct.pack("<I", 56)                 # sizeof(source_info), disp.hpp:96-109
    assert s[8:24] == b"wavelen\0" + struct.pack("<I", 8) + bytes(4)
    assert b"start_time\0\0\0\0\0\0" + struct.pack("<I", 32) in s and b"amplitude\0\0\0\0\0\0\0" + struct.pack("<I", 48) in s
    for dt in (FIELD_TYPE, LOC_TYPE, SRC_TYPE, np.dtype("<u8"), np.dtype("<i4"), np.dtype("<f4")):
        back, end = hdf5.decode_dtype(hdf5.encode_dtype(dt), 0)
        assert back == dt and end == len(hdf5.encode_dtype(dt))
    assert hdf5.encode_dtype(np.uint64) == bytes([0x10, 0, 0, 0, 8, 0, 0, 0, 0, 0, 64, 0])   # NATIVE_HSIZE


@pytest.mark.parametrize("n_points", [1, 7, 40, 700])
def test_round_trip_groups_of_every_size(tmp_path, n_points):
    rng = np.random.default_rng(n_points)
    path = str(tmp_path / "t.h5")
    w = hdf5.H5Writer(path)
    want = {}
    for i in range(n_points):
        a = np.zeros(11, dtype=FIELD_TYPE)
        a["Re"], a["Im"] = rng.standard_normal(11), rng.standard_normal(11)
        want["c/point_%04d/time" % i] = a
        w.create_dataset("c/point_%04d/time" % i, a)
    w.create_dataset("c/locations", np.zeros(0, dtype=LOC_TYPE))
    w.create_group("empty")
    w.create_dataset("n", np.array([n_points], dtype=np.uint64))
    w.close()
    f = hdf5.File(path)
    assert f.keys() == ["c", "empty", "n"] and len(f["empty"]) == 0 and f["n"][0] == n_points
    keys = f["c"].keys()
    assert keys[0] == "locations" and keys[1:] == sorted(k.split("/")[1] for k in want) and len(f["c"]) == n_points + 1
    for k, a in want.items():
        got = f[k].read()
        assert got.dtype == FIELD_TYPE and np.array_equal(got["Re"], a["Re"]) and np.array_equal(got["Im"], a["Im"])
    # every structure lies inside the end-of-file address of the superblock, which equals the file size
    assert f.eof == os.path.getsize(path)
    assert f.leaf_k == max(4, -(-(n_points + 1) // 64))


def test_field_samples_file_replays_the_reference_scripts_reads(tmp_path):
    """field_samples.h5 for a synthetic bound_geom state, read the way scripts/phases.py:684-733 reads it."""
    rng = np.random.default_rng(3)
    n_mon, n_saves = 12, 175
    bg = types.SimpleNamespace(
        monitor_clusters=[4, 12], monitor_locs=rng.uniform(0, 18, (n_mon, 3)), n_t_pts=3500, save_span=20,
        sources=[types.SimpleNamespace(wavelen=0.76, width=1.27, phase=0.25, start_time=5.0, end_time=20.2, amplitude=1.0)],
        problem=types.SimpleNamespace(cgs_params=[("pi", np.pi), ("length", 18.0)]),
        field_times=[rng.standard_normal(n_saves) + 1j * rng.standard_normal(n_saves) for _ in range(n_mon)],
        time_bounds=lambda: [0.0, 300.0, 300.0 * 20 / 3500])
    d = field_samples_dict(bg)
    path = write_field_samples_h5(d, str(tmp_path / "field_samples.h5"))
    f = hdf5.File(path)
    clust = [k for k in f.keys() if "cluster" in k and len(f[k]) > 1 and len(f[k]["locations"]) > 0]
    assert clust == ["cluster_0", "cluster_1"] and "cluster_2" in f.keys()          # the empty trailing cluster
    keylist = list(f[clust[0]].keys())
    assert len(f[clust[0]][keylist[1]]["time"]) == n_saves
    assert f["info"]["time_bounds"][1] == 300.0 and f["info"]["sources"]["wavelen"][0] == 0.76
    assert f["info"]["sources"][0]["end_time"] == 20.2 and f["info"]["sources"].dtype.itemsize == 56
    assert f["info"]["n_clusters"][0] == 2 and f["info"]["n_clusters"].dtype == np.dtype("<u8")
    assert f["info"]["cgs_params"]["length"][0] == 18.0
    pts = list(f[clust[1]].keys())[1:]
    assert pts == ["point_%02d" % i for i in range(4, 12)]
    assert np.array_equal(np.array(f[clust[1]][pts[0]]["time"]["Re"]), bg.field_times[4].real)
    assert f[clust[1]]["locations"]["z"][-1] == bg.monitor_locs[11, 2]
    assert len(f[clust[0]][keylist[1]]["frequency"]) == 128


@pytest.mark.parametrize("n_points", [3, 300])
def test_cpp_writer_is_byte_identical_to_the_python_writer(tmp_path, n_points):
    """host/sj_hdf5.hpp (used by the C++ sim_geom host) and hdf5.py lay the same content out identically."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "h5w")
    subprocess.check_call(["g++", "-O1", "-std=c++11", "-I", os.path.join(root, "host"), "-o", exe,
                           os.path.join(root, "tests", "cpp", "h5_writer_main.cpp")])
    cpp_path, py_path = str(tmp_path / "cpp.h5"), str(tmp_path / "py.h5")
    subprocess.check_call([exe, cpp_path, str(n_points)])
    w = hdf5.H5Writer(py_path)
    w.create_dataset("info/time_bounds", np.array([0.0, 300.0, 1.5]))
    w.create_dataset("info/n_clusters", np.array([1], dtype=np.uint64))
    src = np.zeros(1, dtype=SRC_TYPE)
    src[0] = (0.76, 1.27, 0.25, 5.0, 20.2, 1.0)
    w.create_dataset("info/sources", src)
    w.create_group("info/cgs_params")
    locs = np.zeros(n_points, dtype=LOC_TYPE)
    flat = 0.25 * np.arange(3 * n_points).reshape(-1, 3)
    locs["x"], locs["y"], locs["z"] = flat[:, 0], flat[:, 1], flat[:, 2]
    w.create_dataset("cluster_0/locations", locs)
    w.create_dataset("cluster_1/locations", np.zeros(0, dtype=LOC_TYPE))
    for i in range(n_points):
        t = np.zeros(9, dtype=FIELD_TYPE)
        t["Re"], t["Im"] = i + 0.5 * np.arange(9), 0.0 - np.arange(9.0)
        w.create_dataset("cluster_0/point_%04d/time" % i, t)
    w.close()
    a, b = open(cpp_path, "rb").read(), open(py_path, "rb").read()
    assert len(a) == len(b) and a == b
    f = hdf5.File(cpp_path)
    assert f["info"]["sources"][0]["end_time"] == 20.2 and len(f["cluster_0"]) == n_points + 1
