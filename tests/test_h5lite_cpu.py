"""The HDF5 subset reader / writer behind `spin.read_save` / `spin.write_save` (`rbqoc.jl:98-139`, keys of
`spin13.jl:196-221`).  Pins: a file written by libhdf5 itself (scipy ships one MATLAB 7.3 fixture = HDF5 with a 512-byte user
block), write -> read round trips of the reference's result dictionary, and hand-assembled files for the layouts h5py users
produce (chunked + gzip + shuffle, variable-length strings, version-2 object headers with link messages)."""
import os
import struct
import zlib

import numpy as np
import pytest

from rbqoc_b200 import h5lite, spin


def test_reads_a_file_written_by_libhdf5():
    import scipy.io
    path = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("scipy's MATLAB 7.3 fixture is not installed")
    f = h5lite.File(path)
    assert f.base == 512 and f.keys() == ["testdouble"]
    v = f["testdouble"]
    assert v.dtype == np.float64 and v.shape == (9, 1)
    assert np.array_equal(v[:, 0], np.linspace(0.0, 2.0 * np.pi, 9))


def _reference_result(N=41, n=11):
    rng = np.random.default_rng(3)
    return {                                                   # spin13.jl:196-221, shapes as Julia holds them
        "acontrols": rng.normal(size=(N - 1, 1)), "controls_idx": np.array([10]), "d2controls_dt2_idx": np.array([1]),
        "evolution_time": 4.0, "astates": rng.normal(size=(N, n)), "Q": rng.random(n), "Qf": rng.random(n), "R": rng.random(1),
        "cmax": 1.2e-9, "dt": 0.1, "solver_type": 1, "sqrtbp": 0, "max_penalty": 1e11, "constraint_tol": 1e-8, "al_tol": 1e-4,
        "gate_type": 2, "save_type": 1, "integrator_type": 3, "iterations": 153, "max_iterations": 10000, "pn_steps": 2,
        "max_cost_value": 1e8, "benchmark_result": None, "note": "spin13 π/2",
    }


def test_round_trip_of_the_reference_result_dictionary(tmp_path):
    res = _reference_result()
    path = tmp_path / "00000_spin13.h5"
    spin.write_save(path, res)
    back = spin.read_save(path)
    assert sorted(back) == sorted(k for k, v in res.items() if v is not None)       # 23 entries: three symbol-table nodes
    for k, v in res.items():
        if v is None:
            continue
        if isinstance(v, str):
            assert back[k] == v
        else:
            assert np.array_equal(np.asarray(back[k]), np.asarray(v)), k
            assert np.asarray(back[k]).shape == np.asarray(v).shape, k
    assert isinstance(back["iterations"], (int, np.integer)) and isinstance(back["dt"], (float, np.floating))
    # what the file itself holds is the transposed array (HDF5.jl reverses Julia's dimensions), as h5py would show it
    raw = h5lite.File(path)["astates"]
    assert raw.shape == (11, 41) and np.array_equal(raw.T, res["astates"])
    # grab_controls / sample_controls straight from the path, for the three save formats (rbqoc.jl:98-123)
    c, dt_inv, T = spin.grab_controls(path)
    assert np.array_equal(c[:, 0], res["astates"][:-1, 9]) and dt_inv == pytest.approx(10.0) and T == 4.0
    cs, ds, Tf = spin.sample_controls(path)
    assert cs.shape == (400, 1) and Tf == pytest.approx(4.0)
    p2 = tmp_path / "sample.h5"
    spin.write_save(p2, {"save_type": 2, "controls_sample": cs, "d2controls_dt2_sample": ds, "evolution_time_sample": Tf})
    c2, dti2, T2 = spin.grab_controls(p2)
    assert np.array_equal(c2, cs[:-1]) and dti2 == 100.0 and T2 == Tf
    # the python format (spin14.py:222-235): h5py stores `controls` (N, 1) row-major, Julia reads it as (1, N)
    p3 = tmp_path / "py.h5"
    h5lite.write(p3, {"controls": cs, "save_type": 3, "dt": 1e-2, "evolution_time": Tf})
    c3, _, _ = spin.grab_controls(p3)
    assert np.array_equal(c3, cs)


def test_round_trip_types_groups_and_user_block(tmp_path):
    entries = {"f8": np.arange(24.0).reshape(2, 3, 4), "f4": np.float32([1.5, -2.5]), "i4": np.int32([[1, -2], [3, 4]]),
               "u1": np.uint8([0, 255]), "b": np.array([True, False]), "c16": np.array([1 + 2j, 3 - 4j]), "empty": np.zeros((0,)),
               "scalar_i": 7, "scalar_f": -0.25, "s": "gate → X/2", "grp": {"inner": np.arange(5), "deeper": {"x": 1.0}},
               **{f"key{i:03d}": np.full(3, float(i)) for i in range(40)}}
    data = h5lite.write(None, entries)
    for blob in (data, h5lite.write(None, entries, userblock=1024)):
        f = h5lite.File(blob)
        assert len(f.keys()) == len(entries)
        assert np.array_equal(f["f8"], entries["f8"]) and f["f4"].dtype == np.float32 and f["i4"].dtype == np.int32
        assert np.array_equal(f["i4"], entries["i4"]) and np.array_equal(f["u1"], entries["u1"])
        assert np.array_equal(f["b"], [1, 0]) and f["empty"].shape == (0,)
        c = f["c16"]
        assert np.array_equal(c["r"] + 1j * c["i"], entries["c16"])
        assert f["scalar_i"] == 7 and f["scalar_f"] == -0.25 and f["s"] == entries["s"]
        assert np.array_equal(f["grp/inner"], np.arange(5)) and f["grp"]["deeper"]["x"] == 1.0
        assert all(np.array_equal(f[f"key{i:03d}"], np.full(3, float(i))) for i in range(40))
        with pytest.raises(KeyError):
            f["absent"]
    assert h5lite.File(h5lite.write(None, entries, userblock=1024)).base == 1024
    with pytest.raises(h5lite.H5FormatError):
        h5lite.File(b"not an hdf5 file" * 100)
    with pytest.raises(TypeError):
        h5lite.write(None, {"o": np.array([object()])})
    with pytest.raises(h5lite.H5FormatError):                  # a file cut off in the middle of its metadata
        h5lite.File(data[:len(data) // 2])["f8"]


# ---- hand-assembled files for the layouts the writer never produces -----------------------------------------------------
def _shuffle(raw, es):
    return np.frombuffer(raw, np.uint8).reshape(-1, es).T.tobytes()


def _classic_file_with(dataset_header_builder):
    """Superblock 0 + root symbol table holding one dataset `d` whose header comes from the builder(img)."""
    img = h5lite._Image()
    img.b += b"\x00" * 96
    hdr = img.alloc(dataset_header_builder(img))
    heap = bytearray(b"\x00" * 8) + b"d\x00" + b"\x00" * 6 + struct.pack("<QQ", 1, 16)
    heap_data = img.alloc(bytes(heap))
    heap_addr = img.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 16, heap_data))
    snod = img.alloc(b"SNOD" + struct.pack("<BBH", 1, 0, 1) + struct.pack("<QQII16x", 8, hdr, 0, 0) + b"\x00" * 280)
    node = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, h5lite.UNDEF, h5lite.UNDEF) + struct.pack("<QQQ", 0, snod, 8)
    btree = img.alloc(node + b"\x00" * (544 - len(node)))
    root = img.alloc(h5lite._object_header([h5lite._message(0x11, struct.pack("<QQ", btree, heap_addr))]))
    sb = h5lite.SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, h5lite.UNDEF, len(img.b), h5lite.UNDEF) + struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", btree, heap_addr)
    img.b[:96] = sb
    return bytes(img.b)


def test_chunked_gzip_shuffle_dataset():
    a = np.arange(7 * 10, dtype=np.float64).reshape(7, 10) * 0.5
    cd = (4, 4)

    def build(img):
        keys = []
        for i0 in range(0, 7, 4):
            for j0 in range(0, 10, 4):
                chunk = np.zeros(cd)
                blk = a[i0:i0 + 4, j0:j0 + 4]
                chunk[:blk.shape[0], :blk.shape[1]] = blk
                comp = zlib.compress(_shuffle(chunk.tobytes(), 8))
                keys.append((len(comp), (i0, j0), img.alloc(comp)))
        node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(keys), h5lite.UNDEF, h5lite.UNDEF)
        for size, (i0, j0), addr in keys:
            node += struct.pack("<IIQQQ", size, 0, i0, j0, 0) + struct.pack("<Q", addr)
        node += struct.pack("<IIQQQ", 0, 0, 8, 12, 0)
        bt = img.alloc(node)
        space = struct.pack("<BBB5x", 1, 2, 0) + struct.pack("<QQ", 7, 10)
        layout = struct.pack("<BBBQIII", 3, 2, 3, bt, 4, 4, 8)
        pipeline = struct.pack("<BB6x", 1, 2) + struct.pack("<HHHH", 2, 0, 0, 1) + struct.pack("<II", 8, 0) + \
            struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<II", 6, 0)
        return h5lite._object_header([h5lite._message(1, space), h5lite._message(3, h5lite._dtype_message(np.float64), 1),
                                      h5lite._message(0x0B, pipeline), h5lite._message(8, layout)])
    assert np.array_equal(h5lite.File(_classic_file_with(build))["d"], a)


def test_variable_length_string_in_the_global_heap():
    text = "X/2 corpse µs".encode("utf-8")

    def build(img):
        gcol = b"GCOL" + struct.pack("<B3xQ", 1, 4096) + struct.pack("<HH4xQ", 1, 1, len(text)) + text + b"\x00" * (-len(text) % 8)
        gcol += struct.pack("<HH4xQ", 0, 0, 4096 - len(gcol) - 16)
        addr = img.alloc(gcol + b"\x00" * (4096 - len(gcol)))
        raw = img.alloc(struct.pack("<IQI", len(text), addr, 1))
        space = struct.pack("<BBB5x", 1, 0, 0)
        dtype = struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0, 16) + struct.pack("<BBBBI", 0x13, 0x10, 0, 0, 1)
        layout = struct.pack("<BBQQ", 3, 1, raw, 16)
        return h5lite._object_header([h5lite._message(1, space), h5lite._message(3, dtype, 1), h5lite._message(8, layout)])
    assert h5lite.File(_classic_file_with(build))["d"] == text.decode("utf-8")


def test_version2_superblock_object_headers_and_link_messages():
    a = np.arange(5, dtype=np.int64)
    img = h5lite._Image()
    img.b += b"\x00" * 48

    def ohdr(messages):                                        # version 2, chunk-0 size in one byte, no times, checksum left zero
        body = b"".join(struct.pack("<BHB", t, len(m), 0) + m for t, m in messages)
        return b"OHDR" + struct.pack("<BBB", 2, 0, len(body)) + body + b"\x00" * 4
    raw = img.alloc(a.tobytes())
    compact = np.float64([2.5, -1.0]).tobytes()
    d1 = img.alloc(ohdr([(1, struct.pack("<BBBB", 2, 1, 0, 1) + struct.pack("<Q", 5)), (3, h5lite._dtype_message(np.int64)),
                         (8, struct.pack("<BBQQ", 3, 1, raw, 40))]))
    d2 = img.alloc(ohdr([(1, struct.pack("<BBBB", 2, 1, 0, 1) + struct.pack("<Q", 2)), (3, h5lite._dtype_message(np.float64)),
                         (8, struct.pack("<BBH", 3, 0, len(compact)) + compact)]))
    link = lambda name, addr: (6, struct.pack("<BBB", 1, 0, len(name)) + name + struct.pack("<Q", addr))
    cont = img.alloc(b"OCHK" + struct.pack("<BHB", 6, 3 + 7 + 8, 0) + link(b"compact", d2)[1] + b"\x00" * 4)
    root = img.alloc(ohdr([(2, struct.pack("<BBQQ", 0, 0, h5lite.UNDEF, h5lite.UNDEF)), link(b"ints", d1),
                           (0x10, struct.pack("<QQ", cont, 4 + 4 + 18 + 4))]))
    sb = h5lite.SIGNATURE + struct.pack("<BBBB", 2, 8, 8, 0) + struct.pack("<QQQQ", 0, h5lite.UNDEF, len(img.b), root) + b"\x00" * 4
    img.b[:48] = sb
    f = h5lite.File(bytes(img.b))
    assert f.keys() == ["compact", "ints"]
    assert np.array_equal(f["ints"], a) and np.array_equal(f["compact"], [2.5, -1.0])
