"""CPU suite, part 2: the C-ABI library loads and exports exactly what include/rbq.h declares.
No compute call is made here (there is no GPU in the build container, and no CPU fallback to call)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import rbqoc_b200 as rbq
from rbqoc_b200 import _lib, spin

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "rbq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rbq_[a-z0-9_]+)\s*\(", src)))


def test_header_compiles_as_plain_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "rbq.h"\nint main(void){ rbq_model_params p; rbq_stats s; (void)p; (void)s; return sizeof(p) == 128 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rbq.h but not exported by librbq.so"
    assert sorted(_lib.SIGNATURES) == names, "ctypes signature table and header disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (rbq_[a-z0-9_]+)", out)))
    assert exported == names


def test_struct_layouts_match_the_header():
    assert C.sizeof(_lib.ModelParams) == 128
    assert C.sizeof(_lib.Stats) == 8 * (6 + 256)
    assert rbq.STATS_INT64_WORDS == 262
    assert _lib.ModelParams.fq.offset == 48 and _lib.ModelParams.S_diag.offset == 96


def test_model_dims_without_a_device():
    assert _lib.load().rbq_version() == 100
    dims = {"spin13": (11, 1), "spin12": (43, 1), "spin18": (43, 1), "spin30": (56, 1)}
    for name, (n, m) in dims.items():
        mdl = getattr(spin, "Spin" + name[4:] + "Model")()
        assert (mdl.n, mdl.m) == (n, m)
    assert (spin.Spin15Model(True, True).n, spin.Spin15Model(True, True).m) == (12, 2)
    assert spin.Spin11Model(2).n == 19 and spin.Spin17Model(1).n == 15
    assert spin.Spin30Model(sample_state_count=4, nominal_idxs=[1, 5, 9, 1]).n == 15 + 41 * 4
    bad = _lib.ModelParams()
    bad.model_id = 99
    assert _lib.load().rbq_state_dim(C.byref(bad)) == _lib.E_MODEL
    assert _lib.load().rbq_state_dim(None) == _lib.E_NULL


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_no_cpu_fallback():
    """Without a device the engine refuses to exist; it never computes on the host."""
    with pytest.raises(rbq.RbqError) as e:
        rbq.Engine(0)
    assert e.value.code == _lib.E_NODEVICE
    h = C.c_void_p()
    assert _lib.load().rbq_create(0, C.byref(h)) == _lib.E_NODEVICE and not h.value
    # null handles are argument errors, not crashes
    assert _lib.load().rbq_synchronize(None) == _lib.E_NULL
    assert _lib.load().rbq_launch_count(None) == -1


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_plain_c_client_links_and_fails_loudly_without_a_device(tmp_path):
    """tests/abi_client.c is what a ccall-style consumer looks like; without a GPU it must stop at rbq_create (-6)."""
    exe = tmp_path / "abi_client"
    libdir = os.path.join(ROOT, "rbqoc_b200")
    subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-Werror", os.path.join(ROOT, "tests", "abi_client.c"), "-I",
                    os.path.join(ROOT, "include"), "-L", libdir, "-lrbq", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 2 and "rbq_create(0, &h) -> -6" in r.stderr and "version 100" in r.stdout


def test_product_never_imports_the_oracle():
    """Only tests/, smoke() and bench.py's cpu legs may touch oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rbqoc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(from|import)\s+oracle|liboracle|\borc_|oracle\.(py|cpp)\b.*#include", txt), f"{f} uses the oracle"


def test_stats_merge_is_a_host_function():
    a, b = _lib.Stats(), _lib.Stats()
    a.count, a.sum, a.sumsq, a.min, a.max = 2, 3.0, 5.0, 1.0, 2.0
    b.count, b.sum, b.sumsq, b.min, b.max, b.nonfinite = 1, 0.5, 0.25, 0.5, 0.5, 4
    a.hist[3] = 2
    b.hist[3] = 1
    b.hist[200] = 7
    m = rbq.stats_merge(a, b)
    assert (m.count, m.nonfinite, m.sum, m.sumsq, m.min, m.max) == (3, 4, 3.5, 5.25, 0.5, 2.0)
    assert m.hist[3] == 3 and m.hist[200] == 7
    assert m.mean == pytest.approx(3.5 / 3)
