"""A plain-C client process (no Python, no torch in it) calls librbq.so the way a Julia `ccall` would; its printed results
must match the oracle.  This is the drop-in boundary exercised end to end on the GPU box."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_client_through_the_abi(tmp_path, oracle):
    exe = tmp_path / "abi_client"
    libdir = os.path.join(ROOT, "rbqoc_b200")
    subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-Werror", os.path.join(ROOT, "tests", "abi_client.c"), "-I",
                    os.path.join(ROOT, "include"), "-L", libdir, "-lrbq", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = {}
    for line in r.stdout.strip().splitlines():
        k, *v = line.split()
        out[k] = np.array([float(t) for t in v])
    assert out["version"][0] == 100 and list(out["dims"]) == [43, 1]

    p = oracle.make_params(model_id=12, fq_samples=(1.4e-2 * 1.01, 1.4e-2 * 0.99))
    x = np.array([0.1 * ((i * 7) % 11) - 0.5 for i in range(43)])
    x[9] = 0.37
    u = np.array([0.05])
    ref = oracle.step(p, x, u, 0.1)
    assert np.max(np.abs(out["step"] - ref)) / np.max(np.abs(ref)) < 1e-12
    J = oracle.jacobian(p, x, u, 0.1)                       # (n, n+m); the C side prints column-major
    got = out["jac"].reshape(44, 43).T
    assert np.max(np.abs(got - J)) / np.max(np.abs(J)) < 1e-10

    controls = np.array([0.3 * (j % 7) / 7.0 - 0.1 for j in range(50)])
    fq = 1.4e-2 * (1.0 + 0.01 * (np.arange(4) - 1.5))
    da = 1e-5 * np.arange(4)
    psi0 = np.array([[0.5, 0.5, 0.5 if s % 2 else -0.5, 0.5] for s in range(4)])
    fid = oracle.mc_static(controls, 0.1, 3, 2, fq, da, psi0)
    assert np.max(np.abs(out["fid"].reshape(4, 3) - fid)) < 1e-12
    assert out["stats"][0] == 4 and out["stats"][1] == 0
    assert abs(out["stats"][2] - np.sum(1 - fid[:, -1])) < 1e-12
    assert out["bad_gate"][0] == -4 and out["null_ptr"][0] == -1      # RBQ_E_GATE, RBQ_E_NULL
    assert out["launches"][0] >= 5
