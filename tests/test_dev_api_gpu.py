"""The device-pointer (`_dev`) entry points: same results as the host-pointer forms, statistics ACCUMULATE until reset,
work is enqueued on the stream given to rbq_set_stream (torch provides device memory and streams, nothing else)."""
import numpy as np
import pytest

import rbqoc_b200 as rbq
from rbqoc_b200 import spin

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _dev(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_dev_entry_points_match_host_entry_points(engine):
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    eng = rbq.Engine(0)
    eng.set_stream(stream.cuda_stream)
    controls = spin.pulse_sin(301)[:-1]
    nk, dt, gate, G, S = controls.shape[0], 0.1, rbq.GATE_XPIBY2, 3, 3000
    fq, da, psi0 = engine.philox_samples(7, S, 4, spin.FQ, 0.01, 0.03, 2.574e-5)
    fid_h, st_h = engine.mc_static(controls, dt, gate, G, fq, da, psi0)

    with torch.cuda.stream(stream):
        dc, dfq, dda, dpsi = (_dev(a, dev) for a in (controls, fq, da, psi0))
        dfid = torch.empty(S, G + 1, dtype=torch.float64, device=dev)
        dst = torch.zeros(rbq.STATS_INT64_WORDS, dtype=torch.int64, device=dev)
    eng.stats_reset_dev(dst)
    eng.mc_static_dev(dc, nk, dt, gate, G, S, dfq, dda, dpsi, rbq.ALGO_SU2, dfid, dst)
    eng.synchronize()
    st = rbq.stats_from_words(dst.cpu().numpy())
    assert np.array_equal(dfid.cpu().numpy(), fid_h)
    assert st.count == S and st.sum == st_h.sum and st.hist[:] == st_h.hist[:]
    # statistics accumulate across calls until reset
    eng.mc_static_dev(dc, nk, dt, gate, G, S, dfq, dda, dpsi, rbq.ALGO_SU2, None, dst)
    eng.synchronize()
    st2 = rbq.stats_from_words(dst.cpu().numpy())
    assert st2.count == 2 * S and abs(st2.sum - 2 * st.sum) < 1e-12 * st.sum and st2.min == st.min
    # Philox variant draws the same samples on the device
    eng.stats_reset_dev(dst)
    eng.mc_static_philox_dev(dc, nk, dt, gate, G, 7, S, 4, spin.FQ, 0.01, 0.03, 2.574e-5, rbq.ALGO_SU2, dfid, dst)
    eng.synchronize()
    assert np.array_equal(dfid.cpu().numpy(), fid_h)

    # time-dependent noise
    ndi = 10.0
    nlen = int(np.ceil(nk * dt * G * ndi)) + 1
    noise = engine.pink_noise(7, 500, 4, spin.NAMP_PREFACTOR * 1e3, ndi, nlen)
    fidn_h, stn_h = engine.mc_noise(controls, dt, gate, G, spin.FQ, noise, ndi, psi0[:500])
    with torch.cuda.stream(stream):
        dn = _dev(noise, dev)
        dfn = torch.empty(500, G + 1, dtype=torch.float64, device=dev)
    eng.stats_reset_dev(dst)
    eng.mc_noise_dev(dc, nk, dt, gate, G, 500, spin.FQ, dn, nlen, ndi, dpsi, rbq.ALGO_SU2, dfn, dst)
    eng.synchronize()
    assert np.array_equal(dfn.cpu().numpy(), fidn_h)
    assert rbq.stats_from_words(dst.cpu().numpy()).count == 500

    # Lindblad
    rng = np.random.default_rng(0)
    psi = rng.random((400, 2)) + 1j * rng.random((400, 2))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    rho0 = 0.9 * np.einsum("si,sj->sij", psi, psi.conj()) + 0.05 * np.eye(2)
    fl_h, rho_h, _ = engine.lindblad(controls, dt, gate, G, spin.FQ, da[:400], rho0)
    vec = np.ascontiguousarray(rho0.transpose(0, 2, 1).reshape(400, 4)).view(np.float64)
    with torch.cuda.stream(stream):
        drho = _dev(vec, dev)
        dfl = torch.empty(400, G + 1, dtype=torch.float64, device=dev)
        dro = torch.empty(400, 8, dtype=torch.float64, device=dev)
    eng.stats_reset_dev(dst)
    eng.lindblad_dev(dc, nk, dt, gate, G, 400, spin.FQ, dda, drho, rbq.LINDBLAD_T1, 0.0, 0.0, dfl, dro, dst)
    eng.synchronize()
    assert np.array_equal(dfl.cpu().numpy(), fl_h)
    out = dro.cpu().numpy().view(np.complex128).reshape(400, 2, 2).transpose(0, 2, 1)
    assert np.array_equal(out, rho_h)

    # trajectories
    model = spin.Spin12Model()
    N, B = 101, 5
    U = 1e-3 * rng.standard_normal((B, N - 1, 1))
    dts = np.full(N - 1, 0.1)
    x0 = np.tile(spin.spin12_x0(), (B, 1))
    X_h = engine.rollout(model.params, x0, U, dts)
    with torch.cuda.stream(stream):
        dX = torch.empty(B, N, 43, dtype=torch.float64, device=dev)
        dx0, dU, ddt = _dev(x0, dev), _dev(U, dev), _dev(dts, dev)
    eng.rollout_dev(model.params, B, N, dx0, dU, ddt, dX)
    K = N - 1
    with torch.cuda.stream(stream):
        dAB = torch.empty(K, 44, 43, dtype=torch.float64, device=dev)
    eng.jacobians_dev(model.params, K, dX[0, :-1].contiguous(), dU[0].contiguous(), ddt, dAB)
    eng.synchronize()
    assert np.array_equal(dX.cpu().numpy(), X_h)
    AB_h = engine.jacobians(model.params, X_h[0, :-1], U[0], dts)
    assert np.array_equal(dAB.cpu().numpy().transpose(0, 2, 1), AB_h)
    assert eng.launch_count > 10
    eng.close()
