"""GPU parity: Monte-Carlo robustness sweeps, time-dependent noise, Lindblad rollout, statistics.

The oracle runs the reference's algorithm (per-knot generic Pade exp, 4x4 propagator products,
extended-precision states); the engine runs the closed-form SU(2)/Bloch kernels.  Tolerance: relative
1e-12 on fidelities/states (north star), looser only where stated and why.
"""
import numpy as np
import pytest

import rbqoc_b200 as rbq
from _cases import GOLDEN
from rbqoc_b200 import spin

pytestmark = pytest.mark.gpu

FID_RTOL = 1e-12


def _samples(S, seed=0, fq_sigma=0.01, da_sigma=2.574e-5):
    rng = np.random.default_rng(seed)
    fq = spin.FQ * (1 + np.clip(fq_sigma * rng.standard_normal(S), -0.03, 0.03))
    da = da_sigma * rng.standard_normal(S)
    psi = rng.random((S, 2)) + 1j * rng.random((S, 2))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    psi0 = np.concatenate([psi.real, psi.imag], axis=1)
    return fq, da, psi0


PULSES = {
    "sin301_dt0.1": (spin.pulse_sin(301)[:-1], 0.1, rbq.GATE_XPIBY2),
    "sin1001_dt0.1": (spin.pulse_sin(1001)[:-1], 0.1, rbq.GATE_ZPIBY2),
    "y2_dt0.01": (spin.controls_ypiby2(100.0), 0.01, rbq.GATE_YPIBY2),
    "x2_dt0.01": (spin.controls_xpiby2(100.0), 0.01, rbq.GATE_XPIBY2),
    "z2_dt0.01": (spin.controls_zpiby2(100.0), 0.01, rbq.GATE_ZPIBY2),
    "ragged_2049": (0.4 * np.cos(np.arange(2049) * 0.01), 0.05, rbq.GATE_XPI),   # crosses a tile boundary, odd length
    "single_knot": (np.array([0.25]), 0.1, rbq.GATE_ZPIBY2),
}


@pytest.mark.parametrize("name", list(PULSES))
@pytest.mark.parametrize("algo", [rbq.ALGO_SU2, rbq.ALGO_PADE], ids=["su2", "pade"])
def test_mc_static_matches_oracle(engine, oracle, name, algo):
    controls, dt, gate = PULSES[name]
    S = 1000 if controls.shape[0] > 3000 else 2500
    fq, da, psi0 = _samples(S, seed=len(name))
    gate_count = 5
    fid, st = engine.mc_static(controls, dt, gate, gate_count, fq, da, psi0, algo=algo)
    ref = oracle.mc_static(controls, dt, gate, gate_count, fq, da, psi0)
    assert fid.shape == (S, gate_count + 1)
    # one gate: the north-star tolerance; the reference algorithm's own rounding grows coherently with the knot
    # count on piecewise-constant pulses (~2e-16 per knot), which only the 5680-knot X/2 feels
    nk = controls.shape[0]
    assert np.max(np.abs(fid[:, :2] - ref[:, :2])) < max(FID_RTOL, 5e-16 * nk), name
    # the oracle (= the reference algorithm) itself sits up to ~4e-13 from the exact answer after 5 gates of a
    # ~2000-knot pulse (50-digit referee, tests/test_oracle_cpu.py), so this is at the reference's own noise floor
    assert np.max(np.abs(fid - ref)) < max(5 * FID_RTOL, 2.5e-15 * nk), name
    # statistics of the last-gate infidelity are exactly those of the returned fidelities
    e = 1.0 - fid[:, -1]
    assert st.count == S and st.nonfinite == 0
    assert abs(st.sum - e.sum()) <= 1e-12 * abs(e.sum()) + 1e-18
    assert st.min == e.min() and st.max == e.max()
    ost = oracle.stats(e)
    assert np.sum(np.abs(np.array(st.hist[:]) - np.array(ost.hist[:]))) <= 2   # a bin-edge ulp at most
    assert sum(st.hist[:]) == S


@pytest.mark.parametrize("dt,amp", [(0.5, 0.3), (1.0, 0.4), (2.5, 0.5), (0.2, 0.5)], ids=["class2", "class3", "class3_big", "class2_edge"])
@pytest.mark.parametrize("algo", [rbq.ALGO_SU2, rbq.ALGO_PADE], ids=["su2", "pade"])
def test_mc_sweeps_with_large_rotation_angles(engine, oracle, dt, amp, algo):
    """Coarse grids push x = theta^2 past the tan-form polynomials: the normalised K = 8 series and the sincos class of the
    sweep kernels, and Pade-7/9/13 (+ squaring) of the reference-algorithm kernels, for the static, noisy and Lindblad sweeps."""
    nk = 120
    controls = amp * np.sin(np.arange(nk) * 0.37)
    S = 300
    fq, da, psi0 = _samples(S, seed=int(dt * 10))
    fid, st = engine.mc_static(controls, dt, rbq.GATE_YPIBY2, 3, fq, da, psi0, algo=algo)
    ref = oracle.mc_static(controls, dt, rbq.GATE_YPIBY2, 3, fq, da, psi0)
    assert np.max(np.abs(fid - ref)) < FID_RTOL
    ndi = 1.0 / dt
    nlen = int(np.ceil(nk * dt * 3 * ndi)) + 1
    noise = 0.02 * np.random.default_rng(1).standard_normal((S, nlen))
    fidn, _ = engine.mc_noise(controls, dt, rbq.GATE_YPIBY2, 3, spin.FQ, noise, ndi, psi0, algo=algo)
    refn = oracle.mc_noise(controls, dt, rbq.GATE_YPIBY2, 3, spin.FQ, noise, ndi, psi0)
    assert np.max(np.abs(fidn - refn)) < FID_RTOL
    if algo == rbq.ALGO_SU2:
        rho0 = _rand_rho(64, 3, mixed=True)
        fl, rho, _ = engine.lindblad(controls, dt, rbq.GATE_YPIBY2, 5, spin.FQ, da[:64], rho0)
        rfid, rfc, rrho = oracle.lindblad_t1(controls, dt, rbq.GATE_YPIBY2, 5, spin.FQ, da[:64], rho0)
        assert np.max(np.abs(rho - rrho)) < 1e-12 and np.max(np.abs(fl - rfc)) < 1e-12


def test_mc_noise_sample_leaving_the_polynomial_range(engine, oracle):
    """A noise burst pushes single knots outside the tan-form polynomial's x range: the per-knot guard must take the
    general step for them (and only them)."""
    controls, dt, gate = PULSES["sin301_dt0.1"]
    S, gc, ndi = 96, 2, 10.0
    _, _, psi0 = _samples(S, seed=13)
    nlen = int(np.ceil(controls.shape[0] * dt * gc * ndi)) + 1
    noise = np.zeros((S, nlen))
    noise[::3, 40:55] = 3.0            # |a + delta a| ~ 3 GHz for a few knots of every third sample
    for mode in ("0", "1"):
        import os
        engine.set_option("RBQ_NOISE_GATE_PARALLEL", mode)
        try:
            fid, _ = engine.mc_noise(controls, dt, gate, gc, spin.FQ, noise, ndi, psi0)
        finally:
            engine.set_option("RBQ_NOISE_GATE_PARALLEL", None)
        ref = oracle.mc_noise(controls, dt, gate, gc, spin.FQ, noise, ndi, psi0)
        assert np.max(np.abs(fid - ref)) < FID_RTOL, mode


def test_mc_static_analytic_gates_are_exact(engine):
    """Known answers embedded in the reference (spin.jl:557, 608-612): the idle Z/2 and the
    three-segment Y/2 reproduce their gates for every initial state at zero detuning."""
    S = 512
    fq, _, psi0 = _samples(S, seed=3, fq_sigma=0.0)
    n = 1786
    fid, _ = engine.mc_static(np.zeros(n), spin.TTOT_ZPIBY2 / n, rbq.GATE_ZPIBY2, 8, fq, None, psi0)
    assert np.max(np.abs(fid - 1.0)) < 1e-13
    # exact-duration Y/2: segments as one knot each via three calls is not expressible; use fine dt instead
    c = spin.controls_ypiby2(100.0)
    fid, _ = engine.mc_static(c, 0.01, rbq.GATE_YPIBY2, 1, fq, None, psi0)
    assert np.max(1.0 - fid[:, 1]) < 2e-4   # the two zeroed end samples and the floor() truncation of spin14.py:116


def test_mc_static_detuning_gate_error_pin(engine):
    """paper main.tex:857-859 / trials.xlsx: analytic Z/2 at +-1% detuning averages ~4.7e-5."""
    S = 2000
    _, _, psi0 = _samples(S, seed=9)
    fq = spin.FQ * np.where(np.arange(S) % 2 == 0, 1.01, 0.99)
    n = 1786
    fid, st = engine.mc_static(np.zeros(n), spin.TTOT_ZPIBY2 / n, rbq.GATE_ZPIBY2, 1, fq, None, psi0)
    assert 4.3e-5 < st.mean < 5.1e-5


def test_mc_static_many_gates_and_empty(engine, oracle):
    controls, dt, gate = PULSES["sin301_dt0.1"]
    fq, da, psi0 = _samples(64, seed=4)
    fid, st = engine.mc_static(controls, dt, gate, 200, fq, da, psi0)
    ref = oracle.mc_static(controls, dt, gate, 200, fq, da, psi0)
    assert np.max(np.abs(fid - ref)) < 2e-11        # 200 applications amplify the one-gate rounding ~200x
    fid0, st0 = engine.mc_static(controls, dt, gate, 3, np.empty(0), np.empty(0), np.empty((0, 4)))
    assert fid0.shape == (0, 4) and st0.count == 0 and st0.min == np.inf
    with pytest.raises(rbq.RbqError):
        engine.mc_static(controls, dt, 99, 1, fq, da, psi0)
    with pytest.raises(rbq.RbqError):
        engine.mc_static(controls, dt, gate, 1, fq, da, psi0, algo=7)


def test_long_pulse_many_tiles(engine, oracle):
    """20 001 knots = 10 shared-memory tiles of the static pipeline, 20 of the noise pipeline (odd length, ragged last tile)."""
    nk = 20001
    controls = 0.3 * np.sin(np.arange(nk) * 1e-3) + 0.05 * np.cos(np.arange(nk) * 7e-3)
    S = 48
    fq, da, psi0 = _samples(S, seed=21)
    fid, _ = engine.mc_static(controls, 0.01, rbq.GATE_XPI, 2, fq, da, psi0)
    ref = oracle.mc_static(controls, 0.01, rbq.GATE_XPI, 2, fq, da, psi0)
    assert np.max(np.abs(fid - ref)) < max(5e-12, 2.5e-15 * nk)
    ndi = 10.0
    nlen = int(np.ceil(nk * 0.01 * 2 * ndi)) + 1
    noise = oracle.pink_noise(0, S, 3, spin.NAMP_PREFACTOR * 1e3, ndi, nlen)
    fidn, _ = engine.mc_noise(controls, 0.01, rbq.GATE_XPI, 2, spin.FQ, noise, ndi, psi0)
    refn = oracle.mc_noise(controls, 0.01, rbq.GATE_XPI, 2, spin.FQ, noise, ndi, psi0)
    assert np.max(np.abs(fidn - refn)) < max(5e-12, 2.5e-15 * nk * 2)


def test_zero_gates(engine):
    """gate_count = 0: only the initial fidelity F(psi0, psi0) = 1 is reported; the statistics are those of 1 - 1."""
    controls, dt, gate = PULSES["sin301_dt0.1"]
    fq, da, psi0 = _samples(100, seed=2)
    fid, st = engine.mc_static(controls, dt, gate, 0, fq, da, psi0)
    assert fid.shape == (100, 1) and np.max(np.abs(fid - 1)) < 1e-15 and st.count == 100 and st.max < 1e-15
    noise = np.zeros((100, 1))
    fidn, stn = engine.mc_noise(controls, dt, gate, 0, spin.FQ, noise, 10.0, psi0)
    assert fidn.shape == (100, 1) and np.max(np.abs(fidn - 1)) < 1e-15 and stn.count == 100
    rho0 = _rand_rho(100, 1, mixed=True)
    fidl, rho, stl = engine.lindblad(controls, dt, gate, 0, spin.FQ, None, rho0)
    assert fidl.shape == (100, 1) and np.max(np.abs(rho - rho0)) < 1e-15
    assert np.max(np.abs(fidl[:, 0] - 1.0)) < 1e-12            # sqrt-fidelity of a state with itself


def test_mc_static_nonfinite_samples_are_counted(engine):
    controls, dt, gate = PULSES["sin301_dt0.1"]
    fq, da, psi0 = _samples(300, seed=5)
    fq[7] = np.nan
    da[100] = np.inf
    fid, st = engine.mc_static(controls, dt, gate, 1, fq, da, psi0)
    assert st.nonfinite == 2 and st.count == 298
    assert np.isnan(fid[7, 1]) and np.isnan(fid[100, 1])


def test_philox_samples_match_oracle(engine, oracle):
    args = (12345, 4096, 77, spin.FQ, 0.01, 0.03, 2.574e-5)
    fq, da, psi0 = engine.philox_samples(*args)
    rfq, rda, rpsi0 = oracle.philox_samples(*args)
    assert np.max(np.abs(fq - rfq)) < 1e-15 * spin.FQ * 10
    assert np.max(np.abs(da - rda)) < 1e-15 * 2.574e-5 * 100
    assert np.max(np.abs(psi0 - rpsi0)) < 4e-16
    assert np.max(np.abs(fq / spin.FQ - 1)) <= 0.03 + 1e-15
    # counter-based: a shard starting elsewhere draws the same samples
    fq2, da2, psi2 = engine.philox_samples(12345 + 1000, 100, 77, spin.FQ, 0.01, 0.03, 2.574e-5)
    assert np.array_equal(fq2, fq[1000:1100]) and np.array_equal(psi2, psi0[1000:1100])


def test_mc_static_philox_equals_explicit_samples(engine):
    controls, dt, gate = PULSES["sin1001_dt0.1"]
    S, first, seed = 5000, 999, 3
    fq, da, psi0 = engine.philox_samples(first, S, seed, spin.FQ, 0.01, 0.03, 2.574e-5)
    fid_e, st_e = engine.mc_static(controls, dt, gate, 2, fq, da, psi0)
    fid_p, st_p = engine.mc_static_philox(controls, dt, gate, 2, first, S, seed, spin.FQ, 0.01, 0.03, 2.574e-5, want_fid=True)
    assert np.array_equal(fid_e, fid_p)
    assert st_e.count == st_p.count == S and st_e.sum == st_p.sum and st_e.hist[:] == st_p.hist[:]
    # sharding invariance: two half sweeps merge to the whole
    _, a = engine.mc_static_philox(controls, dt, gate, 2, first, 2000, seed, spin.FQ, 0.01, 0.03, 2.574e-5)
    _, b = engine.mc_static_philox(controls, dt, gate, 2, first + 2000, 3000, seed, spin.FQ, 0.01, 0.03, 2.574e-5)
    m = rbq.stats_merge(a, b)
    assert m.count == S and m.hist[:] == st_p.hist[:] and m.min == st_p.min and m.max == st_p.max
    assert abs(m.sum - st_p.sum) < 1e-12 * st_p.sum


@pytest.mark.parametrize("gate_parallel", ["0", "1"], ids=["sequential", "gate_parallel"])
@pytest.mark.parametrize("algo", [rbq.ALGO_SU2, rbq.ALGO_PADE], ids=["su2", "pade"])
def test_mc_noise_matches_oracle(engine, oracle, algo, gate_parallel, monkeypatch):
    """Both forms of the sweep: one chain per sample, and (few samples) one gate propagator per (sample, gate)."""
    engine.set_option("RBQ_NOISE_GATE_PARALLEL", gate_parallel)
    controls = spin.controls_ypiby2(100.0)
    dt, gate, gate_count, ndi = 0.01, rbq.GATE_YPIBY2, 3, 10.0
    S = 200
    _, _, psi0 = _samples(S, seed=6)
    nlen = int(np.ceil(controls.shape[0] * dt * gate_count * ndi)) + 1
    noise = oracle.pink_noise(0, S, 11, spin.NAMP_PREFACTOR * 1e3, ndi, nlen)   # exaggerated so it matters
    fid, st = engine.mc_noise(controls, dt, gate, gate_count, spin.FQ, noise, ndi, psi0, algo=algo)
    ref = oracle.mc_noise(controls, dt, gate, gate_count, spin.FQ, noise, ndi, psi0)
    assert np.max(np.abs(fid - ref)) < 5e-12      # ~5800 chained steps
    assert st.count == S
    assert np.std(fid[:, -1]) > 0                  # the noise is actually felt


@pytest.mark.parametrize("gate_parallel", ["0", "1"], ids=["sequential", "gate_parallel"])
def test_pink_noise_matches_oracle_and_drives_sweep(engine, oracle, gate_parallel, monkeypatch):
    engine.set_option("RBQ_NOISE_GATE_PARALLEL", gate_parallel)
    nlen = 500
    got = engine.pink_noise(40, 33, 5, spin.NAMP_PREFACTOR, 10.0, nlen)
    ref = oracle.pink_noise(40, 33, 5, spin.NAMP_PREFACTOR, 10.0, nlen)
    assert np.max(np.abs(got - ref)) < 1e-12 * np.max(np.abs(ref))
    # ragged shapes of the two-pass generator (rows that do not fill a 32-row tile, odd lengths, one sample, one entry)
    for first, S, nl in ((0, 70, 67), (7, 1, 1), (3, 32, 64), (11, 5, 33)):
        got = engine.pink_noise(first, S, 9, spin.NAMP_PREFACTOR, 10.0, nl)
        ref = oracle.pink_noise(first, S, 9, spin.NAMP_PREFACTOR, 10.0, nl)
        assert got.shape == ref.shape and np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))
    # on-the-fly generation inside the sweep == the materialised noise fed through rbq_mc_noise
    controls, dt, gate, gc = spin.pulse_sin(301)[:-1], 0.1, rbq.GATE_XPIBY2, 4
    S, first, seed, namp = 130, 40, 5, spin.NAMP_PREFACTOR * 1e3
    nl = int(np.ceil(controls.shape[0] * dt * gc * 10.0)) + 1
    noise = engine.pink_noise(first, S, seed, namp, 10.0, nl)
    _, _, psi0 = engine.philox_samples(first, S, seed, spin.FQ, 0, 0, 0)
    fid_a, st_a = engine.mc_noise(controls, dt, gate, gc, spin.FQ, noise, 10.0, psi0)
    fid_b, st_b = engine.mc_pink_philox(controls, dt, gate, gc, first, S, seed, spin.FQ, namp, 10.0)
    assert np.max(np.abs(fid_a - fid_b)) < 1e-14
    assert st_a.count == st_b.count == S and abs(st_a.sum - st_b.sum) < 1e-12
    # the two forms of the sweep agree with each other to rounding
    engine.set_option("RBQ_NOISE_GATE_PARALLEL", "1" if gate_parallel == "0" else "0")
    fid_c, _ = engine.mc_noise(controls, dt, gate, gc, spin.FQ, noise, 10.0, psi0)
    assert np.max(np.abs(fid_a - fid_c)) < 1e-13


def _rand_rho(S, seed, mixed=False):
    rng = np.random.default_rng(seed)
    psi = rng.random((S, 2)) + 1j * rng.random((S, 2))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    rho = np.einsum("si,sj->sij", psi, psi.conj())
    if mixed:
        lam = rng.uniform(0.6, 1.0, S)[:, None, None]
        rho = lam * rho + (1 - lam) * np.eye(2) / 2
    return rho


@pytest.mark.parametrize("name", ["sin301_dt0.1", "y2_dt0.01", "z2_dt0.01"])
@pytest.mark.parametrize("mixed", [False, True], ids=["pure", "mixed"])
def test_lindblad_t1_matches_oracle(engine, oracle, name, mixed):
    controls, dt, gate = PULSES[name]
    S, gate_count = 300, 6
    rho0 = _rand_rho(S, 8, mixed)
    da = 1e-3 * np.random.default_rng(2).standard_normal(S)
    fid, rho, st = engine.lindblad(controls, dt, gate, gate_count, spin.FQ, da, rho0)
    rfid, rfid_closed, rrho = oracle.lindblad_t1(controls, dt, gate, gate_count, spin.FQ, da, rho0)
    # the oracle's per-knot complex Pade exponentials put it 4e-13..8e-13 from the 50-digit answer after 6 gates of a
    # ~2000-knot pulse (test_lindblad_matches_50_digit_golden below pins the engine at 1e-13), hence 5e-12 here
    assert np.max(np.abs(rho - rrho)) < 5e-12
    # sqrt-fidelity against a PURE target is ill conditioned in any FP64 evaluation: det(target) = 0 is only known to
    # ~1e-16, and F^2 contains 2 sqrt(det rho det target) ~ 2 sqrt(2.5e-4 1e-16) = 3e-10 (conditioning sqrt(eps))
    assert np.max(np.abs(fid - rfid_closed)) < (5e-12 if mixed else 1e-8)
    # the reference's fidelity_mat goes through two matrix square roots (spin.jl:1101-1104); for (near-)pure
    # targets sqrt() turns its 1e-16 rounding into ~1e-8, so the eigen route only agrees to that level
    assert np.max(np.abs(fid - rfid)) < (5e-7 if not mixed else 1e-9)
    assert st.count == S
    assert np.allclose(np.trace(rho, axis1=1, axis2=2), 1.0, atol=1e-14)


def test_mc_static_matches_50_digit_golden(engine):
    from _cases import GOLDEN

    z = np.load(GOLDEN + "/mc_static.npz")
    for algo, tol in ((rbq.ALGO_SU2, 1e-13), (rbq.ALGO_PADE, 2e-12)):
        fid, _ = engine.mc_static(z["controls"], float(z["dt"]), int(z["gate"]), int(z["gate_count"]), z["fq"], z["da"], z["psi0"],
                                  algo=algo)
        assert np.max(np.abs(fid - z["fid"])) < tol, algo


def test_lindblad_matches_50_digit_golden(engine):
    from _cases import GOLDEN

    z = np.load(GOLDEN + "/lindblad.npz")
    G = int(z["gate_count"])
    fid, rho, _ = engine.lindblad(z["controls"], float(z["dt"]), int(z["gate"]), G, float(z["fq"]), z["da"], z["rho0"])
    assert np.max(np.abs(rho - z["rho"][:, G])) < 1e-13


@pytest.mark.parametrize("name", ["sin301_dt0.1", "x2_dt0.01"])
def test_lindblad_shared_map_path(engine, oracle, name, monkeypatch):
    """No per-realisation offsets: the one-gate map is composed once (tree product over knots) and shared by all samples."""
    controls, dt, gate = PULSES[name]
    S, gate_count = 200, 7
    rho0 = _rand_rho(S, 11, mixed=True)
    fid, rho, st = engine.lindblad(controls, dt, gate, gate_count, spin.FQ, None, rho0)
    engine.set_option("RBQ_LINDBLAD_NO_SHARED_MAP", "1")
    fid2, rho2, _ = engine.lindblad(controls, dt, gate, gate_count, spin.FQ, None, rho0)
    # sequential vs tree-ordered products of the same per-knot maps differ by rounding only (~1e-16 per knot and gate)
    nk = controls.shape[0]
    assert np.max(np.abs(rho - rho2)) < max(1e-13, 2e-16 * nk * gate_count) and np.max(np.abs(fid - fid2)) < max(1e-12, 2e-16 * nk * gate_count)
    rfid, rfid_closed, rrho = oracle.lindblad_t1(controls, dt, gate, gate_count, spin.FQ, None, rho0[:40])
    assert np.max(np.abs(rho[:40] - rrho)) < max(5e-12, 2.5e-15 * controls.shape[0])
    assert np.max(np.abs(fid[:40] - rfid_closed)) < max(5e-12, 2.5e-15 * controls.shape[0])
    assert st.count == S


def test_lindblad_d1_pin(engine):
    """D1(Z/2) = T/T1(0) = 5.7441e-5 (paper main.tex:1077, trials.xlsx spin14): the population of a
    z-polarised state relaxes by exactly that much over one idle gate."""
    n = 1786
    rho0 = np.array([[[1.0, 0], [0, 0]]], dtype=np.complex128)
    fid, rho, _ = engine.lindblad(np.zeros(n), spin.TTOT_ZPIBY2 / n, rbq.GATE_ZPIBY2, 1, spin.FQ, None, rho0)
    d1 = spin.TTOT_ZPIBY2 / 310.88e3
    assert abs((1 - rho[0, 0, 0].real) - 0.5 * (1 - np.exp(-2 * d1))) < 1e-13
    assert abs(d1 - 5.7441e-5) < 1e-8


def test_lindblad_t2_dephasing(engine):
    """dynamics_lindbladt2_deqjl (spin.jl:542-554) with the pulse off: coherences decay as
    exp(-(gamma_c t + gamma_f^2 t^2)) while rotating at fq; populations are untouched."""
    n, dt = 2000, 0.05
    gc, gf = spin.GAMMAC, 3e-4
    rho0 = np.array([[[0.5, 0.5], [0.5, 0.5]]], dtype=np.complex128)
    fid, rho, _ = engine.lindblad(np.zeros(n), dt, rbq.GATE_ZPIBY2, 2, spin.FQ, None, rho0, channels=rbq.LINDBLAD_T2,
                                  t2_gamma_c=gc, t2_gamma_f=gf)
    t = 2 * n * dt
    assert abs(abs(rho[0, 0, 1]) - 0.5 * np.exp(-(gc * t + gf * gf * t * t))) < 1e-13
    assert abs(rho[0, 0, 0].real - 0.5) < 1e-15


# ---- analytic gates as segment lists (spin.jl:562-594, 651-670, 765-785, 819-905) ----------------------------------------
@pytest.mark.parametrize("gate", [spin.GateType.zpiby2, spin.GateType.ypiby2, spin.GateType.xpiby2], ids=lambda g: g.name)
def test_analytic_gate_segments_match_oracle_and_golden(engine, oracle, gate):
    z = np.load(GOLDEN + "/segments.npz")
    G = int(z["gate_count"])
    amps, durs = spin.ANALYTIC_SEGMENTS[gate]
    # 50-digit golden: segments of up to 17.9 ns (theta ~ 0.8 rad .. 7 rad with the drive on) -> the sincos class
    fid, st = engine.mc_static_segments(amps, durs, gate, G, z["fq"], z["da"], z["psi0"])
    assert np.max(np.abs(fid - z[f"fid_{gate.name}"])) < 1e-14
    assert np.max(np.abs(fid[0] - 1.0)) < 5e-15                       # nominal sample: the gate is exact
    fidl, rho, stl = engine.lindblad_segments(amps, durs, gate, G, spin.FQ, z["da"], z["rho0"])
    assert np.max(np.abs(rho - z[f"rho_{gate.name}"][:, G])) < 1e-13
    # a sweep against the oracle, ragged sample count, statistics on the last gate
    S = 1237
    fq, da, psi0 = _samples(S, seed=int(gate))
    fid, st = engine.mc_static_segments(amps, durs, gate, 7, fq, da, psi0)
    ref = oracle.mc_static_segments(amps, durs, gate, 7, fq, da, psi0)
    assert np.max(np.abs(fid - ref)) < FID_RTOL
    e = 1.0 - fid[:, -1]
    assert st.count == S and st.min == e.min() and st.max == e.max()
    psi = psi0[:, :2] + 1j * psi0[:, 2:]
    rho0 = np.einsum("si,sj->sij", psi, psi.conj())
    fidl, rho, stl = engine.lindblad_segments(amps, durs, gate, 7, spin.FQ, da, rho0)
    fidc, rref = oracle.lindblad_segments(amps, durs, gate, 7, spin.FQ, da, rho0)
    assert np.max(np.abs(rho - rref)) < FID_RTOL
    assert np.max(np.abs(fidl - fidc)) < 1e-8         # pure targets: sqrt-conditioned in any FP64 evaluation (see test_lindblad_t1_matches_oracle)
    assert stl.count == S


def test_xpiby2da_noise_segments_match_oracle_and_golden(engine, oracle):
    z = np.load(GOLDEN + "/segments.npz")
    amps, durs, nidx = spin.xpiby2da_segments()
    G = int(z["da_gate_count"])
    fid, st = engine.mc_noise_segments(amps, durs, nidx, spin.GateType.xpiby2, G, spin.FQ, z["da_noise"], z["psi0"][:3])
    assert np.max(np.abs(fid - z["da_fid"])) < 1e-13
    S = 333
    _, _, psi0 = _samples(S, seed=5)
    noise = spin.NAMP_PREFACTOR * 30 * np.random.default_rng(2).standard_normal((S, 600))
    fid, st = engine.mc_noise_segments(amps, durs, nidx, spin.GateType.xpiby2, 12, spin.FQ, noise, psi0)
    ref = oracle.mc_noise_segments(amps, durs, nidx, spin.GateType.xpiby2, 12, spin.FQ, noise, psi0)
    assert np.max(np.abs(fid - ref)) < 5 * FID_RTOL     # 12 x 575 chained Pade exponentials on the oracle side
    assert st.count == S
    # bad noise index / short rows are argument errors, not reads past the row
    with pytest.raises(Exception):
        engine.mc_noise_segments(amps, durs, nidx, spin.GateType.xpiby2, 1, spin.FQ, noise[:, :500], psi0)
    # empty sweep
    fid, st = engine.mc_noise_segments(amps, durs, nidx, spin.GateType.xpiby2, 2, spin.FQ, np.zeros((0, 600)), np.zeros((0, 4)))
    assert fid.shape == (0, 3) and st.count == 0


def test_run_sim_prop_analytic_dynamics_types(engine):
    """run_sim_prop(..., dynamics_type=zpiby2nodis / ypiby2t1 / xpiby2da ...) without a pulse file (DT_AN, spin.jl:94-114)."""
    for dyn, gate in ((spin.DynamicsType.zpiby2nodis, spin.GateType.zpiby2), (spin.DynamicsType.ypiby2nodis, spin.GateType.ypiby2),
                      (spin.DynamicsType.xpiby2nodis, spin.GateType.xpiby2)):
        r = spin.run_sim_prop(10, gate, None, 0, seed=np.arange(5), dynamics_type=dyn, engine=engine)
        assert np.max(np.abs(r["fidelities"] - 1)) < 1e-13
    r = spin.run_sim_prop(50, spin.GateType.ypiby2, None, 0, seed=np.arange(4), dynamics_type=spin.DynamicsType.ypiby2t1, engine=engine)
    f = r["fidelities"]
    assert f.shape == (4, 51) and np.all(f[:, -1] < f[:, 0]) and np.all(f[:, -1] > 0.99)
    r = spin.run_sim_prop(20, spin.GateType.xpiby2, None, 0, seed=np.arange(1, 9), dynamics_type=spin.DynamicsType.xpiby2da, engine=engine)
    e = 1 - r["fidelities"][:, -1]
    assert r["fidelities"].shape == (8, 21) and np.all(e >= -1e-13) and np.all(e < 1e-3) and np.any(e > 1e-10)


# ---- the knot-table step (local Taylor polynomial of tan(theta)/theta around the sweep centre) ---------------------------
@pytest.mark.parametrize("name", ["sin1001_dt0.1", "y2_dt0.01", "ragged_2049"])
def test_table_class_equals_polynomial_classes(engine, oracle, name, monkeypatch):
    """Samples near the sweep centre take the table step, the rest the Horner/series step; both carry a < 1e-16 relative
    truncation per knot, so they agree to rounding, and each matches the oracle / the 50-digit golden on its own."""
    controls, dt, gate = PULSES[name]
    S = 4096 + 77
    fq, da, psi0 = _samples(S, seed=21)
    # blocks with mixed warps: every 5th warp is pushed out of the table's range (eps ~ 2 w^2 a da grows with dt^2)
    da[np.arange(S) % 160 < 32] += 2e-3 if dt >= 0.05 else 0.2
    fid_t, st_t = engine.mc_static(controls, dt, gate, 3, fq, da, psi0)
    engine.set_option("RBQ_NO_TABLE_CLASS", "1")
    fid_p, st_p = engine.mc_static(controls, dt, gate, 3, fq, da, psi0)
    engine.set_option("RBQ_NO_TABLE_CLASS", None)
    assert np.max(np.abs(fid_t - fid_p)) < 5e-14
    far = np.arange(S) % 160 < 32
    assert np.array_equal(fid_t[far], fid_p[far])              # out-of-range warps run the same code either way
    assert not np.array_equal(fid_t[~far], fid_p[~far])        # ... and the in-range ones really took the other path
    ref = oracle.mc_static(controls, dt, gate, 3, fq[:600], da[:600], psi0[:600])
    nk = controls.shape[0]
    assert np.max(np.abs(fid_t[:600] - ref)) < max(3 * FID_RTOL, 2.5e-15 * nk)
    assert st_t.count == st_p.count == S


def test_table_class_golden_and_sweep_centre(engine, monkeypatch):
    """50-digit golden with the table step on and off; moving the sweep centre (rbq_set_sweep_centre) only moves rounding."""
    z = np.load(GOLDEN + "/mc_static.npz")
    args = (z["controls"], float(z["dt"]), int(z["gate"]), int(z["gate_count"]), z["fq"], z["da"], z["psi0"])
    fid_t, _ = engine.mc_static(*args)
    engine.set_option("RBQ_NO_TABLE_CLASS", "1")
    fid_p, _ = engine.mc_static(*args)
    engine.set_option("RBQ_NO_TABLE_CLASS", None)
    assert np.max(np.abs(fid_t - z["fid"])) < 1e-13 and np.max(np.abs(fid_p - z["fid"])) < 1e-13
    assert not np.array_equal(fid_t, fid_p)
    # a sweep around another qubit frequency: off-centre it falls back to the polynomial step, re-centred it takes the table
    controls, dt, gate = PULSES["sin301_dt0.1"]
    S = 2000
    fq, da, psi0 = _samples(S, seed=8)
    fq *= 3.0
    da += 1e-3
    fid_off, _ = engine.mc_static(controls, dt, gate, 2, fq, da, psi0)
    try:
        engine.set_sweep_centre(3.0 * spin.FQ, 1e-3)
        fid_on, _ = engine.mc_static(controls, dt, gate, 2, fq, da, psi0)
    finally:
        engine.set_sweep_centre(spin.FQ, 0.0)
    assert np.max(np.abs(fid_on - fid_off)) < 5e-14 and not np.array_equal(fid_on, fid_off)
    with pytest.raises(Exception):
        engine.set_sweep_centre(float("nan"), 0.0)


@pytest.mark.parametrize("gate_count", [1, 3], ids=["pair_store", "row_store"])
def test_mc_static_zero_copy_equals_streamed(engine, monkeypatch, gate_count):
    """Page-locked sample buffers are read (and the fidelities written) by the kernel itself over PCIe; pageable ones are
    streamed in chunks.  Same kernel (RBQ_MC_VARIANT 161: table rows through shared memory in both forms), same samples, same
    bits; the launch count shows which form ran.  With the default variant the streamed form reads the table from constant
    memory (two kernels): the same operations in the same order, so still the same bits."""
    controls, dt, gate = PULSES["sin301_dt0.1"]
    S = 20000 + 13
    fq, da, psi0 = _samples(S, seed=31)
    fid_d, st_d = engine.mc_static(controls, dt, gate, gate_count, fq, da, psi0)           # default variant, streamed
    engine.set_option("RBQ_MC_VARIANT", "161")
    try:
        _zero_copy_equals_streamed(engine, controls, dt, gate, gate_count, S, fq, da, psi0, fid_d, st_d)
    finally:
        engine.set_option("RBQ_MC_VARIANT", "")


def _zero_copy_equals_streamed(engine, controls, dt, gate, gate_count, S, fq, da, psi0, fid_d, st_d):
    fid_s, st_s = engine.mc_static(controls, dt, gate, gate_count, fq, da, psi0)           # pageable -> streamed
    assert np.array_equal(fid_s, fid_d) and st_s.count == st_d.count and st_s.sum == st_d.sum and st_s.hist[:] == st_d.hist[:]
    pfq, pda, ppsi = engine.pinned_empty(S), engine.pinned_empty(S), engine.pinned_empty((S, 4))
    pfid = engine.pinned_empty((S, gate_count + 1))
    pfq[:], pda[:], ppsi[:] = fq, da, psi0
    pfid[:] = -1.0
    before = engine.launch_count
    fid_z, st_z = engine.mc_static(controls, dt, gate, gate_count, pfq, pda, ppsi, fid_out=pfid)
    launches_zero_copy = engine.launch_count - before
    assert fid_z is pfid and np.array_equal(pfid, fid_s)
    assert st_z.count == S and st_z.sum == st_s.sum and st_z.hist[:] == st_s.hist[:]
    # da = None and no per-sample output
    _, st_n = engine.mc_static(controls, dt, gate, gate_count, pfq, None, ppsi, want_fid=False)
    _, st_m = engine.mc_static(controls, dt, gate, gate_count, fq, None, psi0, want_fid=False)
    assert st_n.sum == st_m.sum and st_n.count == S
    # a misaligned fidelity buffer (8-byte offset) still works (row stores instead of the 16-byte pair store)
    raw = engine.pinned_empty(S * (gate_count + 1) + 1)
    off = raw[1:].reshape(S, gate_count + 1)
    engine.mc_static(controls, dt, gate, gate_count, pfq, pda, ppsi, fid_out=off)
    assert np.array_equal(off, fid_s)
    engine.set_option("RBQ_MC_ZEROCOPY", "0")
    before = engine.launch_count
    fid_c, _ = engine.mc_static(controls, dt, gate, gate_count, pfq, pda, ppsi, fid_out=pfid)
    assert np.array_equal(fid_c, fid_s)
    assert engine.launch_count - before >= launches_zero_copy


@pytest.mark.parametrize("gate_parallel", ["0", "1"], ids=["sequential", "gate_parallel"])
def test_noise_table_class_equals_polynomial_classes(engine, oracle, gate_parallel, monkeypatch):
    """The noise sweeps take the knot-table step wherever the noise keeps eps inside the table's range (decided when the
    noise index changes), the polynomial step otherwise; both agree to rounding and each matches the oracle."""
    engine.set_option("RBQ_NOISE_GATE_PARALLEL", gate_parallel)
    controls, dt, gate = PULSES["sin301_dt0.1"]
    S, gc, ndi = 700, 4, 10.0
    _, _, psi0 = _samples(S, seed=17)
    nlen = int(np.ceil(controls.shape[0] * dt * gc * ndi)) + 1
    noise = engine.pink_noise(3, S, 9, spin.NAMP_PREFACTOR, ndi, nlen)
    noise[::7, 30:60] += 5e-3                      # bursts that leave the table's range (but not the polynomial's)
    fid_t, _ = engine.mc_noise(controls, dt, gate, gc, spin.FQ, noise, ndi, psi0)
    engine.set_option("RBQ_NO_TABLE_CLASS", "1")
    fid_p, _ = engine.mc_noise(controls, dt, gate, gc, spin.FQ, noise, ndi, psi0)
    engine.set_option("RBQ_NO_TABLE_CLASS", None)
    assert np.max(np.abs(fid_t - fid_p)) < 5e-14 and not np.array_equal(fid_t, fid_p)
    ref = oracle.mc_noise(controls, dt, gate, gc, spin.FQ, noise[:200], ndi, psi0[:200])
    assert np.max(np.abs(fid_t[:200] - ref)) < FID_RTOL and np.max(np.abs(fid_p[:200] - ref)) < FID_RTOL


# ---- run_sim_deqjl: the piecewise-constant ODE solved exactly (spin.jl:424-467, 1277-1380) ------------------------------------
def test_mc_timeline_matches_oracle(engine, oracle):
    rng = np.random.default_rng(31)
    controls = spin.controls_ypiby2(100.0)
    amps, durs, nidx, save_after = spin.deqjl_timeline(controls, 100.0, spin.TTOT_YPIBY2_JL, 4, 10.0)
    S = 333
    fq, da, psi0 = _samples(S, seed=41)
    noise = spin.NAMP_PREFACTOR * 20 * rng.standard_normal((S, int(nidx.max()) + 1))
    fid, st = engine.mc_timeline(amps, durs, save_after, rbq.GATE_YPIBY2, fq, psi0, da=da, noise=noise, noise_idx=nidx)
    ref = oracle.mc_timeline(amps, durs, nidx, save_after, rbq.GATE_YPIBY2, fq, da, noise, psi0)
    assert fid.shape == (S, 5) and np.max(np.abs(fid - ref)) < 5 * FID_RTOL       # 4 x 1948 chained Pade exponentials on the oracle side
    assert st.count == S and st.min == (1.0 - fid[:, -1]).min()
    # static form (no noise), and argument errors
    fid0, _ = engine.mc_timeline(amps, durs, save_after, rbq.GATE_YPIBY2, fq, psi0)
    ref0 = oracle.mc_timeline(amps, durs, None, save_after, rbq.GATE_YPIBY2, fq, None, None, psi0)
    assert np.max(np.abs(fid0 - ref0)) < 5 * FID_RTOL
    with pytest.raises(Exception):
        engine.mc_timeline(amps, durs, save_after[::-1].copy(), rbq.GATE_YPIBY2, fq, psi0)
    with pytest.raises(Exception):
        engine.mc_timeline(amps, durs, save_after, rbq.GATE_YPIBY2, fq, psi0, noise=noise[:, :10], noise_idx=nidx)


def test_run_sim_deqjl_agrees_with_run_sim_prop_where_the_grids_coincide(engine):
    """When gate_time is a whole number of control knots the ODE's exact solution IS what integrate_prop_schroed! /
    schroedda! compute (the same product of knot exponentials), so the two entry points must agree; with the analytic Y/2
    pulse (gate time 19.4736 ns on a 0.01 ns grid) the ODE form wraps 0.0036 ns into the next period, as the reference does."""
    controls = spin.pulse_sin(301)[:-1]
    seeds = np.arange(6)
    a = spin.run_sim_deqjl(3, spin.GateType.xpiby2, controls, 10.0, 30.0, fq=1.01 * spin.FQ, seed=seeds, engine=engine)
    b = spin.run_sim_prop(3, spin.GateType.xpiby2, controls, 10.0, fq=1.01 * spin.FQ, seed=seeds, engine=engine)
    assert np.max(np.abs(a["fidelities"] - b["fidelities"])) < 1e-12
    # (With noise the two reference paths differ by construction: integrate_prop_schroedda! indexes the noise with a time
    # accumulated by repeated `time += dt` (spin.jl:476-480), which falls just short of a grid point now and then and so reads
    # the previous noise sample, while the ODE form indexes with the exact time.  Both are mirrored as they are: mc_noise
    # against its oracle, the timeline against its own -- so here only the noisy run's sanity is checked.)
    a = spin.run_sim_deqjl(2, spin.GateType.xpiby2, controls, 10.0, 30.0, seed=seeds, dynamics_type=spin.DynamicsType.schroedda, engine=engine)
    b = spin.run_sim_prop(2, spin.GateType.xpiby2, controls, 10.0, seed=seeds, dynamics_type=spin.DynamicsType.schroedda, engine=engine)
    assert a["fidelities"].shape == (6, 3) and np.max(np.abs(a["fidelities"] - b["fidelities"])) < 1e-3
    y = spin.run_sim_deqjl(1, spin.GateType.ypiby2, spin.controls_ypiby2(100.0), 100.0, spin.TTOT_YPIBY2_JL, seed=seeds, engine=engine)
    e = 1 - y["fidelities"][:, -1]
    assert np.all(e < 1e-4) and np.all(e >= -1e-13)            # the sampled Y/2 pulse (zeroed end samples, wrap) is a Y/2 gate to ~1e-5


def test_host_register_enables_zero_copy_for_caller_owned_arrays(engine):
    """rbq_host_register page-locks memory the caller already owns (a Julia Array, a numpy array): the sweep then takes the
    zero-copy form (one sweep kernel instead of one per chunk) and returns the same bits (both forms through the shared-memory
    kernel, RBQ_MC_VARIANT 161)."""
    controls, dt, gate = PULSES["sin301_dt0.1"]
    S = 300000                                             # several one-wave chunks in the streamed form
    fq, da, psi0 = _samples(S, seed=51)
    fid = np.empty((S, 2))
    engine.set_option("RBQ_MC_VARIANT", "161")
    try:
        _host_register_zero_copy(engine, controls, dt, gate, S, fq, da, psi0, fid)
    finally:
        engine.set_option("RBQ_MC_VARIANT", "")


def _host_register_zero_copy(engine, controls, dt, gate, S, fq, da, psi0, fid):
    before = engine.launch_count
    f_s, st_s = engine.mc_static(controls, dt, gate, 1, fq, da, psi0, fid_out=fid)
    streamed = engine.launch_count - before
    keep = f_s.copy()
    for a in (fq, da, psi0, fid):
        engine.host_register(a)
    try:
        before = engine.launch_count
        f_z, st_z = engine.mc_static(controls, dt, gate, 1, fq, da, psi0, fid_out=fid)
        zero_copy = engine.launch_count - before
    finally:
        for a in (fq, da, psi0, fid):
            engine.host_unregister(a)
    assert np.array_equal(f_z, keep) and st_z.sum == st_s.sum
    assert zero_copy < streamed


# ---- dephasing (T2) and T1+T2 with the DRIVE ON: dynamics_lindbladt2_deqjl (spin.jl:542-554) -----------------------------
T2_CASES = {
    # name: (channels, gamma_c, gamma_f)
    "t2_const": (rbq.LINDBLAD_T2, spin.GAMMAC, 0.0),
    "t2_reference_rate": (rbq.LINDBLAD_T2, spin.GAMMAC, spin.GAMMAF_PREFACTOR),
    "t2_rate_x30": (rbq.LINDBLAD_T2, spin.GAMMAC, 30 * spin.GAMMAF_PREFACTOR),
    "t1t2": (rbq.LINDBLAD_T1 | rbq.LINDBLAD_T2, spin.GAMMAC, spin.GAMMAF_PREFACTOR),
}


@pytest.mark.parametrize("case", list(T2_CASES))
@pytest.mark.parametrize("name", ["sin301_dt0.1", "sin1001_dt0.1", "y2_dt0.01"])
@pytest.mark.parametrize("with_da", [False, True], ids=["shared", "offsets"])
def test_lindblad_t2_with_drive_matches_ode_oracle(engine, oracle, case, name, with_da):
    """The oracle is the exact (long-double Taylor) solution of the ODE the reference gives to Vern9 at reltol 1e-12
    (oracle.lindblad_knots -> orc_lindblad_ode); the engine uses one Magnus step per knot in the Bloch picture.
    Tolerance: 5e-12 on densities and fidelities against mixed targets (the Lindblad bar of DESIGN.md §3)."""
    controls, dt, gate = PULSES[name]
    ch, gc, gf = T2_CASES[case]
    S, gate_count = 64, 3
    rho0 = _rand_rho(S, 21, mixed=True)
    da = 1e-3 * np.random.default_rng(5).standard_normal(S) if with_da else None
    fid, rho, st = engine.lindblad(controls, dt, gate, gate_count, spin.FQ, da, rho0, channels=ch, t2_gamma_c=gc, t2_gamma_f=gf)
    rfid, rrho = oracle.lindblad_knots(controls, dt, gate, gate_count, spin.FQ, da, rho0, channels=ch, gamma_c=gc, gamma_f=gf)
    assert np.max(np.abs(rho - rrho)) < 5e-12, (case, name)
    assert np.max(np.abs(fid - rfid)) < 5e-12, (case, name)
    assert st.count == S and np.allclose(np.trace(rho, axis1=1, axis2=2), 1.0, atol=1e-14)


def test_lindblad_t2_matches_50_digit_golden(engine):
    """tests/golden/lindblad_t2.npz: 50-digit solution of the reference's T2 / T1+T2 ODEs with the drive on
    (tests/golden/make_golden_t2.py)."""
    z = np.load(GOLDEN + "/lindblad_t2.npz")
    G = int(z["gate_count"])
    for name, ch in (("t2", rbq.LINDBLAD_T2), ("t1t2", rbq.LINDBLAD_T1 | rbq.LINDBLAD_T2), ("t2c", rbq.LINDBLAD_T2), ("t2big", rbq.LINDBLAD_T2)):
        fid, rho, _ = engine.lindblad(z["controls"], float(z["dt"]), int(z["gate"]), G, float(z["fq"]), z["da"], z["rho0"],
                                      channels=ch, t2_gamma_c=float(z["gamma_c"]), t2_gamma_f=float(z["gf_" + name]))
        # one fourth-order Magnus step per knot: 2e-13 at the reference's rate; the 30x case (rate slope 900x) sits at the Lindblad bar
        assert np.max(np.abs(rho - z["rho_" + name][:, G])) < (5e-12 if name == "t2big" else 2e-13), name


def test_run_sim_prop_lindbladt2_uses_the_reference_rate(engine, oracle):
    """run_sim_prop(dynamics_type=lindbladt2) must integrate Gamma(t) = GAMMAC + 2 GAMMAF_PREFACTOR^2 t (spin.jl:549-552), not
    GAMMAC alone."""
    controls, dt, gate = PULSES["sin301_dt0.1"]
    seeds = np.arange(5)
    res = spin.run_sim_prop(4, gate, controls, 1 / dt, seed=seeds, dynamics_type=spin.DynamicsType.lindbladt2, engine=engine)
    psi0 = np.stack([spin.gen_rand_state_iso(int(s), engine) for s in seeds])
    rho0 = np.stack([np.outer(spin.get_vec_uniso(p), spin.get_vec_uniso(p).conj()) for p in psi0])
    rfid, _ = oracle.lindblad_knots(controls, dt, gate, 4, spin.FQ, None, rho0, channels=2, gamma_c=spin.GAMMAC,
                                    gamma_f=spin.GAMMAF_PREFACTOR)
    # pure targets: the sqrt-fidelity is sqrt(eps)-conditioned (see test_lindblad_t1_matches_oracle)
    assert np.max(np.abs(res["fidelities"] - rfid)) < 1e-8
    rfid0, _ = oracle.lindblad_knots(controls, dt, gate, 4, spin.FQ, None, rho0, channels=2, gamma_c=spin.GAMMAC, gamma_f=0.0)
    assert np.max(np.abs(rfid - rfid0)) > 1e-7      # the dropped term would have been visible


# ---- evaluate_fqdev / evaluate_adev (spin.jl:1611-1672) against a loop of oracle run_sim_prop calls ----------------------
def test_evaluate_fqdev_matches_serial_oracle_loop(engine, oracle):
    controls, dt = spin.controls_zpiby2(100.0), 0.01
    trials, fq_cov = 60, spin.FQ * 1e-2
    res = spin.evaluate_fqdev(controls, 1 / dt, fq_cov=fq_cov, trial_count=trials, gate_type=spin.GateType.zpiby2, engine=engine)
    ge, geb = np.empty(2 * trials), np.empty(8)
    for i in range(-4, trials):                                       # the reference's two loops, one call per (state, sign)
        psi0 = spin.gen_rand_state_iso(i, engine)[None] if i < 0 else oracle.philox_samples(i, 1, 0, spin.FQ, 0.0, 0.0, 0.0)[2]
        for j, fq in enumerate((spin.FQ + fq_cov, spin.FQ - fq_cov)):
            f = oracle.mc_static(controls, dt, rbq.GATE_ZPIBY2, 1, np.array([fq]), None, psi0)
            (geb if i < 0 else ge)[2 * (i + 4 if i < 0 else i) + j] = 1 - f[0, -1]
    assert np.max(np.abs(res["gate_errors"] - ge)) < FID_RTOL
    assert np.max(np.abs(res["gate_errors_basis"] - geb)) < FID_RTOL
    # paper main.tex:857-859: analytic Z/2 at 1 % detuning has a gate error of about 4.7e-5 (state average)
    assert abs(np.mean(res["gate_errors"]) - 4.7e-5) < 1.0e-5


def test_evaluate_adev_matches_serial_oracle_loop(engine, oracle):
    controls, dt = spin.controls_xpiby2(100.0), 0.01
    trials, gates = 5, 3
    res = spin.evaluate_adev(controls, 1 / dt, gate_count=gates, gate_type=spin.GateType.xpiby2, trial_count=trials, engine=engine)
    nlen = int(np.ceil(controls.shape[0] * dt * gates * spin.DT_NOISE_INV)) + 1
    ge = np.empty(trials)
    for k, seed in enumerate(range(1, trials + 1)):                   # spin.jl:1662-1668: seed = 1:trial_count
        noise = oracle.pink_noise(seed, 1, 0, spin.NAMP_PREFACTOR, spin.DT_NOISE_INV, nlen)
        psi0 = oracle.philox_samples(seed, 1, 0, spin.FQ, 0.0, 0.0, 0.0)[2]
        f = oracle.mc_noise(controls, dt, rbq.GATE_XPIBY2, gates, spin.FQ, noise, spin.DT_NOISE_INV, psi0)
        ge[k] = 1 - f[0, -1]
    assert np.max(np.abs(res["gate_errors"] - ge)) < 5e-12            # 3 x 5680 chained Pade exponentials on the oracle side


# ---- Lindblad sweeps with realistic per-realisation offsets: the per-knot Frechet table path ---------------------------------
@pytest.mark.parametrize("name", ["sin301_dt0.1", "sin1001_dt0.1", "y2_dt0.01", "ragged_2049"])
@pytest.mark.parametrize("gate_count", [1, 3, 6], ids=["g1", "g3", "g6compose"])
@pytest.mark.parametrize("channels", [rbq.LINDBLAD_T1, rbq.LINDBLAD_T1 | rbq.LINDBLAD_T2], ids=["t1", "t1t2c"])
def test_lindblad_table_path_matches_series_path_and_oracle(engine, oracle, name, gate_count, channels):
    """Offsets of the size the reference's noise model gives (2.574e-5 GHz) take the table step (E_k + d L1_k + d^2 L2_k +
    d^3 L3_k per knot); a few realisations far outside the table's range share warps with them and take the series step, as
    do the knots next to a kink of the T1 table.  Both forms must agree to rounding, and with the oracle at the Lindblad bar."""
    controls, dt, gate = PULSES[name]
    S = 777                                                   # ragged: 24 full warps + 9 threads
    rho0 = _rand_rho(S, 31, mixed=True)
    rng = np.random.default_rng(6)
    da = 2.574e-5 * rng.standard_normal(S)
    da[5::97] = 3e-3 * rng.standard_normal(da[5::97].shape[0])    # out-of-range realisations: their warps fall back
    gc = spin.GAMMAC if channels & rbq.LINDBLAD_T2 else 0.0
    fid, rho, st = engine.lindblad(controls, dt, gate, gate_count, spin.FQ, da, rho0, channels=channels, t2_gamma_c=gc)
    engine.set_option("RBQ_LINDBLAD_TABLE", "0")
    fid2, rho2, st2 = engine.lindblad(controls, dt, gate, gate_count, spin.FQ, da, rho0, channels=channels, t2_gamma_c=gc)
    nk = controls.shape[0]
    assert np.max(np.abs(rho - rho2)) < max(2e-13, 1e-16 * nk * gate_count), name
    assert np.max(np.abs(fid - fid2)) < max(1e-12, 1e-15 * nk * gate_count), name
    assert st.count == st2.count == S and abs(st.sum - st2.sum) < 1e-11 * S
    sub = slice(0, 120)
    rfid, rrho = oracle.lindblad_knots(controls, dt, gate, gate_count, spin.FQ, da[sub], rho0[sub], channels=channels, gamma_c=gc)
    assert np.max(np.abs(rho[sub] - rrho)) < 5e-12 and np.max(np.abs(fid[sub] - rfid)) < 5e-12, name


def test_lindblad_table_path_matches_50_digit_golden(engine):
    z = np.load(GOLDEN + "/lindblad_t2.npz")
    G = int(z["gate_count"])
    for name, ch, gc in (("t1_small", rbq.LINDBLAD_T1, 0.0), ("t1t2c_small", rbq.LINDBLAD_T1 | rbq.LINDBLAD_T2, float(z["gamma_c"]))):
        for opt in (None, "0"):
            engine.set_option("RBQ_LINDBLAD_TABLE", opt)
            fid, rho, _ = engine.lindblad(z["controls"], float(z["dt"]), int(z["gate"]), G, float(z["fq"]), z["da_small"], z["rho0"],
                                          channels=ch, t2_gamma_c=gc)
            assert np.max(np.abs(rho - z["rho_" + name][:, G])) < 1e-13, (name, opt)


def test_lindblad_table_path_full_size_balanced_grid(engine, oracle):
    """BASELINE configs[3] at full size (1e5 realisations, P_sin(1001)): spot rows against the oracle, every fidelity finite and
    in [0, 1], and the statistics equal those of the returned fidelities."""
    controls, dt, gate = PULSES["sin1001_dt0.1"]
    S = 100_000
    rng = np.random.default_rng(0)
    psi = rng.random((S, 2)) + 1j * rng.random((S, 2))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    rho0 = np.einsum("si,sj->sij", psi, psi.conj())
    da = 2.574e-5 * rng.standard_normal(S)
    fid, rho, st = engine.lindblad(controls, dt, rbq.GATE_XPIBY2, 1, spin.FQ, da, rho0)
    assert st.count == S and np.all(np.isfinite(fid)) and fid.min() >= 0.0 and fid.max() <= 1.0 + 1e-12
    e = 1.0 - fid[:, -1]
    assert abs(st.sum - e.sum()) <= 1e-12 * abs(e.sum()) + 1e-15 and st.min == e.min() and st.max == e.max()
    pick = np.concatenate([np.arange(0, 64), np.arange(S - 64, S), rng.integers(0, S, 128)])
    _, _, rrho = oracle.lindblad_t1(controls, dt, rbq.GATE_XPIBY2, 1, spin.FQ, da[pick], rho0[pick])
    assert np.max(np.abs(rho[pick] - rrho)) < 5e-12


# ---- run_sim_deqjl for the density dynamics and run_sim_fine_deqjl (spin.jl:488-518, 542-554, 1277-1478) -----------------------
@pytest.mark.parametrize("dyn,ch", [(spin.DynamicsType.lindbladnodis, 0), (spin.DynamicsType.lindbladt1, 1), (spin.DynamicsType.lindbladt2, 2)],
                         ids=["lindbladnodis", "lindbladt1", "lindbladt2"])
def test_run_sim_deqjl_density_dynamics_match_the_ode_oracle(engine, oracle, dyn, ch):
    """gate_time is not a whole number of control knots, so the reference's `mod control_knot_count` wrap-around and the cuts at
    the save times are both exercised; seed 0 is the reference's fixed |+> density."""
    controls, cdi, gate_time, gates = spin.pulse_sin(241)[:-1], 10.0, 23.87, 3
    seeds = np.array([0, 1, 2, 7])
    res = spin.run_sim_deqjl(gates, spin.GateType.ypiby2, controls, cdi, gate_time, seed=seeds, dynamics_type=dyn, engine=engine)
    amps, durs, _, save_after = spin.deqjl_timeline(controls, cdi, gate_time, gates)
    rho0 = np.stack([spin.gen_rand_density_iso(int(s), engine) for s in seeds])
    t2 = ch == 2
    rfid, rrho = oracle.lindblad_ode(amps, durs, save_after, rbq.GATE_YPIBY2, spin.FQ, None, rho0, channels=ch,
                                     gamma_c=spin.GAMMAC if t2 else 0.0, gamma_f=spin.GAMMAF_PREFACTOR if t2 else 0.0)
    assert np.allclose(rho0[0], 0.5 * np.ones((2, 2)))
    assert np.max(np.abs(res["densities"] - rrho)) < 5e-12
    assert np.max(np.abs(res["fidelities"] - rfid)) < 1e-8       # pure targets: sqrt(eps) conditioning of the sqrt-fidelity
    if ch == 0:     # no dissipation: purity is conserved
        pur = np.einsum("sgij,sgji->sg", res["densities"], res["densities"]).real
        assert np.max(np.abs(pur - 1.0)) < 1e-12


def test_run_sim_fine_deqjl_states_on_the_save_step_grid(engine, oracle):
    """run_sim_fine_deqjl (Fig. 4b driver): iso states / densities every save_step along the pulse, against the oracle's
    segment walk (states) and exact ODE solution (densities)."""
    controls, cdi, gate_time = spin.controls_ypiby2(100.0), 100.0, spin.TTOT_YPIBY2_JL
    seeds = np.array([1, 2, 3])
    for dyn in (spin.DynamicsType.schroed, spin.DynamicsType.schroedda):
        out = spin.run_sim_fine_deqjl(controls, cdi, gate_time, gate_count=1, seed=seeds, dynamics_type=dyn, save_step=0.1, engine=engine)
        nsave = out["save_times"].shape[0] - 1
        assert nsave == int(np.floor(gate_time / 0.1)) and out["states"].shape == (3, nsave + 1, 4)
        noisy = dyn == spin.DynamicsType.schroedda
        amps, durs, nidx, save_after = spin.deqjl_timeline(controls, cdi, gate_time, 1, 10.0 if noisy else None, out["save_times"][1:])
        psi0 = np.stack([spin.gen_rand_state_iso(int(s), engine) for s in seeds])
        noise = None
        if noisy:
            nlen = int(np.ceil(gate_time * 10.0)) + 1
            noise = np.concatenate([oracle.pink_noise(int(s), 1, 0, spin.NAMP_PREFACTOR, 10.0, nlen) for s in seeds])
        _, rstates = oracle.mc_timeline(amps, durs, nidx, save_after, rbq.GATE_XPIBY2, np.full(3, spin.FQ), None, noise, psi0, want_states=True)
        assert np.max(np.abs(out["states"] - rstates)) < 2e-12, dyn
        assert np.max(np.abs(np.linalg.norm(out["states"], axis=2) - 1.0)) < 1e-13
    out = spin.run_sim_fine_deqjl(controls, cdi, gate_time, gate_count=2, seed=seeds, dynamics_type=spin.DynamicsType.lindbladt1, save_step=0.5,
                                  engine=engine)
    amps, durs, _, save_after = spin.deqjl_timeline(controls, cdi, gate_time, 2, None, out["save_times"][1:])
    rho0 = np.stack([spin.gen_rand_density_iso(int(s), engine) for s in seeds])
    _, rrho = oracle.lindblad_ode(amps, durs, save_after, rbq.GATE_XPIBY2, spin.FQ, None, rho0, channels=1)
    assert out["states"].shape == rrho.shape and np.max(np.abs(out["states"] - rrho)) < 5e-12


def test_single_process_device_group_equals_one_handle(engine):
    """rbq_group: one process driving several handles (here every visible GPU, and the first one twice more so that the
    sharding is exercised on a single-GPU box too).  Samples are keyed by their global index, so the group's fidelities equal
    one handle's to rounding and the merged statistics up to the order of the partial sums."""
    import torch

    controls, dt, gate = PULSES["sin301_dt0.1"]
    devs = list(range(torch.cuda.device_count())) + [0, 0]
    grp = rbq.EngineGroup(devs)
    try:
        first, S, seed = 1234, 10_007, 5        # not divisible by the member count
        fid_g, st_g = grp.mc_static_philox(controls, dt, gate, 2, first, S, seed, spin.FQ, 0.01, 0.03, 2.574e-5, want_fid=True)
        fid_1, st_1 = engine.mc_static_philox(controls, dt, gate, 2, first, S, seed, spin.FQ, 0.01, 0.03, 2.574e-5, want_fid=True)
        # same samples whatever the sharding; which warp a sample shares (and so which of the equally accurate step classes
        # it takes) depends on where its shard starts, hence rounding-level differences only
        assert np.max(np.abs(fid_g - fid_1)) < 1e-14
        assert st_g.count == st_1.count == S and abs(st_g.min - st_1.min) < 1e-14 and abs(st_g.max - st_1.max) < 1e-14
        assert np.abs(np.array(st_g.hist[:]) - np.array(st_1.hist[:])).sum() <= 4      # a sample on a bin edge may move
        assert abs(st_g.sum - st_1.sum) < 1e-12 * st_1.sum
        _, st_e = grp.mc_static_philox(controls, dt, gate, 2, 0, 1, seed, spin.FQ, 0.01, 0.03, 2.574e-5)     # fewer samples than members
        assert st_e.count == 1
    finally:
        grp.close()


# ---- constant-memory form of the static sweep (mc_static_ctab_kernel + mc_static_finish_kernel) ------------------------------
@pytest.mark.parametrize("variant", ["3161", "2161", "2181", "42161", "43161", "1161"])
def test_constant_memory_form_matches_shared_memory_form_and_oracle(engine, oracle, variant):
    """Table rows as uniform-register operands from constant memory (variants 3xxx, the default: Horner with c1 through shared
    memory -- bit for bit the shared-memory kernel; 2xxx: the quadratic factored as c2 (eps (eps + rho) + sigma); 1xxx: Horner with
    vector loads) against the rows through shared memory (161), the oracle, and the 50-digit golden; ragged sample count, a knot
    count that is no multiple of the batch, recorded gates 1 and 3."""
    controls, dt, gate = PULSES["sin301_dt0.1"]
    controls = controls[:297]                                 # 297 = 37 batches of 8 + 1: the tail loop runs
    S = 5000 + 19
    fq, da, psi0 = _samples(S, seed=77)
    try:
        for gates in (1, 3):
            engine.set_option("RBQ_MC_VARIANT", "161")
            f0, s0 = engine.mc_static(controls, dt, gate, gates, fq, da, psi0)
            engine.set_option("RBQ_MC_VARIANT", variant)
            before = engine.launch_count
            f1, s1 = engine.mc_static(controls, dt, gate, gates, fq, da, psi0)
            assert engine.launch_count - before > 0
            assert np.max(np.abs(f1 - f0)) < 2e-14 and s1.count == s0.count == S and abs(s1.sum - s0.sum) <= 1e-12 * s0.sum
            if variant[-4] in "13":                         # Horner forms: the same operations as the shared-memory kernel
                assert np.array_equal(f1, f0) and s1.sum == s0.sum and s1.hist[:] == s0.hist[:]
            assert sum(s1.hist[:]) == S
            ref = oracle.mc_static(controls, dt, gate, gates, fq[:400], da[:400], psi0[:400])
            assert np.max(np.abs(f1[:400] - ref)) < FID_RTOL
        z = np.load(GOLDEN + "/mc_static.npz")
        fg, _ = engine.mc_static(z["controls"], float(z["dt"]), int(z["gate"]), int(z["gate_count"]), z["fq"], z["da"], z["psi0"])
        assert np.max(np.abs(fg - z["fid"])) < 1e-13
    finally:
        engine.set_option("RBQ_MC_VARIANT", "")


def test_constant_memory_table_is_shared_correctly_between_handles(engine):
    """The constant-memory table is one per device: two handles with different pulses, called alternately on their own
    streams, each get their own pulse's result (owner tag + event in launch_mc_static_ctab)."""
    other = rbq.Engine(0)
    try:
        c1, dt, gate = PULSES["sin301_dt0.1"]
        c2 = 0.5 * c1[::-1].copy()
        S = 3000
        fq, da, psi0 = _samples(S, seed=78)
        a1, _ = engine.mc_static(c1, dt, gate, 1, fq, da, psi0)
        b1, _ = other.mc_static(c2, dt, gate, 1, fq, da, psi0)
        assert not np.array_equal(a1, b1)
        for _ in range(3):
            a2, _ = engine.mc_static(c1, dt, gate, 1, fq, da, psi0)
            b2, _ = other.mc_static(c2, dt, gate, 1, fq, da, psi0)
            assert np.array_equal(a2, a1) and np.array_equal(b2, b1)
        # the same handle with a new pulse of the same length: the table is uploaded again
        a3, _ = engine.mc_static(c2, dt, gate, 1, fq, da, psi0)
        assert np.array_equal(a3, b1)
    finally:
        other.close()


def test_constant_memory_form_windows_of_a_long_pulse(engine, oracle):
    """A pulse longer than the constant-memory table (2000 knots) is swept in windows of 1792 knots, one launch each, the
    quaternion travelling through scratch; blocks with out-of-range warps walk the whole pulse in the first launch."""
    nk = 2 * 1792 + 517                                      # three windows, the last one partial and no multiple of 8
    t = np.arange(nk) * 0.01
    controls = 0.02 * np.sin(2 * np.pi * t / (nk * 0.01)) + 0.004 * np.sin(2 * np.pi * 7 * t / (nk * 0.01))
    S = 3000 + 5
    fq, da, psi0 = _samples(S, seed=79)
    da[np.arange(S) % 640 < 32] += 0.2                       # every 20th warp leaves the table's range: its block takes the tile path
    try:
        engine.set_option("RBQ_MC_VARIANT", "161")
        f0, s0 = engine.mc_static(controls, 0.01, rbq.GATE_XPIBY2, 2, fq, da, psi0)
        engine.set_option("RBQ_MC_VARIANT", "3161")
        before = engine.launch_count
        f1, s1 = engine.mc_static(controls, 0.01, rbq.GATE_XPIBY2, 2, fq, da, psi0)
        assert engine.launch_count - before >= 3 + 1           # three sweep launches + the finish kernel (+ table, statistics)
    finally:
        engine.set_option("RBQ_MC_VARIANT", "")
    assert np.array_equal(f1, f0) and s1.count == S and s1.sum == s0.sum     # Horner form: the same operations, window by window
    ref = oracle.mc_static(controls, 0.01, rbq.GATE_XPIBY2, 2, fq[:200], da[:200], psi0[:200])
    assert np.max(np.abs(f1[:200] - ref)) < max(FID_RTOL, 2.5e-15 * nk)


def test_constant_memory_form_seed_in_call_equals_sample_arrays_across_windows(engine):
    """rbq_mc_static_philox (samples drawn in the kernel from their global Philox index) against rbq_mc_static on the arrays
    rbq_philox_samples returns, on a pulse that needs several constant-memory windows and with a non-zero first index: the same
    fidelities bit for bit, the same statistics."""
    controls, dt, gate = PULSES["x2_dt0.01"]                  # 5680 knots: four windows
    first, S, seed = 12345, 3000 + 7, 99
    fq, da, psi0 = engine.philox_samples(first, S, seed, spin.FQ, 0.01, 0.03, 2.574e-5)
    f_a, st_a = engine.mc_static(controls, dt, gate, 2, fq, da, psi0)
    f_p, st_p = engine.mc_static_philox(controls, dt, gate, 2, first, S, seed, spin.FQ, 0.01, 0.03, 2.574e-5, want_fid=True)
    assert np.array_equal(f_a, f_p)
    assert st_a.count == st_p.count == S and st_a.sum == st_p.sum and st_a.hist[:] == st_p.hist[:]
