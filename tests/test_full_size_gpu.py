"""GPU tests at BASELINE.json's full sizes (the oracle cannot run these in seconds), through size-independent
properties of the domain: unitarity, exact analytic gates, global-phase invariance, sharding invariance of the
additive statistics, bit-reproducibility, and for the Lindblad map trace preservation / contraction."""
import numpy as np
import pytest

import rbqoc_b200 as rbq
from rbqoc_b200 import spin

pytestmark = pytest.mark.gpu

S_FULL = 1 << 20          # the bench's 1M-sample sweep
NK = 1000


@pytest.fixture(scope="module")
def sweep(engine):
    controls = spin.pulse_sin(NK + 1)[:-1]
    fq, da, psi0 = engine.philox_samples(0, S_FULL, 2026, spin.FQ, 0.01, 0.03, 2.574e-5)
    fid, st = engine.mc_static(controls, 0.1, rbq.GATE_XPIBY2, 1, fq, da, psi0)
    return controls, fq, da, psi0, fid, st


def test_full_sweep_statistics_are_those_of_the_fidelities(sweep):
    controls, fq, da, psi0, fid, st = sweep
    assert fid.shape == (S_FULL, 2)
    assert np.all(np.abs(fid[:, 0] - 1.0) < 1e-14)                    # F(psi0, psi0) = |psi0|^4 = 1
    assert np.all(fid[:, 1] <= 1.0 + 1e-13) and np.all(fid[:, 1] >= -1e-13)
    e = 1.0 - fid[:, 1]
    assert st.count == S_FULL and st.nonfinite == 0
    assert abs(st.sum - e.sum()) < 1e-11 * e.sum()
    assert abs(st.sumsq - np.sum(e * e)) < 1e-11 * np.sum(e * e)
    assert st.min == e.min() and st.max == e.max()
    assert sum(st.hist[:]) == S_FULL


def test_full_sweep_is_bit_reproducible(engine, sweep):
    controls, fq, da, psi0, fid, st = sweep
    fid2, st2 = engine.mc_static(controls, 0.1, rbq.GATE_XPIBY2, 1, fq, da, psi0)
    assert np.array_equal(fid, fid2)
    assert st.sum == st2.sum and st.sumsq == st2.sumsq and st.hist[:] == st2.hist[:]


def test_full_sweep_sharding_invariance(engine, sweep):
    """8 contiguous shards drawn by global Philox index merge to the statistics of the whole (what 8 GPUs compute)."""
    controls, fq, da, psi0, fid, st = sweep
    total = None
    for r in range(8):
        first, cnt = r * (S_FULL // 8), S_FULL // 8
        _, part = engine.mc_static_philox(controls, 0.1, rbq.GATE_XPIBY2, 1, first, cnt, 2026, spin.FQ, 0.01, 0.03, 2.574e-5)
        total = part if total is None else rbq.stats_merge(total, part)
    assert total.count == st.count and total.hist[:] == st.hist[:]
    assert total.min == st.min and total.max == st.max
    assert abs(total.sum - st.sum) < 1e-12 * st.sum


def test_full_sweep_global_phase_invariance(engine, sweep):
    """psi0 -> exp(i phi) psi0 is the iso rotation [[c,-s],[s,c]] (x) 1: fidelities cannot change."""
    controls, fq, da, psi0, fid, st = sweep
    phi = 0.7321
    c, s = np.cos(phi), np.sin(phi)
    rot = np.concatenate([c * psi0[:, :2] - s * psi0[:, 2:], s * psi0[:, :2] + c * psi0[:, 2:]], axis=1)
    fid2, _ = engine.mc_static(controls, 0.1, rbq.GATE_XPIBY2, 1, fq, da, rot)
    assert np.max(np.abs(fid2 - fid)) < 1e-13


def test_full_sweep_exact_gate_has_unit_fidelity(engine, sweep):
    """The idle Z/2 of spin.jl:557 at zero detuning, 1M random states, 4 consecutive gates (the 4-cyclic targets)."""
    _, _, _, psi0, _, _ = sweep
    n = 1786
    fid, st = engine.mc_static(np.zeros(n), spin.TTOT_ZPIBY2 / n, rbq.GATE_ZPIBY2, 4, np.full(S_FULL, spin.FQ), None, psi0)
    assert np.max(np.abs(fid - 1.0)) < 2e-13
    assert st.max < 2e-13


def test_full_size_lindblad_map_properties(engine):
    """1e5 realisations (BASELINE configs[3]): trace preserved, Bloch vector inside the ball, purity non-increasing for
    the unital T1 map, fidelities in [0, 1]."""
    S = 100_000
    rng = np.random.default_rng(5)
    psi = rng.random((S, 2)) + 1j * rng.random((S, 2))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    rho0 = np.einsum("si,sj->sij", psi, psi.conj())
    lam = rng.uniform(0.5, 1.0, S)[:, None, None]
    rho0 = lam * rho0 + (1 - lam) * np.eye(2) / 2
    da = 2.574e-5 * rng.standard_normal(S)
    controls = spin.pulse_sin(NK + 1)[:-1]
    fid, rho, st = engine.lindblad(controls, 0.1, rbq.GATE_XPIBY2, 3, spin.FQ, da, rho0)
    assert np.max(np.abs(np.trace(rho, axis1=1, axis2=2) - 1.0)) < 1e-14
    assert np.max(np.abs(rho - rho.conj().transpose(0, 2, 1))) == 0.0
    purity0 = np.real(np.einsum("sij,sji->s", rho0, rho0))
    purity = np.real(np.einsum("sij,sji->s", rho, rho))
    assert np.all(purity <= purity0 + 1e-14)
    assert np.all(purity >= 0.5 - 1e-14)
    assert np.all(fid >= 0.0) and np.all(fid <= 1.0 + 1e-13)
    assert st.count == S


def test_full_sweep_zero_copy_streamed_and_polynomial_forms_agree(engine, sweep, monkeypatch):
    """The 1M-sample sweep through page-locked buffers (zero-copy: one kernel reads / writes over PCIe, table rows through shared
    memory) returns the bits of the streamed form (two kernels, table rows from constant memory: the same 14 operations per knot in
    the same order), whichever kernel the streamed form is given; the all-uniform variant (2161: the quadratic factored as
    c2 (eps (eps + rho) + sigma)) and the polynomial step (table class off) agree with them to rounding."""
    controls, fq, da, psi0, fid, st = sweep
    pfq, pda, ppsi = engine.pinned_empty(S_FULL), engine.pinned_empty(S_FULL), engine.pinned_empty((S_FULL, 4))
    pfid = engine.pinned_empty((S_FULL, 2))
    pfq[:], pda[:], ppsi[:] = fq, da, psi0
    fz, stz = engine.mc_static(controls, 0.1, rbq.GATE_XPIBY2, 1, pfq, pda, ppsi, fid_out=pfid)
    assert np.array_equal(fz, fid) and stz.sum == st.sum and stz.hist[:] == st.hist[:]
    engine.set_option("RBQ_MC_VARIANT", "161")           # the streamed form through the shared-memory kernel: the same bits
    fs, sts = engine.mc_static(controls, 0.1, rbq.GATE_XPIBY2, 1, fq, da, psi0)
    engine.set_option("RBQ_MC_VARIANT", "2161")          # ... through the all-uniform constant-memory variant: rounding
    fu, stu = engine.mc_static(controls, 0.1, rbq.GATE_XPIBY2, 1, fq, da, psi0)
    engine.set_option("RBQ_MC_VARIANT", "")
    assert np.array_equal(fz, fs) and stz.sum == sts.sum and stz.hist[:] == sts.hist[:]
    assert np.max(np.abs(fu - fid)) < 2e-14 and stu.count == S_FULL and abs(stu.sum - st.sum) < 1e-12 * st.sum
    engine.set_option("RBQ_NO_TABLE_CLASS", "1")
    fp, stp = engine.mc_static(controls, 0.1, rbq.GATE_XPIBY2, 1, fq, da, psi0)
    assert np.max(np.abs(fp - fid)) < 2e-14 and stp.count == S_FULL
    assert abs(stp.sum - st.sum) < 1e-12 * st.sum


def test_full_size_backward_pass(engine):
    """1001 knots, n = 56 (the unscented model's dimensions), 6 problems: gains and value decrease against the numpy
    restatement; a batch returns the same bits for the same problem wherever it sits in the batch."""
    from oracle import riccati
    from test_oracle_cpu import _lq_problem

    rng = np.random.default_rng(8)
    N, n, m = 1001, 56, 1
    A, Bm, AB, Q, q, R, Eux, H, r = _lq_problem(rng, N, n, m)
    A *= 0.55                                            # rho(A) < 1: a 1000-knot recursion that stays well conditioned
    AB = np.concatenate([A, Bm], axis=2)
    st = lambda x: np.stack([x] * 6)
    K, d, dV, fail = engine.backward_pass(st(AB), st(Q), st(q), st(R), st(Eux), st(r), rho=1e-2)
    Kr, dr, dVr, failr = riccati.backward_pass(AB, Q, q, R, Eux, r, rho=1e-2)
    assert list(fail) == [-1] * 6 and failr == -1
    rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    assert rel(K[0], Kr) < 1e-10 and rel(d[0], dr) < 1e-10 and np.allclose(dV[0], dVr, rtol=1e-10)
    for b in range(1, 6):
        assert np.array_equal(K[b], K[0]) and np.array_equal(d[b], d[0]) and np.array_equal(dV[b], dV[0])
