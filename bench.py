#!/usr/bin/env python
"""bench.py — FP64 trajectory-steps/s of the Monte-Carlo robustness rollout (BASELINE.json config 5,
the configuration the north-star target is quoted on: "the 1M-sample robustness rollout").

One step = one pass of the hot path over one batch: a fixed pulse P_sin(1001) (1000 control knots,
dt = 0.1 ns) rolled over S = 2^20 (detuning, amplitude-error, initial-state) samples PER GPU, one gate,
fidelity per sample + infidelity statistics.  A trajectory-step is one 4-real iso state advanced by
one knot, so a step is S * 1000 trajectory-steps per GPU.

  value      device-resident inputs, rbq_mc_static_dev + (N>1) the one all-gather of rbq_stats
  e2e        the reference-facing call: seeds in, per-sample fidelities + statistics out in HOST memory
             (rbq_mc_static_philox, the shape of run_sim_prop(...; seed, state_seed), spin.jl:1553-1608); the second
             form, explicit host sample arrays through rbq_mc_static (48 B/sample up), is reported beside it
  parity     the first 10 000 samples of this very sweep against the CPU oracle, element-wise (tol 1e-12); a miss fails the run
  roofline   FP64 pipe: flops the SU(2) kernel executes per trajectory-step / measured DFMA peak
  configs    BASELINE configs[0..3] (spin13 / spin12 / spin30 rollout + Jacobians, Lindblad 1e5) each with its own value, e2e,
             roofline and CPU baseline (all host threads and one thread)
  strong_scaling  S_total = 1e6 / 1e7 / 1e8 samples split over the ranks of this run (seed-in form, statistics only)
  cpu_baseline / --impl reference: the CPU restatement of the reference algorithm (the Julia reference cannot run in this
             image or on the GPU box: `julia` is absent on both), built -O3 -march=native on the box, all host threads
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NK = 1000            # control knots of P_sin(1001)
DT = 0.1
S_PER_GPU = 1 << 20
GATE_COUNT = 1
SEED = 2026
PHILOX = (0.01, 0.03, 2.574e-5)   # fq_rel_sigma, fq_rel_clip, da_sigma (figures.jl:261,386; spin.jl:164)
PARITY_N = 10_000
PARITY_TOL = 1e-12
# flops per trajectory-step on this grid (FMA = 2), counted in the SASS [DESIGN.md §4.2] and cross-checked against ncu
# smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on (profiles/):
#   knot-table step (samples near the sweep centre -- every warp of this workload but ~0.3 %), ALGORITHMIC count:
#     1 DFMA (eps) + 2 DFMA (T) + 1 DADD (q) + 2 DMUL (T p, T q) + 8 DFMA (v += T G v) = 14 FP64 instructions = 25 flop
#   the shipped constant-memory form (mc_static_ctab_kernel) executes exactly these 14 (q_j, c0, c2 as uniform-register operands,
#   c1 through shared memory); its all-uniform variant (RBQ_MC_VARIANT=2161) factors the quadratic and executes 15 / 26
FLOPS_PER_STEP_SU2 = 25.0
FP64_INSTR_PER_STEP_SU2 = 14.0
FLOPS_EXECUTED_PER_STEP_SU2 = 25.0
# operand-delivery model of the shipped loop (tools/sass_fp64_cost.py on librbq.so; tools/proto/regbank.cu measured on B200: a
# warp-wide FP64 instruction holds the pipe 2.0 cycles with up to two register operands to fetch, 2.9 with three): per knot
# 11.75 instructions at 2.0 and 2.25 at 2.9
FP64_MODEL_CYCLES_PER_STEP_SU2 = 30.03
BYTES_PER_SAMPLE = 8 + 8 + 32 + 16   # fq, da, psi0 in; 2 fidelities out
# FP64 instructions per density-step of the Lindblad table step (lindblad_tab_kernel): 27 DFMA (Horner of the 9
# entries of E_lo + da (L1 + da (L2 + da L3))) + 3 DMUL + 6 DFMA (the small part applied) + 9 DFMA (E_hi applied) = 45
# instructions = 87 flop
LINDBLAD_INSTR_PER_STEP = 45.0
LINDBLAD_FLOPS_PER_STEP = 87.0


def sample_clocks(stop, out):
    """SM clock, max clock, power and throttle reasons every ~5 ms through NVML (nvidia-smi is too slow to start for a
    sub-second timed region); rows mimic the nvidia-smi csv layout clocks_summary() reads."""
    idx = int(os.environ.get("LOCAL_RANK", "0"))
    try:
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not stop.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            try:
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            except Exception:
                pw = 0.0
            r = reasons_fn(h)
            out.append([str(idx), str(sm), str(mx), f"{pw:.1f}", hex(r)] +
                       ["Active" if r & bits[k] else "Not Active" for k in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
            stop.wait(0.005)
        return
    except Exception:
        pass
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(idx)],
                               capture_output=True, text=True, timeout=5)
            for line in r.stdout.strip().splitlines():
                out.append([c.strip() for c in line.split(",")])
        except Exception:
            pass
        stop.wait(0.1)


def clocks_summary(rows):
    sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
    mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
    reasons = set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for r in rows:
        if len(r) >= 9:
            for name, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
    pw = [float(r[3]) for r in rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None,
            "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
            "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which must not shrink the CPU arm)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


WORKLOAD = (f"mc_static sweep (BASELINE configs[4]): P_sin({NK + 1}) = {NK} knots, dt={DT} ns, S={{S}} samples per GPU "
            f"(Philox seed {SEED} keyed by global index; fq~N(0,1%) clipped 3%, da~N(0,2.574e-5)), target Z/2, "
            f"gate_count={GATE_COUNT}, per-sample fidelities + statistics")


def cpu_arm_inputs(orc, S, first=0):
    """The bench sweep's own samples (same Philox stream as the GPU arm), drawn by the oracle's generator."""
    from rbqoc_b200 import spin

    return orc.philox_samples(first, S, SEED, spin.FQ, *PHILOX)


def cpu_reference_rate(orc, controls, n_threads, S):
    """Times the CPU arm (reference algorithm: per-knot Pade exp + 4x4 products, one sample per call) on S samples of the
    bench sweep."""
    fq, da, psi = cpu_arm_inputs(orc, S)
    t0 = time.perf_counter()
    orc.mc_static(controls, DT, 1, GATE_COUNT, fq, da, psi, nthreads=n_threads)
    t = time.perf_counter() - t0
    return S * NK / t, t


def cpu_baseline_block(orc, controls, threads, S_full, budget_s=12.0):
    """cpu_baseline of the headline workload: the full S when it fits the time budget, else a bounded slice (stated)."""
    rate0, t0 = cpu_reference_rate(orc, controls, threads, 256 * threads)          # probe (also warms the threads up)
    S = int(min(S_full, max(256 * threads, budget_s * rate0 / NK)))
    rate, t = cpu_reference_rate(orc, controls, threads, S)
    rate1, t1 = cpu_reference_rate(orc, controls, 1, max(64, int(min(S, 2.0 * rate / threads / NK))))
    return {"value": rate, "unit": "trajectory-steps/s", "cores": threads, "kind": "port",
            "build": "g++ " + " ".join(orc.NATIVE_FLAGS[:2]) + " (oracle/_native, built on this host)",
            "sample": (f"{S} of the sweep's {S_full} samples x {NK} knots in {t:.1f} s" if S < S_full else
                       f"the full workload: {S} samples x {NK} knots in {t:.1f} s") +
                      " (reference algorithm: generic Pade exp + 4x4 product per knot, OpenMP over samples)",
            "one_thread": {"value": rate1, "unit": "trajectory-steps/s", "cores": 1}}


# ---------------------------------------------------------------------------------------------------------------------------
# BASELINE configs[0..3]: each with its own value / e2e / roofline / cpu_baseline
# ---------------------------------------------------------------------------------------------------------------------------
def measured_configs(eng, rbq, spin, torch, dev, stream, flush, orc, peaks, fp64_peak):
    out = {}
    threads = host_threads()
    hbm = peaks.get("hbm_gbs") or 6542.0

    def dev_ms(fn, reps=5, warm=2):
        """median device time of fn (which only enqueues on the engine's stream), L2 flushed before every repetition"""
        for _ in range(warm):
            fn()
        evs = []
        for _ in range(reps):
            with torch.cuda.stream(stream):
                flush.zero_()
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record(stream)
            fn()
            with torch.cuda.stream(stream):
                b.record(stream)
            evs.append((a, b))
        stream.synchronize()
        return float(np.median([a.elapsed_time(b) for a, b in evs]))

    def host_ms(fn, reps=5):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts)) * 1e3

    def cpu_ms(fn, reps=3):
        fn()
        return min(host_ms(fn, reps=1) for _ in range(reps))

    dd = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    N = 1001
    K1 = N - 1
    for name, model, x0, iso in (("spin13", spin.Spin13Model(), spin.spin13_x0(), 2), ("spin12", spin.Spin12Model(), spin.spin12_x0(), 10),
                                 ("spin30", spin.Spin30Model(), spin.spin30_x0(), 13)):
        n, m = model.n, model.m
        blk = n * (n + m) * 8
        dtv = np.full(K1, 0.1)
        U = 1e-3 * np.sin(np.arange(K1) * 0.05)[:, None] * np.ones((1, m))
        X = eng.rollout(model.params, x0, U, dtv)
        # --- one trajectory (what one ALTRO iteration asks for): rollout + all Jacobians -------------------------------
        dx0, dU, ddt = dd(x0[None]), dd(U), dd(dtv)
        dX = torch.empty(N, n, dtype=torch.float64, device=dev)
        dAB = torch.empty(K1, n * (n + m), dtype=torch.float64, device=dev)
        ms_ro = dev_ms(lambda: eng.rollout_dev(model.params, 1, N, dx0, dU, ddt, dX))
        ms_jac = dev_ms(lambda: eng.jacobians_dev(model.params, K1, dX, dU, ddt, dAB))
        ab_pin = eng.pinned_empty((K1, n + m, n))
        e2e_ms = host_ms(lambda: (eng.rollout(model.params, x0, U, dtv), eng.jacobians(model.params, X[:-1], U, dtv, out=ab_pin)))
        orc.set_num_threads(threads)
        cpu_all = cpu_ms(lambda: (orc.rollout(model.params, x0, U, dtv), orc.jacobians(model.params, X[:-1], U, dtv)))
        orc.set_num_threads(1)
        cpu_one = cpu_ms(lambda: (orc.rollout(model.params, x0, U, dtv), orc.jacobians(model.params, X[:-1], U, dtv)))
        steps1 = K1 * iso
        single = {
            "workload": f"{name}: one trajectory, N={N} knots (n={n}): open-loop rollout + all {K1} knot-point Jacobians [A_k B_k]",
            "value": steps1 / ((ms_ro + ms_jac) * 1e-3), "unit": "trajectory-steps/s", "rollout_ms": ms_ro, "jacobians_ms": ms_jac,
            "e2e": {"value": steps1 / (e2e_ms * 1e-3), "unit": "trajectory-steps/s", "ms": e2e_ms,
                    "h2d_bytes": (n + K1 * m + K1) * 8 + (K1 * n + K1 * m + K1) * 8, "d2h_bytes": N * n * 8 + K1 * blk,
                    "api": "rbq_rollout + rbq_jacobians (host pointers, Jacobians into page-locked memory)"},
            "roofline": {"bound": "latency", "frac": None,
                         "how": ("one trajectory cannot fill the GPU: the floor is 2 kernel launches (~2-3 us each), the sequential control "
                                 "chain of the rollout and, through the host ABI, 2 synchronous calls with their PCIe round trips "
                                 f"(~10 us each) plus {K1 * blk / 1e6:.1f} MB of Jacobians over PCIe (~{K1 * blk / 55e9 * 1e3:.2f} ms at 55 GB/s)"),
                         "jacobian_write_gbs": K1 * blk / (ms_jac * 1e-3) * 1e-9, "hbm_peak_gbs": hbm},
            "cpu_baseline": {"value": steps1 / (cpu_all * 1e-3), "unit": "trajectory-steps/s", "cores": threads, "ms": cpu_all, "kind": "port",
                             "sample": "the same trajectory: orc.rollout + orc.jacobians (reference step maps under forward-mode duals, "
                                       "OpenMP over knots for the Jacobians)",
                             "one_thread": {"value": steps1 / (cpu_one * 1e-3), "ms": cpu_one, "cores": 1}},
        }
        # --- a batch of independent problems (hyper-parameter sweeps): the metric read literally -----------------------------
        Bt = {"spin13": 4096, "spin12": 1024, "spin30": 512}[name]
        bX0 = dd(np.tile(x0, (Bt, 1))); bU = dd(np.tile(U, (Bt, 1, 1)))
        bX = torch.empty(Bt, N, n, dtype=torch.float64, device=dev)
        # Jacobians at every state the rollout wrote, the terminal one included (N per trajectory, counted as N-1 below), so
        # that the kernel reads X in place: no repacking between the two calls
        bU2 = dd(np.tile(np.concatenate([U, U[-1:]]), (Bt, 1, 1))); bdt2 = dd(np.tile(np.concatenate([dtv, dtv[-1:]]), Bt))
        chunk = 256
        bAB = torch.empty(chunk * N, n * (n + m), dtype=torch.float64, device=dev)      # one chunk of blocks, overwritten

        def batch_rollout():
            eng.rollout_dev(model.params, Bt, N, bX0, bU, ddt, bX)

        def batch_jacobians():
            for b0 in range(0, Bt, chunk):
                cnt = min(chunk, Bt - b0)
                eng.jacobians_dev(model.params, cnt * N, bX[b0:b0 + cnt], bU2[b0:b0 + cnt], bdt2[b0 * N:(b0 + cnt) * N], bAB)

        ms_bro = dev_ms(batch_rollout, reps=3, warm=1)
        ms_bjac = dev_ms(batch_jacobians, reps=3, warm=1)
        stepsB = Bt * K1 * iso
        jac_bytes = Bt * N * blk
        # e2e of the batch through host buffers: a smaller batch (pinned output), PCIe bound by construction
        Be = {"spin13": 256, "spin12": 32, "spin30": 16}[name]
        x0e = np.tile(x0, (Be, 1)); Ue = np.tile(U, (Be, 1, 1))
        Xe = eng.rollout(model.params, x0e, Ue, dtv)
        abe = eng.pinned_empty((Be * K1, n + m, n))
        Xk, Uk, dk = Xe[:, :-1].reshape(-1, n), Ue.reshape(-1, m), np.tile(dtv, Be)
        e2e_b = host_ms(lambda: (eng.rollout(model.params, x0e, Ue, dtv), eng.jacobians(model.params, Xk, Uk, dk, out=abe)), reps=3)
        # CPU arm on a bounded slice of the batch
        Bc = {"spin13": 64, "spin12": 8, "spin30": 4}[name]
        x0c = np.tile(x0, (Bc, 1)); Uc = np.tile(U, (Bc, 1, 1))
        Xc = orc.rollout(model.params, x0c, Uc, dtv)
        Xck, Uck, dck = Xc[:, :-1].reshape(-1, n), Uc.reshape(-1, m), np.tile(dtv, Bc)
        orc.set_num_threads(threads)
        cpu_b = cpu_ms(lambda: (orc.rollout(model.params, x0c, Uc, dtv), orc.jacobians(model.params, Xck, Uck, dck)), reps=2)
        batched = {
            "workload": f"{name}: {Bt} independent problems x N={N} knots, device resident: rollout + Jacobians at every knot",
            "value": stepsB / ((ms_bro + ms_bjac) * 1e-3), "unit": "trajectory-steps/s", "rollout_ms": ms_bro, "jacobians_ms": ms_bjac,
            "e2e": {"value": Be * K1 * iso / (e2e_b * 1e-3), "unit": "trajectory-steps/s", "ms": e2e_b, "problems": Be,
                    "d2h_bytes": Be * N * n * 8 + Be * K1 * blk, "pcie_floor_ms": Be * K1 * blk / 55e9 * 1e3,
                    "api": "rbq_rollout + rbq_jacobians (host pointers): PCIe bound, the Jacobian blocks are the payload"},
            "roofline": {"bound": "hbm", "achieved": jac_bytes / (ms_bjac * 1e-3) * 1e-9, "peak": hbm, "unit": "GB/s",
                         "frac": jac_bytes / (ms_bjac * 1e-3) * 1e-9 / hbm, "traffic": None,
                         "how": f"Jacobian kernel: n(n+m) 8 B = {blk} B written per knot x {Bt * N} knots / kernel time, against the measured copy peak "
                                "(MEASURED_PEAKS.json hbm_gbs)"},
            "cpu_baseline": {"value": Bc * K1 * iso / (cpu_b * 1e-3), "unit": "trajectory-steps/s", "cores": threads, "kind": "port",
                             "sample": f"{Bc} of the {Bt} problems in {cpu_b:.0f} ms (OpenMP over trajectories / knots)"},
        }
        out[name] = {"single_trajectory": single, "batched": batched}
        del bX0, bU, bX, bAB, bU2, bdt2, ab_pin, abe
    # --- configs[3]: Lindblad T1 rollout over 1e5 noise realisations ---------------------------------------------------------
    S = 100_000
    rng = np.random.default_rng(0)
    psi = rng.random((S, 2)) + 1j * rng.random((S, 2))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    rho0 = np.einsum("si,sj->sij", psi, psi.conj())
    vec = np.ascontiguousarray(rho0.transpose(0, 2, 1).reshape(S, 4)).view(np.float64)
    controls = spin.pulse_sin(N)[:-1]
    da_h = 2.574e-5 * rng.standard_normal(S)
    dcn, dda, drho = (torch.from_numpy(a).to(dev) for a in (controls, da_h, vec))
    dfid = torch.empty(S, 2, dtype=torch.float64, device=dev)
    dro = torch.empty(S, 8, dtype=torch.float64, device=dev)
    dst = torch.zeros(rbq.STATS_INT64_WORDS, dtype=torch.int64, device=dev)
    ms_l = dev_ms(lambda: eng.lindblad_dev(dcn, K1, 0.1, rbq.GATE_XPIBY2, 1, S, spin.FQ, dda, drho, rbq.LINDBLAD_T1, 0.0, 0.0, dfid, dro, dst))
    # e2e: the C ABI with host buffers in its own layout, page-locked (7.2 MB up, 8 MB down inside the timed call); the pageable
    # numpy form with the wrapper's transposes is reported beside it
    p_da, p_vec = eng.pinned_empty(S), eng.pinned_empty((S, 8))
    p_fid, p_ro = eng.pinned_empty((S, 2)), eng.pinned_empty((S, 8))
    p_da[:] = da_h; p_vec[:] = vec
    e2e_l = host_ms(lambda: eng.lindblad_into(controls, 0.1, rbq.GATE_XPIBY2, 1, spin.FQ, p_da, p_vec, p_fid, p_ro), reps=5)
    e2e_l_pageable = host_ms(lambda: eng.lindblad(controls, 0.1, rbq.GATE_XPIBY2, 1, spin.FQ, da_h, rho0), reps=3)
    fid_chk, _, _ = eng.lindblad(controls, 0.1, rbq.GATE_XPIBY2, 1, spin.FQ, da_h[:64], rho0[:64])
    if not np.array_equal(fid_chk, p_fid[:64]):
        raise SystemExit("lindblad: the page-locked ABI call and the numpy wrapper disagree")
    Sc = 2000
    orc.set_num_threads(threads)
    cpu_l = cpu_ms(lambda: orc.lindblad_t1(controls, 0.1, rbq.GATE_XPIBY2, 1, spin.FQ, da_h[:Sc], rho0[:Sc], nthreads=threads), reps=2)
    cpu_l1 = cpu_ms(lambda: orc.lindblad_t1(controls, 0.1, rbq.GATE_XPIBY2, 1, spin.FQ, da_h[:128], rho0[:128], nthreads=1), reps=2)
    ach = S * K1 * LINDBLAD_FLOPS_PER_STEP / (ms_l * 1e-3) * 1e-12
    out["lindblad_t1"] = {
        "workload": f"Lindblad T1 rollout (BASELINE configs[3]): P_sin({N}), {S} realisations (random pure rho0, static da~N(0,2.574e-5)), "
                    "1 gate, sqrt-fidelity + final density per realisation",
        "value": S * K1 / (ms_l * 1e-3), "unit": "density-steps/s", "ms": ms_l,
        "e2e": {"value": S * K1 / (e2e_l * 1e-3), "unit": "density-steps/s", "ms": e2e_l, "h2d_bytes": S * 72 + K1 * 8, "d2h_bytes": S * 80 + 2096,
                "api": "rbq_lindblad (host pointers, page-locked buffers in the ABI's layout)",
                "pageable_numpy_wrapper_ms": e2e_l_pageable},
        "roofline": {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak if fp64_peak else None,
                     "fp64_instr_per_density_step": LINDBLAD_INSTR_PER_STEP, "flops_per_density_step": LINDBLAD_FLOPS_PER_STEP,
                     "fp64_pipe_issue_util": S * K1 * LINDBLAD_INSTR_PER_STEP / (ms_l * 1e-3) * 2e-12 / fp64_peak if fp64_peak else None,
                     "how": "table step: 45 FP64 instructions (42 DFMA + 3 DMUL = 87 flop) per density-step counted in the SASS of lindblad_tab_kernel (DESIGN.md §4.4); knots next to a kink of the T1 table take the longer series step"},
        "cpu_baseline": {"value": Sc * K1 / (cpu_l * 1e-3), "unit": "density-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{Sc} of the {S} realisations in {cpu_l:.0f} ms (integrate_prop_lindbladt1!: complex 4x4 Pade exp + product per knot)",
                         "one_thread": {"value": 128 * K1 / (cpu_l1 * 1e-3), "cores": 1}},
    }
    return out


def run_reference(args):
    """--impl reference: the reference's CPU path.  The reference is Julia (absent here and on the GPU box), so this is the
    C++ restatement of its algorithm built -O3 -march=native on this host, all host threads, the FULL per-step workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    from rbqoc_b200 import spin

    orc.build()
    orc.use_native_build()
    controls = spin.pulse_sin(NK + 1)[:-1]
    threads = host_threads()
    S = args.samples
    # a probe decides whether the full S fits ~6 s per step; otherwise the step is a bounded slice (stated in `sample`)
    rate0, _ = cpu_reference_rate(orc, controls, threads, 256 * threads)
    S_step = int(min(S, max(256 * threads, 6.0 * rate0 / NK)))
    fq, da, psi = cpu_arm_inputs(orc, S_step)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        orc.mc_static(controls, DT, 1, GATE_COUNT, fq, da, psi, nthreads=threads)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = float(np.mean(times)) * 1e3
    value = S_step * NK / (ms * 1e-3)
    sample = (f"the full workload per step: {S_step} samples x {NK} knots" if S_step == S else
              f"bounded slice per step: {S_step} of {S} samples x {NK} knots (throughput normalised)")
    line = {
        "impl": "reference", "metric": "fp64_trajectory_steps_per_sec", "value": value, "unit": "trajectory-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(S=S), "samples_per_gpu": S, "knots": NK, "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": "trajectory-steps/s", "cores": threads, "kind": "port",
                         "build": "g++ " + " ".join(orc.NATIVE_FLAGS[:2]) + " (oracle/_native, built on this host)",
                         "sample": sample + ", OpenMP over samples; reference algorithm = generic Pade exp + 4x4 product per knot "
                                            "(the Julia reference itself cannot run: no julia on this box)"},
        "e2e": {"value": value, "unit": "trajectory-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=S_PER_GPU, help="samples per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs[0..3] and the strong-scaling table")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch

    import rbqoc_b200 as rbq
    from rbqoc_b200 import dist as rdist
    from rbqoc_b200 import spin

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cpus = None
    if world > 1:
        import torch.distributed as tdist

        # one process per GPU: keep the process (and the pinned buffers it first-touches) on the GPU's own NUMA node
        if os.environ.get("RBQ_NO_NUMA_BIND", "0") in ("", "0"):
            numa_cpus = rdist.bind_to_gpu_numa_node(local)

        # NCCL prints its version banner to stdout at levels VERSION and WARN; the contract is ONE JSON line on stdout
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "NONE"
        tdist.init_process_group("nccl", device_id=dev)
    S = args.samples
    first = rank * S                      # weak scaling: every rank owns S samples of the global index range
    eng = rbq.Engine(local)
    stream = torch.cuda.Stream(device=dev)
    eng.set_stream(stream.cuda_stream)

    controls_h = spin.pulse_sin(NK + 1)[:-1]
    with torch.cuda.stream(stream):
        d_controls = torch.from_numpy(controls_h).to(dev)
        d_fq = torch.empty(S, dtype=torch.float64, device=dev)
        d_da = torch.empty(S, dtype=torch.float64, device=dev)
        d_psi0 = torch.empty(S, 4, dtype=torch.float64, device=dev)
        d_fid = torch.empty(S, GATE_COUNT + 1, dtype=torch.float64, device=dev)
        d_stats = torch.zeros(rbq.STATS_INT64_WORDS, dtype=torch.int64, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    # the sample set: Philox keyed by global index (materialised once; resident in HBM for `value`)
    fq_h, da_h, psi0_h = eng.philox_samples(first, S, SEED, spin.FQ, *PHILOX)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    fq_p, da_p, psi0_p = pin(fq_h), pin(da_h), pin(psi0_h)
    fid_p = torch.empty(S, GATE_COUNT + 1, dtype=torch.float64).pin_memory()
    fid_p2 = torch.empty(S, GATE_COUNT + 1, dtype=torch.float64).pin_memory()
    with torch.cuda.stream(stream):
        d_fq.copy_(fq_p, non_blocking=True); d_da.copy_(da_p, non_blocking=True); d_psi0.copy_(psi0_p, non_blocking=True)
    stream.synchronize()

    # N > 1: the all-gather of step i's 2.1 KB statistics, its copy to pinned host memory and the host merge run on a side
    # stream / the host while step i+1's kernel computes (two rotating statistics buffers); every merge is complete
    # before the timed region closes, and the tail after the last kernel is added to the device time.
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    W = rbq.STATS_INT64_WORDS
    if world > 1:
        with torch.cuda.stream(stream):
            d_stats2 = [torch.zeros(W, dtype=torch.int64, device=dev) for _ in range(2)]
            gathered = [torch.zeros(world * W, dtype=torch.int64, device=dev) for _ in range(2)]
        pinned_g = [torch.zeros(world * W, dtype=torch.int64).pin_memory() for _ in range(2)]
        done_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        pending = [False, False]
    merged = {"last": None, "count": 0}

    def collect(b):
        """Host side of a finished exchange: wait for its pinned copy, merge the per-rank blocks in rank order."""
        if not pending[b]:
            return
        done_ev[b].synchronize()
        host = pinned_g[b].numpy().reshape(world, W)
        total = rbq.stats_from_words(host[0])
        for r in range(1, world):
            rbq.stats_merge(total, rbq.stats_from_words(host[r]))
        merged["last"] = total
        merged["count"] += 1
        pending[b] = False

    def device_step(i=0):
        if world == 1:
            eng.stats_reset_dev(d_stats)
            eng.mc_static_dev(d_controls, NK, DT, rbq.GATE_ZPIBY2, GATE_COUNT, S, d_fq, d_da, d_psi0, rbq.ALGO_SU2, d_fid, d_stats)
            return
        b = i & 1
        collect(b)                                   # the exchange that last used this buffer (two steps ago)
        eng.stats_reset_dev(d_stats2[b])
        eng.mc_static_dev(d_controls, NK, DT, rbq.GATE_ZPIBY2, GATE_COUNT, S, d_fq, d_da, d_psi0, rbq.ALGO_SU2, d_fid, d_stats2[b])
        ready = torch.cuda.Event()
        ready.record(stream)
        with torch.cuda.stream(comm):
            comm.wait_event(ready)
            tdist.all_gather_into_tensor(gathered[b], d_stats2[b])        # the one collective of the step
            pinned_g[b].copy_(gathered[b], non_blocking=True)
            done_ev[b].record(comm)
        pending[b] = True

    def drain():
        if world > 1:
            collect(0)
            collect(1)

    def barrier():
        stream.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident, CUDA events on the launching stream, L2 flushed between steps ----
    stop, rows = threading.Event(), []
    th = threading.Thread(target=sample_clocks, args=(stop, rows), daemon=True)
    th.start()
    for wi in range(args.warmup):
        with torch.cuda.stream(stream):
            flush.zero_()
        device_step(wi)
    drain()
    barrier()
    launches0 = eng.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    merged["count"] = 0
    for i, (a, b) in enumerate(ev):
        with torch.cuda.stream(stream):
            flush.zero_()
            a.record(stream)
        device_step(i)
        with torch.cuda.stream(stream):
            b.record(stream)
    last_b = (args.steps - 1) & 1
    drain()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = eng.launch_count - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    if world > 1:
        assert merged["count"] == args.steps, "every step's statistics must have been exchanged and merged"
        dev_ms += max(0.0, ev[-1][1].elapsed_time(done_ev[last_b]))   # tail: last exchange after the last kernel
    # kernel-only time of the same step (no stats reset / collective): the roofline numerator's clock
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in kev:
        with torch.cuda.stream(stream):
            flush.zero_()
            a.record(stream)
        eng.mc_static_dev(d_controls, NK, DT, rbq.GATE_ZPIBY2, GATE_COUNT, S, d_fq, d_da, d_psi0, rbq.ALGO_SU2, d_fid, d_stats)
        with torch.cuda.stream(stream):
            b.record(stream)
    barrier()
    kern_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    eng.stats_reset_dev(d_stats)      # leave the statistics of exactly one pass in d_stats
    eng.mc_static_dev(d_controls, NK, DT, rbq.GATE_ZPIBY2, GATE_COUNT, S, d_fq, d_da, d_psi0, rbq.ALGO_SU2, d_fid, d_stats)

    # ---- e2e (headline): the seed-in call -- host controls + (seed, first, S) in, per-sample fidelities + statistics out in
    #      page-locked HOST memory (rbq_mc_static_philox); the reference's API takes seeds, not sample arrays ----
    fid_np, fid_np2 = fid_p.numpy(), fid_p2.numpy()
    e2e_steps = max(3, args.steps // 2)

    def timed_host_calls(call):
        for _ in range(max(1, args.warmup // 2)):
            call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            res = call()
        barrier()
        return (time.perf_counter() - t0) / e2e_steps, res

    seed_call = lambda: eng.mc_static_philox(controls_h, DT, rbq.GATE_ZPIBY2, GATE_COUNT, first, S, SEED, spin.FQ, *PHILOX, fid_out=fid_np2)
    e2e_seed_s, (_, st_seed) = timed_host_calls(seed_call)
    # ---- e2e, second form: explicit host sample arrays (48 B per sample up), zero-copy kernel over PCIe ----
    arr_call = lambda: eng.mc_static(controls_h, DT, rbq.GATE_ZPIBY2, GATE_COUNT, fq_p.numpy(), da_p.numpy(), psi0_p.numpy(), fid_out=fid_np)
    e2e_arr_s, (_, st_e2e) = timed_host_calls(arr_call)
    stop.set()
    th.join(timeout=2)

    # raw PCIe rates from the same pinned buffers: the ceiling of the explicit-sample e2e path (48 B up, 16 B down per sample)
    def copy_rate(dst, src, reps=5):
        with torch.cuda.stream(stream):
            dst.copy_(src, non_blocking=True)
        stream.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(stream):
                dst.copy_(src, non_blocking=True)
        stream.synchronize()
        return src.numel() * src.element_size() * reps / (time.perf_counter() - t0) * 1e-9

    h2d_gbs = copy_rate(d_psi0, psi0_p)
    d2h_gbs = copy_rate(fid_p, d_fid)

    # ---- strong scaling: S_total samples split over the ranks of this run (seed-in form, statistics only) ----
    strong = {}
    if not args.no_configs:
        d_st_s = torch.zeros(W, dtype=torch.int64, device=dev)
        g_s = torch.zeros(world * W, dtype=torch.int64, device=dev)
        for S_total in (10 ** 6, 10 ** 7, 10 ** 8):
            f0, cnt = rdist.shard_range(S_total, rank, world)

            def strong_step():
                eng.stats_reset_dev(d_st_s)
                eng.mc_static_philox_dev(d_controls, NK, DT, rbq.GATE_ZPIBY2, GATE_COUNT, f0, cnt, SEED, spin.FQ, *PHILOX, rbq.ALGO_SU2, None, d_st_s)
                if world > 1:
                    with torch.cuda.stream(stream):
                        tdist.all_gather_into_tensor(g_s, d_st_s)

            strong_step()
            barrier()
            reps = 3
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                a.record(stream)
            for _ in range(reps):
                strong_step()
            with torch.cuda.stream(stream):
                b.record(stream)
            barrier()
            t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
            if world > 1:
                tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
            ms = float(t.cpu())
            host = (g_s if world > 1 else d_st_s).cpu().numpy().reshape(world, W)
            total = rbq.stats_from_words(host[0])
            for r in range(1, world):
                rbq.stats_merge(total, rbq.stats_from_words(host[r]))
            strong[f"{S_total:.0e}"] = {"samples_total": S_total, "ms": ms, "value": S_total * NK / (ms * 1e-3), "merged_count": int(total.count),
                                        "mean_infidelity": total.mean}

    # ---- reduce over ranks (max time), print on rank 0 ----
    times = torch.tensor([dev_ms, e2e_seed_s * 1e3, e2e_arr_s * 1e3, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(times, op=tdist.ReduceOp.MAX)
    dev_ms, e2e_seed_ms, e2e_arr_ms, kern_ms = (float(x) for x in times.cpu())
    stream.synchronize()
    st = rbq.stats_from_words(d_stats.cpu().numpy())     # this rank's last single pass (kernel-only loop epilogue)
    if rank == 0:
        steps_per_step_all = float(world) * S * NK
        value = steps_per_step_all * args.steps / (dev_ms * 1e-3)
        peak_tflops = eng.fp64_peak(20000)
        ach_tflops = S * NK * FLOPS_PER_STEP_SU2 / (kern_ms * 1e-3) * 1e-12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        clk = clocks_summary(rows)
        # ---- parity of this very sweep: first PARITY_N samples against the CPU oracle (parity build, not the timing build) ----
        from oracle import oracle as orc

        orc.build()
        npar = min(PARITY_N, S)
        ref = orc.mc_static(controls_h, DT, rbq.GATE_ZPIBY2, GATE_COUNT, fq_h[:npar], da_h[:npar], psi0_h[:npar])
        got = d_fid[:npar].cpu().numpy()
        par_abs = float(np.max(np.abs(got - ref)))
        # the seed-in call draws the same samples on the device (Philox keyed by global index): same numbers, bit for bit
        seed_vs_arrays = float(np.max(np.abs(fid_np2 - d_fid.cpu().numpy())))
        ofq, oda, opsi = orc.philox_samples(first, npar, SEED, spin.FQ, *PHILOX)
        samples_abs = float(max(np.max(np.abs(ofq - fq_h[:npar])), np.max(np.abs(oda - da_h[:npar])), np.max(np.abs(opsi - psi0_h[:npar]))))
        parity = {"n": npar, "max_abs": par_abs, "tol": PARITY_TOL, "against": "orc.mc_static (reference algorithm, parity build) on the sweep's own first samples",
                  "seed_in_vs_sample_arrays_max_abs": seed_vs_arrays, "device_vs_oracle_philox_samples_max_abs": samples_abs}
        line = {
            "metric": "fp64_trajectory_steps_per_sec", "value": value, "unit": "trajectory-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(S=S), "samples_per_gpu": S, "knots": NK,
                       "algo": "su2 closed form, tan form, knot-table step (local quadratic of tan(theta)/theta around the sweep centre), "
                               "table rows as uniform-register operands from constant memory (c1 through shared memory)",
                       "l2": "flushed between steps (256 MiB memset, untimed)",
                       "sharding": ("contiguous sample ranges per rank; per step ONE all-gather of the 2.1 KB statistics, copied to pinned host "
                                    "memory and merged on the host while the next step computes") if world > 1 else "single GPU"},
            "wall_s_timed_region": t_wall,
            "parity": parity,
            "roofline": {"bound": "fp64", "achieved": ach_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": ach_tflops / peak_tflops if peak_tflops else None,
                         "traffic": None,
                         "traffic_note": "per-launch dram read+write of this kernel: see the `ncu --set full` capture of this round under profiles/ "
                                         "(r1: 50.4 MB read + 1.8 MB written against 64 MB algorithmic; the 16 MB of fidelities stay in L2)",
                         "fp64_pipe_issue_util": S * NK * FP64_INSTR_PER_STEP_SU2 / (kern_ms * 1e-3) * 2e-12 / peak_tflops if peak_tflops else None,
                         "peak_source": "measured live: rbq_fp64_peak DFMA saturation loop (MEASURED_PEAKS.json has no FP64 entry)",
                         # register-operand model of the FP64 pipe: a warp-wide instruction holds the pipe 2.0 cycles with up to two
                         # register operands to fetch, 2.9 with three (operands kept by the reuse cache, uniform registers and
                         # immediates are free)
                         "issue_model": {"cycles_per_warp_step_floor": FP64_MODEL_CYCLES_PER_STEP_SU2,
                                         "cycles_per_warp_step_all_operands_reused": 2.0 * FP64_INSTR_PER_STEP_SU2,
                                         "cycles_per_warp_step_measured": (clk["sm_mhz"] or 1965.0) * 1e6 * 4 * 148 * (kern_ms * 1e-3) / (S * NK / 32.0),
                                         "note": "floor = what the SASS of the shipped loop costs under the measured operand-delivery model (tools/sass_fp64_cost.py); floor / measured = how close the kernel runs to it"},
                         "flops_per_trajectory_step": FLOPS_PER_STEP_SU2, "fp64_instr_per_trajectory_step": FP64_INSTR_PER_STEP_SU2,
                         "flops_executed_per_trajectory_step": FLOPS_EXECUTED_PER_STEP_SU2,
                         "frac_executed_flops": (S * NK * FLOPS_EXECUTED_PER_STEP_SU2 / (kern_ms * 1e-3) * 1e-12 / peak_tflops) if peak_tflops else None,
                         "kernel_ms": kern_ms,
                         "hbm": {"achieved_gbs": S * BYTES_PER_SAMPLE / (kern_ms * 1e-3) * 1e-9, "peak_gbs": peaks.get("hbm_gbs"),
                                 "note": "algorithmic bytes; shows the path is not HBM bound"}},
            "e2e": {"value": steps_per_step_all / (e2e_seed_ms * 1e-3), "unit": "trajectory-steps/s",
                    "h2d_bytes_per_step": NK * 8, "d2h_bytes_per_step": S * 16 + 2096, "ms_per_step": e2e_seed_ms,
                    "numa_bound_cpus": len(numa_cpus) if numa_cpus else None,
                    "api": "rbq_mc_static_philox: host controls + (seed, first, S) in; S x 2 fidelities + rbq_stats out in page-locked host memory "
                           "(written by the kernel over PCIe while it computes); the shape of the reference's run_sim_prop(...; seed, state_seed)",
                    "explicit_sample_arrays": {"value": steps_per_step_all / (e2e_arr_ms * 1e-3), "ms_per_step": e2e_arr_ms,
                                               "h2d_bytes_per_step": S * 48 + NK * 8, "d2h_bytes_per_step": S * 16 + 2096,
                                               "api": "rbq_mc_static (host sample arrays, pinned): zero-copy, the kernel reads the samples and writes the "
                                                      "fidelities over PCIe"},
                    "pcie_measured": {"h2d_gbs": h2d_gbs, "d2h_gbs": d2h_gbs,
                                      "h2d_floor_ms_explicit": (S * 48 + NK * 8) / (h2d_gbs * 1e6) if h2d_gbs else None,
                                      "d2h_floor_ms": (S * 16) / (d2h_gbs * 1e6) if d2h_gbs else None}},
            "gpu_launches": launches,
            "clocks": clk,
            "result": {"mean_infidelity": st.mean, "count": st.count, "mean_infidelity_seed_in_call": st_seed.mean,
                       "merged_count_all_ranks": merged["last"].count if merged["last"] is not None else None},
        }
        if strong:
            line["strong_scaling"] = dict(strong, note="S_total samples split over this run's ranks (contiguous global index ranges), "
                                                       "rbq_mc_static_philox_dev + one all-gather of the statistics per pass; device time, max over ranks")
        native_ok = False
        if not args.no_cpu_baseline or (world == 1 and not args.no_configs):
            try:
                orc.use_native_build()      # timing build of the CPU arm from here on (parity above used the parity build)
                native_ok = True
            except Exception as exc:
                line["cpu_native_build_error"] = repr(exc)
        if world == 1 and not args.no_configs:
            try:
                line["configs"] = measured_configs(eng, rbq, spin, torch, dev, stream, flush, orc, peaks, peak_tflops)
            except Exception as exc:   # never let the extra configurations break the contract line
                line["configs"] = {"error": repr(exc)}
        if not args.no_cpu_baseline:
            threads = host_threads()
            cb = cpu_baseline_block(orc, controls_h, threads, S)
            if not native_ok:
                cb["build"] = "liboracle.so (-O2 -mavx2 -ffp-contract=off; the native build failed)"
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
        if par_abs > PARITY_TOL or seed_vs_arrays > PARITY_TOL:
            print(f"PARITY FAILURE: max_abs={par_abs:.3e} (tol {PARITY_TOL}), seed-in vs arrays {seed_vs_arrays:.3e}", file=sys.stderr, flush=True)
            if world > 1:
                tdist.barrier()
                tdist.destroy_process_group()
            sys.exit(3)
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
