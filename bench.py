#!/usr/bin/env python
"""bench.py — FP64 trajectory-steps/s of the Monte-Carlo robustness rollout (BASELINE.json config 5,
the configuration the north-star target is quoted on: "the 1M-sample robustness rollout").

One step = one pass of the hot path over one batch: a fixed pulse P_sin(1001) (1000 control knots,
dt = 0.1 ns) rolled over S = 2^20 (detuning, amplitude-error, initial-state) samples PER GPU, one gate,
fidelity per sample + infidelity statistics.  A trajectory-step is one 4-real iso state advanced by
one knot, so a step is S * 1000 trajectory-steps per GPU.

  value      device-resident inputs, rbq_mc_static_dev + (N>1) the one all-gather of rbq_stats
  e2e        the host-pointer C-ABI call rbq_mc_static with pinned host buffers (H2D + D2H inside)
  roofline   FP64 pipe: flops the SU(2) kernel executes per trajectory-step / measured DFMA peak
  cpu_baseline / --impl reference: the CPU oracle (port of the reference algorithm; the Julia
             reference cannot run in this image) on all host threads, bounded sample
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
# flops the closed-form kernel executes per trajectory-step on this grid (FMA = 2), counted in the SASS [DESIGN.md §4.2]:
#   knot-table step (samples near the sweep centre -- every warp of this workload but ~0.3 %):
#     1 DFMA (eps) + 2 DFMA (T) + 1 DADD (q) + 2 DMUL (T p, T q) + 8 DFMA (v += T G v) = 14 FP64 instructions = 25 flop
#   polynomial step (RBQ_NO_TABLE_CLASS=1, or samples far from the centre; tan form, degree-6 polynomial, class B):
#     1 DADD (q) + 1 DFMA (x) + 6 DFMA (T(x)) + 2 DMUL + 8 DFMA = 18 FP64 instructions = 33 flop
TABLE_CLASS = os.environ.get("RBQ_NO_TABLE_CLASS", "0") in ("", "0")
FLOPS_PER_STEP_SU2 = 25.0 if TABLE_CLASS else 33.0
FP64_INSTR_PER_STEP_SU2 = 14.0 if TABLE_CLASS else 18.0
# FP64 instructions per step with three 64-bit register operands (state updates; with the table also eps and the two
# Horner steps, whose coefficients are per-knot registers instead of constants)
FP64_INSTR3_PER_STEP_SU2 = 11.0 if TABLE_CLASS else 8.0
BYTES_PER_SAMPLE = 8 + 8 + 32 + 16   # fq, da, psi0 in; 2 fidelities out
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel at the default sizes, from the
# `ncu --set full` capture summarised in profiles/r1b_ncu_mc_static_full.csv (50.44 MB read + 1.79 MB written: the 16 MB of
# fidelities stay in the 126 MB L2); algorithmic bytes are 48 MB in + 16 MB out
NCU_DRAM_TRAFFIC_BYTES_DEFAULT = 50_435_072 + 1_789_952


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


def cpu_reference_rate(orc, controls, n_threads, target_seconds=12.0):
    """Times the oracle (reference algorithm: per-knot Pade exp + 4x4 products, one sample per call)
    on all host threads over a bounded sample of the same workload."""
    rng = np.random.default_rng(0)

    def inputs(S):
        fq = 1.4e-2 * (1 + np.clip(0.01 * rng.standard_normal(S), -0.03, 0.03))
        da = 2.574e-5 * rng.standard_normal(S)
        psi = rng.random((S, 4))
        psi /= np.linalg.norm(psi, axis=1, keepdims=True)
        return fq, da, psi

    S0 = 256 * n_threads
    fq, da, psi = inputs(S0)
    t0 = time.perf_counter()
    orc.mc_static(controls, DT, 1, GATE_COUNT, fq, da, psi, nthreads=n_threads)
    t_probe = time.perf_counter() - t0
    S = int(max(S0, min(S0 * target_seconds / max(t_probe, 1e-6), 4_000_000)))
    fq, da, psi = inputs(S)
    t0 = time.perf_counter()
    orc.mc_static(controls, DT, 1, GATE_COUNT, fq, da, psi, nthreads=n_threads)
    t = time.perf_counter() - t0
    return S * NK / t, S, t


def other_configs(eng, rbq, spin, torch, dev):
    """BASELINE configs[0..3] at their named sizes, for context next to the headline (not bench lines of their own):
    what one ALTRO iteration's worth of hot-path calls costs through the host-pointer C ABI, and the Lindblad sweep."""
    out = {}

    def timed_host(fn, reps=5):
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps * 1e3

    N = 1001
    for name, model, x0 in (("spin13", spin.Spin13Model(), spin.spin13_x0()), ("spin12", spin.Spin12Model(), spin.spin12_x0()),
                            ("spin30", spin.Spin30Model(), spin.spin30_x0())):
        n, m = model.n, model.m
        dtv = np.full(N - 1, 0.1)
        U = 1e-3 * np.sin(np.arange(N - 1) * 0.05)[:, None] * np.ones((1, m))
        X = eng.rollout(model.params, x0, U, dtv)
        K = 1e-3 * np.ones((N - 1, m, n))
        d = 1e-4 * np.ones((N - 1, m))
        alphas = 0.5 ** np.arange(8)
        out[name] = {
            "knots": N, "n": n,
            "rollout_ms": timed_host(lambda: eng.rollout(model.params, x0, U, dtv)),
            "jacobians_ms": timed_host(lambda: eng.jacobians(model.params, X[:-1], U, dtv)),
            "jacobians_pinned_out_ms": (lambda buf: timed_host(lambda: eng.jacobians(model.params, X[:-1], U, dtv, out=buf)))(
                eng.pinned_empty((N - 1, n + m, n))),
            "closedloop_8_candidates_ms": timed_host(lambda: eng.rollout_closedloop(model.params, X, U, K, d, dtv, alphas)),
            "api": "rbq_rollout / rbq_jacobians / rbq_rollout_closedloop, pageable host buffers (pinned where named), copies included",
        }
        # the backward pass on the Jacobians where rbq_jacobians_dev leaves them (device resident; LQR blocks of spin13.jl:128-135)
        dd = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        dX, dU, ddt = dd(X[:-1]), dd(U), dd(dtv)
        dAB = torch.empty(N - 1, n * (n + m), dtype=torch.float64, device=dev)
        dExx = dd(np.tile(np.eye(n), (N, 1, 1))); dEx = dd(0.1 * X)
        dEuu = dd(np.tile(0.1 * np.eye(m), (N - 1, 1, 1))); dEux = dd(np.zeros((N - 1, n, m))); dEu = dd(0.1 * U)
        dK = torch.empty(N - 1, n * m, dtype=torch.float64, device=dev); dd_ = torch.empty(N - 1, m, dtype=torch.float64, device=dev)
        ddV = torch.empty(2, dtype=torch.float64, device=dev); dfl = torch.empty(1, dtype=torch.int64, device=dev)

        def expansion_and_pass():
            eng.jacobians_dev(model.params, N - 1, dX, dU, ddt, dAB)
            eng.backward_pass_dev(n, m, N, 1, dAB, dExx, dEx, dEuu, dEux, dEu, 1e-3, rbq.REG_CONTROL, dK, dd_, ddV, dfl)
            eng.synchronize()

        out[name]["jacobians_plus_backward_pass_dev_ms"] = timed_host(expansion_and_pass)
        # ... and the whole hot path of one iLQR iteration on the device: cost expansion (below), dynamics expansion,
        # backward pass, 8 line-search candidates and their merit values
        dXfull, dal = dd(X), dd(alphas)
        dXc = torch.empty(8, N, n, dtype=torch.float64, device=dev); dUc = torch.empty(8, N - 1, m, dtype=torch.float64, device=dev)

        # the cost side of the iteration: LQR objective + the reference's constraint list as augmented-Lagrangian terms
        table = {"spin13": spin.spin13_constraints, "spin12": spin.spin12_constraints, "spin30": spin.spin30_constraints}[name](N, X[-1])
        table.upload(dev)
        dlam = torch.zeros(table.dual_size, dtype=torch.float64, device=dev); dmu = torch.ones(table.dual_size, dtype=torch.float64, device=dev)
        dc = torch.empty(table.dual_size, dtype=torch.float64, device=dev); dJ = torch.empty(N, dtype=torch.float64, device=dev)
        dcc = torch.empty(8, table.dual_size, dtype=torch.float64, device=dev); dJc = torch.empty(8, N, dtype=torch.float64, device=dev)
        dQ = dd(np.ones(n)); dR = dd(0.1 * np.ones(m)); dQf = dd(float(N) * np.ones(n)); dxf = dd(X[-1])
        dExxa = torch.empty(N, n * n, dtype=torch.float64, device=dev); dExa = torch.empty(N, n, dtype=torch.float64, device=dev)
        dEuua = torch.empty(N - 1, m * m, dtype=torch.float64, device=dev); dEuxa = torch.empty(N - 1, m * n, dtype=torch.float64, device=dev)
        dEua = torch.empty(N - 1, m, dtype=torch.float64, device=dev)

        def al_cost():
            eng.al_cost_expansion_dev(n, m, N, dXfull, dU, dQ, dR, dQf, dxf, None, table, dlam, dmu, dExxa, dExa, dEuua, dEuxa, dEua, dc, dJ)

        out[name]["al_cost_expansion_dev_ms"] = timed_host(lambda: (al_cost(), eng.synchronize()))

        def iteration_dev():
            al_cost()
            eng.jacobians_dev(model.params, N - 1, dXfull, dU, ddt, dAB)
            eng.backward_pass_dev(n, m, N, 1, dAB, dExxa, dExa, dEuua, dEuxa, dEua, 1e-3, rbq.REG_CONTROL, dK, dd_, ddV, dfl)
            eng.rollout_closedloop_dev(model.params, N, dXfull, dU, dK, dd_, ddt, 8, dal, dXc, dUc)
            eng.al_cost_expansion_dev(n, m, N, dXc, dUc, dQ, dR, dQf, dxf, None, table, dlam, dmu, None, None, None, None, None, dcc, dJc, B=8)
            eng.synchronize()

        out[name]["ilqr_iteration_dev_ms"] = timed_host(iteration_dev)
        # BASELINE's metric read literally (rollout + Jacobians): a batch of independent problems, device resident.  One
        # trajectory-step = one 4-real iso state vector advanced one knot (spin13 carries 2, spin12 10, spin30 13 per knot).
        Bt = {"spin13": 4096, "spin12": 1024, "spin30": 512}[name]
        iso_vectors = {"spin13": 2, "spin12": 10, "spin30": 13}[name]
        bX0 = dd(np.tile(x0, (Bt, 1))); bU = dd(np.tile(U, (Bt, 1, 1)))
        bX = torch.empty(Bt, N, n, dtype=torch.float64, device=dev)
        # Jacobians at every state the rollout wrote, the terminal one included (N per trajectory, counted as N-1 below), so
        # that the kernel reads X in place: no repacking between the two calls
        bU2 = dd(np.tile(np.concatenate([U, U[-1:]]), (Bt, 1, 1))); bdt2 = dd(np.tile(np.concatenate([dtv, dtv[-1:]]), Bt))
        chunk = 256
        bAB = torch.empty(chunk * N, n * (n + m), dtype=torch.float64, device=dev)      # one chunk of blocks, overwritten

        def batch_rollout_jacobians():
            eng.rollout_dev(model.params, Bt, N, bX0, bU, ddt, bX)
            for b0 in range(0, Bt, chunk):
                cnt = min(chunk, Bt - b0)
                eng.jacobians_dev(model.params, cnt * N, bX[b0:b0 + cnt], bU2[b0:b0 + cnt], bdt2[b0 * N:(b0 + cnt) * N], bAB)
            eng.synchronize()

        ms_b = timed_host(batch_rollout_jacobians, reps=2)
        out[name]["batch_rollout_plus_jacobians"] = {"problems": Bt, "ms": ms_b, "iso_vectors_per_knot": iso_vectors,
                                                     "trajectory_steps_per_s": Bt * (N - 1) * iso_vectors / ms_b * 1e3,
                                                     "jacobian_bytes_written": Bt * N * n * (n + m) * 8}
        del bX0, bU, bX, bAB, bU2, bdt2
    S = 100_000
    rng = np.random.default_rng(0)
    psi = rng.random((S, 2)) + 1j * rng.random((S, 2))
    psi /= np.linalg.norm(psi, axis=1, keepdims=True)
    rho0 = np.einsum("si,sj->sij", psi, psi.conj())
    vec = np.ascontiguousarray(rho0.transpose(0, 2, 1).reshape(S, 4)).view(np.float64)
    controls = spin.pulse_sin(N)[:-1]
    dcn, dda, drho = (torch.from_numpy(a).to(dev) for a in (controls, 2.574e-5 * rng.standard_normal(S), vec))
    dfid = torch.empty(S, 2, dtype=torch.float64, device=dev)
    dro = torch.empty(S, 8, dtype=torch.float64, device=dev)
    dst = torch.zeros(rbq.STATS_INT64_WORDS, dtype=torch.int64, device=dev)

    def lind():
        eng.lindblad_dev(dcn, N - 1, 0.1, rbq.GATE_XPIBY2, 1, S, spin.FQ, dda, drho, rbq.LINDBLAD_T1, 0.0, 0.0, dfid, dro, dst)
        eng.synchronize()

    ms = timed_host(lind)
    out["lindblad_t1"] = {"realisations": S, "knots": N - 1, "gate_count": 1, "ms": ms, "density_steps_per_s": S * (N - 1) / ms * 1e3,
                          "api": "rbq_lindblad_dev (device resident)"}
    return out


def run_reference(args):
    """--impl reference: the reference's CPU path.  The reference is Julia (absent here), so this is the
    oracle port of its algorithm, all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    from rbqoc_b200 import spin

    orc.build()
    controls = spin.pulse_sin(NK + 1)[:-1]
    threads = host_threads()
    budget = 60.0 / max(args.steps + args.warmup, 1)
    rates, S_used = [], 0
    for i in range(args.warmup + args.steps):
        rate, S_used, _ = cpu_reference_rate(orc, controls, threads, target_seconds=min(budget, 10.0))
        if i >= args.warmup:
            rates.append(rate)
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": "fp64_trajectory_steps_per_sec", "value": value, "unit": "trajectory-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * S_used * NK / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"mc_static P_sin({NK + 1}) dt={DT} gate_count={GATE_COUNT}, bounded sample of {S_used} samples per step",
                   "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": value, "unit": "trajectory-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{S_used} samples x {NK} knots per step, OpenMP over samples"},
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
    seed = 2026
    with torch.cuda.stream(stream):
        d_controls = torch.from_numpy(controls_h).to(dev)
        d_fq = torch.empty(S, dtype=torch.float64, device=dev)
        d_da = torch.empty(S, dtype=torch.float64, device=dev)
        d_psi0 = torch.empty(S, 4, dtype=torch.float64, device=dev)
        d_fid = torch.empty(S, GATE_COUNT + 1, dtype=torch.float64, device=dev)
        d_stats = torch.zeros(rbq.STATS_INT64_WORDS, dtype=torch.int64, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    # the sample set: Philox keyed by global index (materialised once; resident in HBM for `value`)
    fq_h, da_h, psi0_h = eng.philox_samples(first, S, seed, spin.FQ, 0.01, 0.03, 2.574e-5)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    fq_p, da_p, psi0_p = pin(fq_h), pin(da_h), pin(psi0_h)
    fid_p = torch.empty(S, GATE_COUNT + 1, dtype=torch.float64).pin_memory()
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

    # ---- e2e: host-pointer C-ABI call, pinned host buffers, copies inside the timed region ----
    fid_np = fid_p.numpy()
    for _ in range(max(1, args.warmup // 2)):
        eng.mc_static(controls_h, DT, rbq.GATE_ZPIBY2, GATE_COUNT, fq_p.numpy(), da_p.numpy(), psi0_p.numpy(), fid_out=fid_np)
    barrier()
    e2e_steps = max(3, args.steps // 2)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _, st_e2e = eng.mc_static(controls_h, DT, rbq.GATE_ZPIBY2, GATE_COUNT, fq_p.numpy(), da_p.numpy(), psi0_p.numpy(),
                                  fid_out=fid_np)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
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

    # ---- reduce over ranks (max time), print on rank 0 ----
    times = torch.tensor([dev_ms, e2e_s * 1e3, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(times, op=tdist.ReduceOp.MAX)
    dev_ms, e2e_ms, kern_ms = (float(x) for x in times.cpu())
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
        line = {
            "metric": "fp64_trajectory_steps_per_sec", "value": value, "unit": "trajectory-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"mc_static sweep (BASELINE configs[4]): P_sin({NK + 1}) = {NK} knots, dt={DT} ns, "
                                   f"S={S} samples per GPU (Philox keyed by global index; fq~N(0,1%) clipped 3%, "
                                   f"da~N(0,2.574e-5)), gate_count={GATE_COUNT}, per-sample fidelities + statistics",
                       "samples_per_gpu": S, "knots": NK,
                       "algo": ("su2 closed form, tan form, knot-table step (local quadratic of tan(theta)/theta around the sweep centre)"
                                if TABLE_CLASS else "su2 closed form, tan form, degree-6 polynomial step"), "l2": "flushed between steps (256 MiB memset, untimed)",
                       "sharding": ("contiguous sample ranges per rank; per step ONE all-gather of the 2.1 KB statistics, copied to pinned host "
                                    "memory and merged on the host while the next step computes") if world > 1 else "single GPU"},
            "wall_s_timed_region": t_wall,
            "roofline": {"bound": "fp64", "achieved": ach_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": ach_tflops / peak_tflops if peak_tflops else None,
                         "traffic": NCU_DRAM_TRAFFIC_BYTES_DEFAULT if (S == S_PER_GPU) else None,
                         "traffic_unit": "bytes per launch (ncu dram read+write, profiles/r1b_ncu_mc_static_full.csv)",
                         "fp64_pipe_issue_util": S * NK * FP64_INSTR_PER_STEP_SU2 / (kern_ms * 1e-3) * 2e-12 / peak_tflops if peak_tflops else None,
                         "peak_source": "measured live: rbq_fp64_peak DFMA saturation loop (MEASURED_PEAKS.json has no FP64 entry)",
                         # register-operand model of the FP64 pipe (profiles/r1_fp64_microbench.txt): 2 issue cycles per warp-DFMA
                         # with <= 2 fresh 64-bit register operands, ~3 with 3 (the register file feeds two per slot)
                         "issue_model": {"cycles_per_warp_step_floor": (FP64_INSTR_PER_STEP_SU2 - FP64_INSTR3_PER_STEP_SU2) * 2 + FP64_INSTR3_PER_STEP_SU2 * 3,
                                         "cycles_per_warp_step_measured": (clk["sm_mhz"] or 1965.0) * 1e6 * 4 * 148 * (kern_ms * 1e-3) / (S * NK / 32.0),
                                         "note": "floor / measured = fraction of the operand-bandwidth-limited issue rate"},
                         "flops_per_trajectory_step": FLOPS_PER_STEP_SU2, "fp64_instr_per_trajectory_step": FP64_INSTR_PER_STEP_SU2,
                         "kernel_ms": kern_ms,
                         "hbm": {"achieved_gbs": S * BYTES_PER_SAMPLE / (kern_ms * 1e-3) * 1e-9, "peak_gbs": peaks.get("hbm_gbs"),
                                 "note": "algorithmic bytes; shows the path is not HBM bound"}},
            "e2e": {"value": steps_per_step_all / (e2e_ms * 1e-3), "unit": "trajectory-steps/s",
                    "h2d_bytes_per_step": S * 48 + NK * 8, "d2h_bytes_per_step": S * 16 + 2096, "ms_per_step": e2e_ms,
                    "numa_bound_cpus": len(numa_cpus) if numa_cpus else None,
                    "api": ("rbq_mc_static (host pointers, pinned): " + ("streamed in one-wave chunks" if os.environ.get("RBQ_MC_ZEROCOPY") == "0"
                                                                          else "zero-copy, the kernel reads the samples and writes the fidelities over PCIe")),
                    "pcie_measured": {"h2d_gbs": h2d_gbs, "d2h_gbs": d2h_gbs,
                                      "h2d_floor_ms": (S * 48 + NK * 8) / (h2d_gbs * 1e6) if h2d_gbs else None}},
            "gpu_launches": launches,
            "clocks": clk,
            "result": {"mean_infidelity": st.mean, "count": st.count,
                       "merged_count_all_ranks": merged["last"].count if merged["last"] is not None else None},
        }
        if world == 1:
            try:
                line["other_configs"] = other_configs(eng, rbq, spin, torch, dev)
            except Exception as exc:   # context only: never let it break the contract line
                line["other_configs"] = {"error": repr(exc)}
        if not args.no_cpu_baseline:
            from oracle import oracle as orc

            orc.build()
            threads = host_threads()
            rate, S_cpu, t_cpu = cpu_reference_rate(orc, controls_h, threads)
            line["cpu_baseline"] = {"value": rate, "unit": "trajectory-steps/s", "cores": threads, "kind": "port",
                                    "sample": f"{S_cpu} samples x {NK} knots of the same workload in {t_cpu:.1f} s "
                                              "(reference algorithm: generic Pade exp + 4x4 product per knot, OpenMP over samples)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
