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
# flops the closed-form kernel executes per trajectory-step on this grid (tan form, degree-6 polynomial, class B):
#   1 DFMA (q = w a_j + w da) + 1 DFMA (x = q^2 + p^2) + 6 DFMA (T(x)) + 2 DMUL (T p, T q) + 8 DFMA (v += T G v)
#   = 18 FP64 instructions = 34 flop (FMA = 2); counted in the SASS: (128 DFMA*2 + 16 DMUL) / 8 sample-steps  [DESIGN.md §4.2]
FLOPS_PER_STEP_SU2 = 34.0
FP64_INSTR_PER_STEP_SU2 = 18.0
BYTES_PER_SAMPLE = 8 + 8 + 32 + 16   # fq, da, psi0 in; 2 fidelities out


def sample_clocks(stop, out):
    """One long-lived `nvidia-smi -lms 50` (spawning it per sample is too slow for a sub-second timed region)."""
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50", "-i",
                                 os.environ.get("LOCAL_RANK", "0")], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return

    def reader():
        for line in proc.stdout:
            out.append([c.strip() for c in line.split(",")])

    t = threading.Thread(target=reader, daemon=True)
    t.start()
    stop.wait()
    proc.terminate()
    t.join(timeout=2)


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
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
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
    if world > 1:
        import torch.distributed as tdist

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

    def device_step():
        eng.stats_reset_dev(d_stats)
        eng.mc_static_dev(d_controls, NK, DT, rbq.GATE_ZPIBY2, GATE_COUNT, S, d_fq, d_da, d_psi0, rbq.ALGO_SU2, d_fid, d_stats)
        if world > 1:
            with torch.cuda.stream(stream):
                return rdist.allgather_merge_stats(d_stats)
        return None

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
    for _ in range(args.warmup):
        with torch.cuda.stream(stream):
            flush.zero_()
        device_step()
    barrier()
    launches0 = eng.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for a, b in ev:
        with torch.cuda.stream(stream):
            flush.zero_()
            a.record(stream)
        device_step()
        with torch.cuda.stream(stream):
            b.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = eng.launch_count - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
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

    # ---- reduce over ranks (max time), print on rank 0 ----
    times = torch.tensor([dev_ms, e2e_s * 1e3, kern_ms], dtype=torch.float64, device=dev)
    if world > 1:
        tdist.all_reduce(times, op=tdist.ReduceOp.MAX)
    dev_ms, e2e_ms, kern_ms = (float(x) for x in times.cpu())
    stream.synchronize()
    st = rbq.stats_from_words(d_stats.cpu().numpy())
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
        line = {
            "metric": "fp64_trajectory_steps_per_sec", "value": value, "unit": "trajectory-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"mc_static sweep (BASELINE configs[4]): P_sin({NK + 1}) = {NK} knots, dt={DT} ns, "
                                   f"S={S} samples per GPU (Philox keyed by global index; fq~N(0,1%) clipped 3%, "
                                   f"da~N(0,2.574e-5)), gate_count={GATE_COUNT}, per-sample fidelities + statistics",
                       "samples_per_gpu": S, "knots": NK, "algo": "su2 closed form (tan form, degree-6 polynomial)", "l2": "flushed between steps (256 MiB memset, untimed)",
                       "sharding": "contiguous sample ranges per rank, one all-gather of 2.1 KB statistics" if world > 1 else "single GPU"},
            "wall_s_timed_region": t_wall,
            "roofline": {"bound": "fp64", "achieved": ach_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": ach_tflops / peak_tflops if peak_tflops else None, "traffic": None,
                         "fp64_pipe_issue_util": S * NK * FP64_INSTR_PER_STEP_SU2 / (kern_ms * 1e-3) * 2e-12 / peak_tflops if peak_tflops else None,
                         "peak_source": "measured live: rbq_fp64_peak DFMA saturation loop (MEASURED_PEAKS.json has no FP64 entry)",
                         "flops_per_trajectory_step": FLOPS_PER_STEP_SU2, "fp64_instr_per_trajectory_step": FP64_INSTR_PER_STEP_SU2,
                         "kernel_ms": kern_ms,
                         "hbm": {"achieved_gbs": S * BYTES_PER_SAMPLE / (kern_ms * 1e-3) * 1e-9, "peak_gbs": peaks.get("hbm_gbs"),
                                 "note": "algorithmic bytes; shows the path is not HBM bound"}},
            "e2e": {"value": steps_per_step_all / (e2e_ms * 1e-3), "unit": "trajectory-steps/s",
                    "h2d_bytes_per_step": S * 48 + NK * 8, "d2h_bytes_per_step": S * 16 + 2096, "ms_per_step": e2e_ms,
                    "api": "rbq_mc_static (host pointers, pinned)"},
            "gpu_launches": launches,
            "clocks": clocks_summary(rows),
            "result": {"mean_infidelity": st.mean, "count": st.count},
        }
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
