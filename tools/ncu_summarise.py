"""Turns raw ncu output (gpurun_out/) into the small tracked summaries under profiles/.

  python tools/ncu_summarise.py launches gpurun_out/X_launches.csv profiles/Y_launches.csv "<command line>"
  python tools/ncu_summarise.py full     gpurun_out/X.ncu-rep      profiles/Y_full.csv     "<command line>"
"""
import collections
import csv
import subprocess
import sys

FULL_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "pcie__read_bytes.sum", "pcie__write_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed_pipe_tensor_op_dmma.sum", "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "sm__warps_active.avg.per_cycle_active",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def launches(src, dst, cmd):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui, gi, bi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3}.get(r[ui], 1.0)
        a = agg.setdefault((r[ki], r[gi], r[bi]), [0, 0.0])
        a[0] += 1
        a[1] += v
    with open(dst, "w") as f:
        f.write(f"# {cmd}\n# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write("kernel,launches,total_us,avg_us,grid,block\n")
        for (k, g, b), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'"{k}",{n},{t:.1f},{t / n:.2f},"{g}","{b}"\n')


def full(src, dst, cmd):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# {cmd}\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            f.write(f"# kernel: {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}\nmetric,value,unit\n")
            for k in FULL_KEYS:
                if k in d and d[k] != "":
                    f.write(f"{k},{d[k]},{u.get(k, '')}\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
