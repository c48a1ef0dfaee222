"""Summarise an .ncu-rep into one CSV row per kernel launch (the columns profiles/*_summary.csv use).

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rN_x_summary.csv
"""
import csv
import io
import subprocess
import sys

COLS = [
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
]
STALL = "smsp__average_warps_issue_stalled_{}_per_issue_active.ratio"
STALLS = ["barrier", "branch_resolving", "dispatch_stall", "drain", "lg_throttle", "long_scoreboard",
          "math_pipe_throttle", "membar", "mio_throttle", "misc", "no_instruction", "not_selected",
          "selected", "short_scoreboard", "sleeping", "tex_throttle", "wait"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in COLS if c in idx]
    stalls = [s for s in STALLS if STALL.format(s) in idx]
    w = csv.writer(sys.stdout)
    w.writerow(["Kernel Name"] + cols + ["stall_" + s for s in stalls])
    w.writerow([""] + [units[idx[c]] for c in cols] + ["ratio"] * len(stalls))
    for r in data:
        w.writerow([r[idx["Kernel Name"]]] + [r[idx[c]] for c in cols] + [r[idx[STALL.format(s)]] for s in stalls])


if __name__ == "__main__":
    main()
