#!/usr/bin/env python
"""Turn an `ncu --set full` report into the markdown table kept under profiles/.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep [--px-per-launch N[,N...]] [--title "..."] > profiles/x_ncu_summary.md

One row per captured launch: duration, DRAM bytes, issue-slot and pipe utilisation, warps active, registers, the stall
reasons per issued instruction, shared-memory wavefronts / bank conflicts, and (with --px-per-launch) thread-instructions per pixel.
"""
import argparse
import csv
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration us", 1.0),
    ("launch__grid_size", "grid", 1.0),
    ("launch__registers_per_thread", "regs", 1.0),
    ("dram__bytes_read.sum", "DRAM read", 1.0),
    ("dram__bytes_write.sum", "DRAM write", 1.0),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %", 1.0),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue %", 1.0),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1.0),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma %", 1.0),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe cycles % (a packed FFMA2/FADD2/FMUL2 holds the pipe two cycles)", 1.0),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu %", 1.0),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu %", 1.0),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu %", 1.0),
    ("smsp__inst_executed.sum", "warp-instr", 1.0),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts", 1.0),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts", 1.0),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb", 1.0),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb", 1.0),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier", 1.0),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math", 1.0),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio", 1.0),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_sel", 1.0),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait", 1.0),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--px-per-launch", default="")
    ap.add_argument("--title", default="")
    ap.add_argument("--command", default="")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    px = [float(v) for v in args.px_per_launch.split(",") if v]
    print(f"# {args.title or args.report}\n")
    if args.command:
        print(f"Command: `{args.command}` (after the same command exited 0 without ncu). Times are cold-cache single launches under the profiler.\n")
    for k, r in enumerate(data):
        name = r[idx["Kernel Name"]].replace("polcue::<unnamed>::", "").replace("void ", "")
        print(f"## launch {k}: `{name[:110]}`\n")
        print("| metric | value | unit |\n|---|---|---|")
        for key, label, _ in METRICS:
            if key in idx:
                print(f"| {label} (`{key}`) | {r[idx[key]]} | {units[idx[key]]} |")
        if k < len(px) and px[k] > 0 and "smsp__inst_executed.sum" in idx:
            print(f"| thread-instructions per pixel | {float(r[idx['smsp__inst_executed.sum']].replace(',', '')) * 32 / px[k]:.1f} | at {px[k]:.0f} px |")
        print()


if __name__ == "__main__":
    main()
