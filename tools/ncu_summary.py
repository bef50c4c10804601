#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i`, no GPU needed): headline metrics of every captured
launch and, with --source, the hottest SASS lines by stall samples.
usage: python tools/ncu_summary.py REPORT.ncu-rep [--source N]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:100])
        for i, h in enumerate(hdr):
            if h in KEYS or "issue_stalled" in h and "ratio" in h and "not_issued" not in h:
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if "issue_stalled" in h and v < 0.3:
                    continue
                print(f"   {h:90s} {r[i]:>16s} {units[i]}")
    if "--source" in sys.argv:
        n = int(sys.argv[sys.argv.index("--source") + 1])
        out = ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"])
        rows = list(csv.reader(io.StringIO(out)))
        h = None
        body = []
        for r in rows:
            if h is None:
                if "Source" in r and any("Sampling" in c for c in r):
                    h = r
                continue
            if r == h:
                continue   # the header repeats for every captured launch
            body.append(r)
        if h is None:
            print("no source page")
            return
        si = h.index("Source")
        samp = [i for i, c in enumerate(h) if c.startswith("# Samples") or c == "Warp Stall Sampling (All Samples)"]
        smp = samp[0] if samp else None
        ex = h.index("# Warp Instructions Executed") if "# Warp Instructions Executed" in h else None
        tot = sum(float(r[smp] or 0) for r in body if len(r) > smp)
        print("total samples", tot, "sass lines", len(body))
        ranked = sorted(range(len(body)), key=lambda k: -float(body[k][smp] or 0))[:n]
        for k in sorted(ranked):
            r = body[k]
            print(f"{k:5d} {float(r[smp] or 0) / tot * 100:5.1f}%  exec={r[ex] if ex is not None else ''}  {r[si][:110]}")


if __name__ == "__main__":
    main()
