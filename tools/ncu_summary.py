#!/usr/bin/env python
"""Key metrics of every kernel launch in an ncu report (--set full), as a markdown table + stall-reason breakdown.

    python tools/ncu_summary.py gpurun_out/r2_acw2.ncu-rep [> profiles/r02_xxx.md]
Needs `ncu` on PATH (reads the report on the CPU box; no GPU required).
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("smsp__inst_executed.sum", "warp instr"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("| # | kernel | " + " | ".join(k[1] for k in KEYS) + " | top stall reasons (share of samples) |")
    print("|---|---|" + "---|" * (len(KEYS) + 1))
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    for n, r in enumerate(rows[2:]):
        name = r[col["Kernel Name"]].split("(")[0]
        vals = []
        for k, _ in KEYS:
            i = col.get(k)
            vals.append(f"{r[i]} {units[i]}".strip() if i is not None and r[i] else "-")
        st = [(hdr[i].replace("smsp__pcsamp_warps_issue_stalled_", ""), float(r[i] or 0)) for i in stall_cols]
        tot = sum(v for _, v in st) or 1.0
        top = ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in sorted(st, key=lambda kv: -kv[1])[:5])
        print(f"| {n} | `{name}` | " + " | ".join(vals) + f" | {top} |")


if __name__ == "__main__":
    main()
