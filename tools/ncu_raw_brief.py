#!/usr/bin/env python
"""Short per-launch summary of an exported ncu raw page (`ncu -i rep --page raw --csv`): time, registers, occupancy, issue slots,
FP64 pipe, shared-memory wavefronts, DRAM bytes, top stall reasons.

    python tools/ncu_raw_brief.py gpurun_out/raw_ccw.csv
"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "t"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("launch__block_size", "blk"), ("launch__occupancy_limit_registers", "occ_reg"), ("launch__occupancy_limit_shared_mem", "occ_smem"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("smsp__inst_executed.sum", "winst")]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    stall = [i for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
    for r in rows[2:]:
        print(r[col["Kernel Name"]].split("(")[0])
        print("   " + "  ".join(f"{n}={r[col[k]]}{units[col[k]] if n in ('t', 'dram_rd', 'dram_wr') else ''}" for k, n in KEYS if k in col))
        st = [(hdr[i].replace("smsp__pcsamp_warps_issue_stalled_", ""), float(r[i].replace(",", "") or 0)) for i in stall]
        tot = sum(v for _, v in st) or 1.0
        print("   stalls: " + ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in sorted(st, key=lambda kv: -kv[1])[:7]))


if __name__ == "__main__":
    main()
