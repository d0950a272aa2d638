#!/usr/bin/env python
"""Markdown table per kernel FUNCTION (launches of one extraction pass aggregated, time-weighted) from an exported ncu raw page
(`ncu -i rep --page raw --csv`, see tools/gpu_measure_r02.sh): time under ncu and share, registers, occupancy limiters, warps
active, issue slots, FP64 pipe, shared-memory wavefronts, DRAM bytes per launch, top stall reasons.

    python tools/ncu_raw_table.py gpurun_out/r02_top_kernels_raw.csv > profiles/r02_top_kernels.md
"""
import collections
import csv
import re
import sys

M = {"t": "gpu__time_duration.sum", "regs": "launch__registers_per_thread", "warps": "sm__warps_active.avg.pct_of_peak_sustained_active",
     "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active", "fp64": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
     "smem": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "rd": "dram__bytes_read.sum",
     "wr": "dram__bytes_write.sum", "grid": "launch__grid_size", "blk": "launch__block_size"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    stall = [i for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]

    def val(r, key):
        i = col[M[key]]
        v = float(r[i].replace(",", "") or 0)
        return v * UNIT.get(units[i], 1.0)

    agg = collections.OrderedDict()
    for r in rows[2:]:
        name = re.sub(r"^void\s+", "", r[col["Kernel Name"]].split("(")[0])
        a = agg.setdefault(name, {"n": 0, "t": 0.0, "w": collections.Counter(), "st": collections.Counter(), "regs": 0, "grid": 0, "blk": 0,
                                  "rd": 0.0, "wr": 0.0})
        t = val(r, "t")
        a["n"] += 1; a["t"] += t
        for k in ("warps", "issue", "fp64", "smem"):
            a["w"][k] += val(r, k) * t
        a["rd"] += val(r, "rd"); a["wr"] += val(r, "wr")
        a["regs"] = int(val(r, "regs")); a["grid"] = max(a["grid"], int(val(r, "grid"))); a["blk"] = int(val(r, "blk"))
        for i in stall:
            a["st"][hdr[i].replace("smsp__pcsamp_warps_issue_stalled_", "")] += float(r[i].replace(",", "") or 0)
    total = sum(a["t"] for a in agg.values())
    print("| kernel | launches | ms under ncu | share | regs | block | warps active % | issue slots % | FP64 pipe % | smem wavefronts % | DRAM MB / launch (rd + wr) | top stall reasons |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        t = a["t"] or 1e-12
        tot = sum(a["st"].values()) or 1.0
        top = ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in a["st"].most_common(4))
        print(f"| `{name}` | {a['n']} | {a['t']:.2f} | {100 * a['t'] / total:.1f}% | {a['regs']} | {a['blk']} | {a['w']['warps'] / t:.1f} | "
              f"{a['w']['issue'] / t:.1f} | {a['w']['fp64'] / t:.1f} | {a['w']['smem'] / t:.1f} | {(a['rd'] + a['wr']) / a['n'] / 1e6:.0f} | {top} |")


if __name__ == "__main__":
    main()
