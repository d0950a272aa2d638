#!/usr/bin/env python
"""Builds profiles/r02_flop_model.json from an ncu launch list of ONE extraction pass (tools/profile_pass.py).

    ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/flops.csv \
        --metrics gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,\
smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,\
dram__bytes_read.sum,dram__bytes_write.sum  python tools/profile_pass.py 32 30
    python tools/flop_model.py gpurun_out/flops.csv 960 "32 x 30 s synthetic clips"

Per kernel function: float64 FLOP (2 per DFMA + DMUL + DADD, predicated-on thread instructions), launches, device time under
ncu (cold, serialised: shares only) and DRAM bytes per launch.  bench.py multiplies flop_per_audio_second by the audio-seconds
of its timed step and divides by the live CUDA-event time: the FLOP count is a property of the algorithm on this kind of
signal, the time is measured in the run.
"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    return re.sub(r"\(.*$", "", name)


def main():
    path, audio_s = sys.argv[1], float(sys.argv[2])
    workload = sys.argv[3] if len(sys.argv) > 3 else f"{audio_s:g} audio-s"
    out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles", "r02_flop_model.json")
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    per_launch = collections.defaultdict(dict)
    names = {}
    for r in rd:
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        per_launch[r["ID"]][r["Metric Name"]] = v
        names[r["ID"]] = short(r["Kernel Name"])
    agg = collections.OrderedDict()
    for i, m in per_launch.items():
        k = agg.setdefault(names[i], {"launches": 0, "dfma": 0.0, "dmul": 0.0, "dadd": 0.0, "ns": 0.0, "dram": 0.0})
        k["launches"] += 1
        k["dfma"] += m.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0.0)
        k["dmul"] += m.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0.0)
        k["dadd"] += m.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0.0)
        k["ns"] += m.get("gpu__time_duration.sum", 0.0)
        k["dram"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    kernels = {}
    total = 0.0
    for name, k in agg.items():
        flop = 2.0 * k["dfma"] + k["dmul"] + k["dadd"]
        total += flop
        kernels[name] = {"launches": k["launches"], "flop": flop, "dfma": k["dfma"], "dmul": k["dmul"], "dadd": k["dadd"],
                         "ms_under_ncu": k["ns"] / 1e6, "dram_bytes_per_launch": k["dram"] / max(k["launches"], 1),
                         "flop_share": 0.0}
    for k in kernels.values():
        k["flop_share"] = k["flop"] / total if total else 0.0
    model = {"workload": workload, "audio_seconds": audio_s, "flop_per_audio_second": total / audio_s,
             "counted": "2 * dfma + dmul + dadd thread instructions (smsp__sass_thread_inst_executed_op_d*_pred_on.sum), one pass",
             "kernels": dict(sorted(kernels.items(), key=lambda kv: -kv[1]["flop"]))}
    with open(out, "w") as f:
        json.dump(model, f, indent=1)
    print(f"{total / audio_s / 1e9:.3f} GFLOP per audio-second over {len(kernels)} kernel functions -> {out}")
    for name, k in list(model["kernels"].items())[:16]:
        print(f"  {k['flop_share'] * 100:5.1f}%  {k['flop'] / audio_s / 1e6:9.2f} MFLOP/audio-s  {k['ms_under_ncu']:9.2f} ms  {name}")


if __name__ == "__main__":
    main()
