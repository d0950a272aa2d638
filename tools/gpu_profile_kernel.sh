#!/bin/bash
# One `ncu --set full` capture of the launches of ONE kernel inside an extraction pass (tools/profile_pass.py), exported as
# CSV pages into gpurun_out/ (the .ncu-rep stays on the box: gpurun_out/ is limited to 64 MiB).
#   usage: bash tools/gpu_profile_kernel.sh <kernel-name-regex> [tag] [clips] [seconds]
set -u
K="$1"; TAG="${2:-$1}"; N="${3:-32}"; D="${4:-30}"
mkdir -p gpurun_out
timeout 300 python tools/profile_pass.py $N $D > gpurun_out/profile_pass_$TAG.log 2>&1 || { tail -5 gpurun_out/profile_pass_$TAG.log; exit 1; }
REP=/tmp/prof_$TAG
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:$K" -o $REP -f \
    python tools/profile_pass.py $N $D > gpurun_out/ncu_$TAG.log 2>&1
echo "capture rc=$?"
ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv 2>> gpurun_out/ncu_$TAG.log
ncu -i $REP.ncu-rep --page source --csv --kernel-id "::regex:$K:1" > gpurun_out/src_$TAG.csv 2>> gpurun_out/ncu_$TAG.log
ls -la gpurun_out/raw_$TAG.csv gpurun_out/src_$TAG.csv
