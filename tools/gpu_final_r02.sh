#!/bin/bash
# Round-2 final pass on one GPU: the whole -m gpu suite, smoke(), the default bench line (with CPU leg and row check), the
# reference arm, and the bench lines of configs 0 / 2 / 3 / 4.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_gputest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/final_gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
timeout 600 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "reference arm rc=$?"
for c in 0 2 3 4; do
  timeout 600 python bench.py --config $c --steps 2 --warmup 1 > gpurun_out/final_config$c.json 2> gpurun_out/final_config$c.err
  echo "config $c rc=$?"
done
python - <<'PY'
import json
for f in ["final_bench", "final_bench_reference", "final_config0", "final_config2", "final_config3", "final_config4"]:
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d.get("value", 0), 1), d.get("unit"), "e2e", round(d.get("e2e", {}).get("value", 0), 1), d.get("parity", {}).get("parity_ok"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
