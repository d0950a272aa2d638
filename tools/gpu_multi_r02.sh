#!/bin/bash
# Round-2 multi-GPU pass (run through `gpurun --gpus N`): the 2-GPU bit-identity test of the sharded entry point and the
# strong-scaling bench lines of BASELINE.json configs[2] (230 ragged recordings) and configs[3] (100,000 x 2 s clips).
#   usage: bash tools/gpu_multi_r02.sh N [configs...]
set -u
N="$1"; shift
CFGS="${*:-2 3}"
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpu" > gpurun_out/gputest_2gpu.log 2>&1
  echo "2-GPU test rc=$?"; tail -3 gpurun_out/gputest_2gpu.log
fi
for c in $CFGS; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --config $c --steps 2 --warmup 1 > gpurun_out/bench_config${c}_n$N.json 2> gpurun_out/bench_config${c}_n$N.err
  echo "config $c N=$N rc=$?"; tail -c 600 gpurun_out/bench_config${c}_n$N.json
done
