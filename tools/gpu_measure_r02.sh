#!/bin/bash
# Round-2 measurement pass (run on the GPU box through gpurun): the bench configs that are not the driver's default, an ncu
# launch list of one extraction pass with FP64 op counters + DRAM bytes (-> tools/flop_model.py), and one `--set full`
# capture of the frame / refinement kernels (-> tools/ncu_summary.py).  Every ncu command runs only after the same program
# exited 0 without ncu; numbers printed under ncu are never bench values.
#   usage: bash tools/gpu_measure_r02.sh [configs] [launches] [full]
set -u
mkdir -p gpurun_out
what="${*:-configs launches full}"
M="gpu__time_duration.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum"
K='regex:k_ac_frames_w|k_ac_candidates|k_pitch_frames|k_hnr_refine|k_pitch_refine|k_cepstrogram_w|k_cpp_frames_warp|k_formant_frames|k_spec_frames_w|k_sinc_fir|k_fft_inner|k_fft_strided|k_cc_frames_w|k_pulses_stretch|k_pitch_viterbi|k_ltas_accum|k_intensity_frames'
if [[ $what == *configs* ]]; then
  for c in 0 2 3 4; do
    timeout 600 python bench.py --config $c --steps 2 --warmup 1 > gpurun_out/bench_config$c.json 2> gpurun_out/bench_config$c.err
    echo "config $c rc=$?"
  done
fi
if [[ $what == *launches* ]]; then
  timeout 300 python tools/profile_pass.py 32 30 > gpurun_out/profile_pass.log 2>&1 && \
  timeout 900 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file gpurun_out/r02_launches_flops_ncu.csv \
      python tools/profile_pass.py 32 30 > gpurun_out/ncu_launches.log 2>&1
  echo "launch list rc=$?"
fi
if [[ $what == *full* ]]; then
  timeout 300 python tools/profile_pass.py 32 30 > gpurun_out/profile_pass.log 2>&1 || exit 1
  # the report itself (~150 MB) stays on the box: gpurun_out/ is limited to 64 MiB, so the pages are exported here
  REP=/tmp/r02_top_kernels
  timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k "$K" \
      -o $REP -f python tools/profile_pass.py 32 30 > gpurun_out/ncu_full.log 2>&1
  echo "full capture rc=$?"
  ncu -i $REP.ncu-rep --page raw --csv > gpurun_out/r02_top_kernels_raw.csv 2>> gpurun_out/ncu_full.log
  for k in k_ac_frames_w k_ac_candidates k_pitch_frames k_hnr_refine k_pitch_refine k_cepstrogram_w k_cpp_frames_warp k_formant_frames \
           k_cc_frames_w k_sinc_fir k_fft_inner k_pulses_stretch k_ltas_accum k_intensity_frames k_spec_frames_w; do
    ncu -i $REP.ncu-rep --page source --csv --kernel-id "::regex:$k:1" > gpurun_out/r02_src_$k.csv 2>/dev/null
    [ -s gpurun_out/r02_src_$k.csv ] || rm -f gpurun_out/r02_src_$k.csv
  done
  du -sh gpurun_out
fi
