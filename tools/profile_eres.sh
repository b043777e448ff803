#!/bin/bash
# Launch list + DRAM bytes per kernel of one ERes2NetV2 extraction pass (3 s segments), both variants, and ECAPA 10 s.
set -x
for m in eres:256:3.0 eres_w24:128:3.0 ecapa:128:10.0; do
  IFS=: read name seg sec <<< "$m"
  CMD="python tools/run_forward.py --model $name --segments $seg --seconds $sec --iters 1"
  $CMD > gpurun_out/r02_${name}_plain.log 2>&1 || continue
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1500 --csv \
      --log-file gpurun_out/r02_${name}_traffic.csv $CMD > gpurun_out/r02_${name}_ncu.log 2>&1
done
python tools/bench_models.py eres eres_w24 ecapa > gpurun_out/r02_bench_models.json 2>&1
tail -3 gpurun_out/r02_bench_models.json
