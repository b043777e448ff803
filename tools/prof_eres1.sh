for m in eres:163 eres_w24:128; do
IFS=: read name seg <<< "$m"
CMD="python tools/run_forward.py --model $name --segments $seg --seconds 3.0 --iters 1"
$CMD > gpurun_out/r02c_${name}_plain.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02c_${name}_traffic.csv $CMD > gpurun_out/r02c_${name}_ncu.log 2>&1
tail -1 gpurun_out/r02c_${name}_plain.log
done
