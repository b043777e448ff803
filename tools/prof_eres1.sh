CMD="python tools/run_forward.py --model eres --segments 163 --seconds 3.0 --iters 1"
$CMD > gpurun_out/r02b_eres_plain.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02b_eres_traffic.csv $CMD > gpurun_out/r02b_eres_ncu.log 2>&1
tail -1 gpurun_out/r02b_eres_plain.log
