#!/bin/bash
# Round-2 profiling pass, run ON THE GPU BOX through gpurun (one GPU).  Every ncu command follows a plain run of
# the same command line (B200_PROFILING.md).  Outputs land in gpurun_out/; tools/measure_traffic.py and
# tools/launch_summary.py turn them into the summaries committed under profiles/.
set -x
FWD="python tools/run_forward.py --segments 2048 --iters 2"
$FWD > gpurun_out/r02_plain.log 2>&1 || exit 1
# (1) every launch of two forward passes with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv $FWD > gpurun_out/r02_ncu1.log 2>&1
# (2) DRAM bytes per kernel of ONE forward pass (bench.py's roofline.traffic)
ONE="python tools/run_forward.py --segments 2048 --iters 1"
$ONE > gpurun_out/r02_plain1.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/r02_traffic.csv $ONE > gpurun_out/r02_ncu2.log 2>&1
# (3) full captures: fbank, stats pooling, the bottleneck GEMM and the FCM slab kernel (one launch each, warm)
ncu --set full --clock-control none --import-source on -k regex:fbank_kernel -s 1 -c 1 -o gpurun_out/r02_prof_fbank $FWD > gpurun_out/r02_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stats_pool -s 1 -c 1 -o gpurun_out/r02_prof_statspool $FWD > gpurun_out/r02_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 60 -c 1 -o gpurun_out/r02_prof_gemm $FWD > gpurun_out/r02_ncu5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_slab3_kernel -s 14 -c 1 -o gpurun_out/r02_prof_slab3 $FWD > gpurun_out/r02_ncu6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cam_local_kernel -s 60 -c 1 -o gpurun_out/r02_prof_cam $FWD > gpurun_out/r02_ncu7.log 2>&1
tail -2 gpurun_out/r02_ncu*.log
