#!/bin/bash
# Round-2 profiling pass, run ON THE GPU BOX through gpurun (one GPU).  Every ncu command follows a plain run of
# the same command line (B200_PROFILING.md).  Outputs land in gpurun_out/; tools/measure_traffic.py,
# tools/launch_summary.py and tools/ncu_digest.py turn them into the summaries committed under profiles/.
set -x
export SPK_GRAPH_MAX_BATCH=0          # profile the launches themselves, not a graph replay of them
FWD="python tools/run_forward.py --segments 8192 --iters 2"
$FWD > gpurun_out/r02_plain.log 2>&1 || exit 1
# (1) every launch of two forward passes with its device time
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches.csv $FWD > gpurun_out/r02_ncu1.log 2>&1
# (2) DRAM bytes per kernel of ONE forward pass (bench.py's roofline.traffic)
ONE="python tools/run_forward.py --segments 8192 --iters 1"
$ONE > gpurun_out/r02_plain1.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/r02_traffic.csv $ONE > gpurun_out/r02_ncu2.log 2>&1
# (3) full captures (one launch each, warm): fbank, the bottleneck GEMM, the FCM slab kernel, the fused CAM layer
ncu --set full --clock-control none --import-source on -k regex:fbank_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_fbank $FWD > gpurun_out/r02_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 60 -c 1 -f -o gpurun_out/r02_prof_gemm $FWD > gpurun_out/r02_ncu5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_slab3_kernel -s 14 -c 1 -f -o gpurun_out/r02_prof_slab3 $FWD > gpurun_out/r02_ncu6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cam_local_kernel -s 60 -c 1 -f -o gpurun_out/r02_prof_cam $FWD > gpurun_out/r02_ncu7.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stem_block_kernel -s 4 -c 1 -f -o gpurun_out/r02_prof_stem $FWD > gpurun_out/r02_ncu10.log 2>&1
# (4) ERes2NetV2 (3 s segments): launch list with DRAM bytes, full captures of the column-part slab kernel and of an im2col GEMM
ER="python tools/run_forward.py --model eres --segments 163 --seconds 3.0 --iters 1"
$ER > gpurun_out/r02_eres_plain.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/r02_eres_traffic.csv $ER > gpurun_out/r02_eres_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_slab4_kernel -s 7 -c 1 -f -o gpurun_out/r02_prof_slab4 $ER > gpurun_out/r02_ncu8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 30 -c 1 -f -o gpurun_out/r02_prof_im2col $ER > gpurun_out/r02_ncu9.log 2>&1
# (5) ECAPA-TDNN (10 s chunks): launch list with DRAM bytes
EC="python tools/run_forward.py --model ecapa --segments 128 --seconds 10.0 --iters 1"
$EC > gpurun_out/r02_ecapa_plain.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/r02_ecapa_traffic.csv $EC > gpurun_out/r02_ecapa_ncu.log 2>&1
ls -la gpurun_out/r02_*
