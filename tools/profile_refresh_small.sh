#!/bin/bash
# DRAM-traffic pass over the conv_gemm_umma launches of the 2nd evaluation ($1 = launches per evaluation)
NG=${1:-168}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_gemm_umma \
    -s $NG -c $NG --csv --log-file gpurun_out/gemm_traffic.csv python tools/prof_forward.py 256 > gpurun_out/prof2.log 2>&1
tail -2 gpurun_out/prof2.log
