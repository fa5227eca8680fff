#!/bin/bash
# ncu evidence (run under gpurun, 1 GPU) on two eager network evaluations of the deep NCSN++ at batch 256:
#   launches.csv      every launch of the 2nd evaluation with its duration (cold cache, serialised: compare shares)
#   gemm_traffic.csv  DRAM bytes + duration of every conv_gemm_umma launch of the 2nd evaluation
#   prof_*.ncu-rep    --set full captures of representative instances of the dominant kernels
mkdir -p gpurun_out
N=$(python tools/prof_forward.py 256 | awk '/launches/{print $2}')
PER=$((N / 2))
SKIP=$((PER + 3))
echo "launches per forward: $PER" | tee gpurun_out/prof.log
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $PER --csv --log-file gpurun_out/launches.csv \
    python tools/prof_forward.py 256 >> gpurun_out/prof.log 2>&1
NG=$(grep -c conv_gemm_umma gpurun_out/launches.csv)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_gemm_umma \
    -s $NG -c $NG --csv --log-file gpurun_out/gemm_traffic.csv python tools/prof_forward.py 256 >> gpurun_out/prof.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:conv_gemm_umma_kernelILi256ELi0ELi1 \
    -s 120 -c 6 -f -o gpurun_out/prof_gemm256 python tools/prof_forward.py 256 >> gpurun_out/prof.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:conv_gemm_umma_kernelILi128ELi0ELi2 \
    -s 45 -c 3 -f -o gpurun_out/prof_gemm128 python tools/prof_forward.py 256 >> gpurun_out/prof.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gn_apply_kernelILi0E \
    -s 160 -c 3 -f -o gpurun_out/prof_gn python tools/prof_forward.py 256 >> gpurun_out/prof.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn256_kernel \
    -s 11 -c 2 -f -o gpurun_out/prof_attn python tools/prof_forward.py 256 >> gpurun_out/prof.log 2>&1
tail -3 gpurun_out/prof.log
