#!/bin/bash
# ncu evidence (run under gpurun, 1 GPU): launch list of one network evaluation + full captures of the hot kernels.
mkdir -p gpurun_out
N=$(python tools/prof_forward.py 256 | awk '/launches/{print $2}')
PER=$((N / 2))
echo "launches per forward: $PER" | tee gpurun_out/prof.log
SKIP=$((PER + 3))   # first forward + the 3 time-embedding launches of the second
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $PER --csv --log-file gpurun_out/launches.csv \
    python tools/prof_forward.py 256 >> gpurun_out/prof.log 2>&1
# conv3x3 128->128 @32x32 (block_n 128), conv3x3 256->256 @16x16 (block_n 256): launch indices inside the 2nd forward
ncu --set full --clock-control none --import-source on -k regex:conv_gemm_umma -s 120 -c 12 -f -o gpurun_out/prof_gemm \
    python tools/prof_forward.py 256 >> gpurun_out/prof.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gn_ -s 100 -c 8 -f -o gpurun_out/prof_gn \
    python tools/prof_forward.py 256 >> gpurun_out/prof.log 2>&1
tail -5 gpurun_out/prof.log
