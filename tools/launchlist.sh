#!/bin/bash
mkdir -p gpurun_out
N=$(python tools/prof_forward.py 256 | awk '/launches/{print $2}')
PER=$((N / 2))
SKIP=$((PER + 3))
echo "launches per forward: $PER" > gpurun_out/prof.log
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $PER --csv --log-file gpurun_out/launches.csv \
    python tools/prof_forward.py 256 >> gpurun_out/prof.log 2>&1
tail -2 gpurun_out/prof.log
