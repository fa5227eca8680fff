#!/bin/bash
# full ncu capture of one kernel instantiation selected by its mangled-name fragment ($1), $2 = skip, $3 = count
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$1 -s ${2:-2} -c ${3:-2} -f \
    -o gpurun_out/prof_one python tools/prof_forward.py 256 > gpurun_out/prof_one.log 2>&1
tail -3 gpurun_out/prof_one.log
