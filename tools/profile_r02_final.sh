#!/bin/bash
# Final round-2 ncu evidence for what changed after tools/profile_r02.sh ran (run under gpurun, 1 GPU):
#   r02f_launches.csv      launch list (durations) of the first launches of `python bench.py --steps 2 --warmup 1`
#   r02f_head_update       --set full: the head convolution of the deep network with the gDDIM update in its epilogue
#   r02f_head_plain        --set full: the same launch with GDDIM_NO_HEAD_UPDATE=1 (plain six-column epilogue) ...
#   r02f_cld_step          ... and the update kernel that then follows it
O=gpurun_out
T=/tmp/r02f_prof
mkdir -p $O $T
full() {  # name, kernel regex (mangled), skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:$rx -s $skip -c $cnt -f -o $T/r02f_$name "$@" >> $O/r02f_prof.log 2>&1
  ncu -i $T/r02f_$name.ncu-rep --page raw --csv > $O/r02f_${name}_raw.csv 2>> $O/r02f_prof.log
}
HEAD='conv_gemm_umma_kernelILi32ELi0ELi1ELi1ELb0'
full head_update $HEAD 2 1 python tools/prof_sampler.py cld 256 4 deep
full head_plain  $HEAD 2 1 env GDDIM_NO_HEAD_UPDATE=1 python tools/prof_sampler.py cld 256 4 deep
full cld_step    'cld_step_c3_kernel' 2 1 env GDDIM_NO_HEAD_UPDATE=1 python tools/prof_sampler.py cld 256 4 deep
if [ "$1" = "launches" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r02f_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r02f_bench_under_ncu.log 2>&1
fi
ls -la $O/r02f_* | awk '{print $5, $9}'
