mkdir -p gpurun_out /tmp/cap
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:attn256_kernel -s 4 -c 1 -f -o /tmp/cap/at python tools/prof_forward.py 256 > gpurun_out/cap_attn.log 2>&1
ncu -i /tmp/cap/at.ncu-rep --page raw --csv > gpurun_out/at_raw.csv 2>> gpurun_out/cap_attn.log
ncu -i /tmp/cap/at.ncu-rep --page source --csv > gpurun_out/at_source.csv 2>> gpurun_out/cap_attn.log
ls -la gpurun_out/at_*.csv
