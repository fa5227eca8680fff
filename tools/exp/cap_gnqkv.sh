mkdir -p gpurun_out /tmp/cap
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:gn_qkv_persist -s 4 -c 1 -f -o /tmp/cap/gq python tools/prof_forward.py 256 > gpurun_out/cap.log 2>&1
ncu -i /tmp/cap/gq.ncu-rep --page raw --csv > gpurun_out/gq_raw.csv 2>> gpurun_out/cap.log
ncu -i /tmp/cap/gq.ncu-rep --page source --csv > gpurun_out/gq_source.csv 2>> gpurun_out/cap.log
ls -la gpurun_out/gq_*.csv
