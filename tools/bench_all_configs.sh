mkdir -p gpurun_out
for c in 2 1 3 4 5; do
  extra="--no-cpu-baseline"; [ $c = 2 ] && extra=""
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 $extra > gpurun_out/r02f_bench_config${c}_1gpu.json 2> gpurun_out/r02f_bench_config${c}_1gpu.err
  python - <<P
import json
try:
  d=json.loads(open("gpurun_out/r02f_bench_config${c}_1gpu.json").read().strip().splitlines()[-1])
  r=d["roofline"]; h=d["roofline_hbm"]
  print("config ${c}", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "ms/step", round(d["ms_per_step"],1), "frac", round(r["frac"],3), "frac+gn", round(r.get("frac_conv_plus_groupnorm") or 0,3), "e2e_tensor", round(d.get("tensor_frac_end_to_end") or 0,3), "gn_frac", round(h["groupnorm"]["frac"],3), "upd", round(h["update"]["frac"],3), r["ms_by_kernel_family"], d["clocks"], d.get("cpu_baseline",{}) and d["cpu_baseline"].get("value"))
except Exception as e:
  print("config ${c} FAILED", e); print(open("gpurun_out/r02f_bench_config${c}_1gpu.err").read()[-800:])
P
done
