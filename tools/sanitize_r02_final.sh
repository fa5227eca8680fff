#!/bin/bash
# final round-2 additions: compute-sanitizer memcheck over the smoke sampler call (update inside the head convolution's
# epilogue), the fused-vs-kernel update test, and the nf = 32 network (pixel-paired convolutions) forward + sampler tests
mkdir -p gpurun_out
run() {  # name, tool, command...
  local name=$1 tool=$2; shift 2
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 "$@" > gpurun_out/r02f_sanitize_$name.log 2>&1
  echo "== $name ($tool)"; grep -E "ERROR SUMMARY|passed|failed|smoke:" gpurun_out/r02f_sanitize_$name.log | tail -4
}
run smoke memcheck python -c "import __graft_entry__ as g; g.smoke()"
run head_update memcheck python -m pytest tests/test_gpu_sampler.py -m gpu -q -x -k "update_fused_into_head"
run nf32 memcheck python -m pytest tests/test_gpu_net.py tests/test_gpu_sampler.py -m gpu -q -x -k "nf32"
