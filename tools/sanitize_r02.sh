#!/bin/bash
# round 2: compute-sanitizer memcheck over the smoke sampler call, the GroupNorm-epilogue / CTA-pair / halo kernel tests,
# the resampling + blur / DCT kernel tests; racecheck on one forward of the small net
mkdir -p gpurun_out
run() {  # name, tool, command...
  local name=$1 tool=$2; shift 2
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 "$@" > gpurun_out/r02_sanitize_$name.log 2>&1
  echo "== $name ($tool)"; grep -E "ERROR SUMMARY|passed|failed|smoke:" gpurun_out/r02_sanitize_$name.log | tail -4
}
run smoke memcheck python -c "import __graft_entry__ as g; g.smoke()"
run gnf_pairs_halo memcheck python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "groupnorm_epilogue or halo or pairs"
run resample_blur memcheck python -m pytest tests/test_ref_golden.py tests/test_gpu_kernels.py -m gpu -q -x -k "resampl or blur or dct or group_norm"
run forward racecheck python -c "
import sys; sys.path.insert(0,'tests')
import numpy as np
from helpers import build
cfg, model, net_fn = build('cld_deep')
x = np.random.default_rng(0).standard_normal((2,32,32,6)).astype(np.float32)
print(float(np.abs(model.forward(x, 0.5)).max()))
"
