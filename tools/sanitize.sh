#!/bin/bash
# compute-sanitizer over one small end-to-end sampler call (smoke) -- memcheck, then racecheck + synccheck on a forward
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_memcheck.log 2>&1
tail -6 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -c "
import sys; sys.path.insert(0,'tests')
import numpy as np
from helpers import build
cfg, model, net_fn = build('cld_deep')
x = np.random.default_rng(0).standard_normal((2,32,32,6)).astype(np.float32)
print(float(np.abs(model.forward(x, 0.5)).max()))
" > gpurun_out/sanitize_racecheck.log 2>&1
tail -6 gpurun_out/sanitize_racecheck.log
# CTA pairs (cta_group::2, remote mbarriers) and halo tiles at kernel level
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "halo or pairs" > gpurun_out/sanitize_pairs_halo.log 2>&1
tail -6 gpurun_out/sanitize_pairs_halo.log
