#!/bin/bash
# Runs on the GPU box (gpurun): staged so that a hang in the tcgen05 path cannot hide the results of the
# CUDA-core reference path.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== stage A: kernels without tcgen05 ==" | tee gpurun_out/ci.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "ref or group_norm or multistep or scalar or dct" 2>&1 | tail -25 | tee -a gpurun_out/ci.log
echo "== stage B: tcgen05 kernels ==" | tee -a gpurun_out/ci.log
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "umma or attention or gn_fused" 2>&1 | tail -40 | tee -a gpurun_out/ci.log
echo "== stage C: network + samplers ==" | tee -a gpurun_out/ci.log
timeout 900 python -m pytest tests/test_gpu_net.py tests/test_gpu_sampler.py tests/test_golden.py -m gpu -q -s > gpurun_out/stageC.log 2>&1; tail -30 gpurun_out/stageC.log | tee -a gpurun_out/ci.log
echo "== done ==" | tee -a gpurun_out/ci.log
if [ "$1" = "bench" ]; then
  echo "== bench ==" | tee -a gpurun_out/ci.log
  timeout 1200 python bench.py --steps 2 --warmup 3 --profile-csv gpurun_out/profile_ops.csv > gpurun_out/bench.json 2> gpurun_out/bench.err
  tail -c 3000 gpurun_out/bench.json | tee -a gpurun_out/ci.log; tail -5 gpurun_out/bench.err | tee -a gpurun_out/ci.log
fi
