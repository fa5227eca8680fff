#!/bin/bash
# GPU box: PDL microbenchmark, then the sampler under the scheduling toggles (same box, back to back).
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pdl_bench tools/exp/pdl_bench.cu && /tmp/pdl_bench 2>&1 | tee gpurun_out/pdl_bench.txt
: > gpurun_out/ab.txt
for t in "GDDIM_PDL=0 GDDIM_ZIGZAG=0" "GDDIM_PDL=1 GDDIM_ZIGZAG=0" "GDDIM_PDL=0 GDDIM_ZIGZAG=1" "GDDIM_PDL=1 GDDIM_ZIGZAG=1" "GDDIM_PDL=0 GDDIM_ZIGZAG=0"; do
  env $t timeout 300 python tools/ab.py 256 3 2>&1 | grep -E "^AB|Error|error" | tee -a gpurun_out/ab.txt
done
