#!/bin/bash
# round 2, call D: per-phase clock traces of the decode step (current build) at B = 256 / 64 / 8
mkdir -p gpurun_out
for b in 256 64 8; do
  python profiles/trace_step.py --batch $b --step 100 > gpurun_out/r2d_trace_b$b.txt 2>&1; echo "trace b=$b rc=$?"
done
python profiles/trace_step.py --batch 256 --step 20 > gpurun_out/r2d_trace_b256_t20.txt 2>&1
cat gpurun_out/r2d_trace_b256.txt
