#!/bin/bash
# round 2, call R: robustness sweeps on the final build: batch-invariance fuzz (22 batch sizes x 6 lengths), beam repeatability, B = 1 latency
mkdir -p gpurun_out
timeout 900 python scratch/fuzz_batch.py > gpurun_out/r2r_fuzz.txt 2>&1; tail -3 gpurun_out/r2r_fuzz.txt
timeout 600 python scratch/beam_repeat_probe3.py 200 64 5 > gpurun_out/r2r_repeat_beam5_64.txt 2>&1; tail -1 gpurun_out/r2r_repeat_beam5_64.txt
timeout 600 python scratch/beam_repeat_probe3.py 200 128 0 > gpurun_out/r2r_repeat_greedy128.txt 2>&1; tail -1 gpurun_out/r2r_repeat_greedy128.txt
python profiles/trace_step.py --batch 8 --step 100 > gpurun_out/r2r_trace_b8.txt 2>&1; head -3 gpurun_out/r2r_trace_b8.txt; tail -5 gpurun_out/r2r_trace_b8.txt
