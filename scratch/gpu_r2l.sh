#!/bin/bash
# round 2, call L: decode time vs steps_per_launch; repeatability probe of the final decode kernel at full occupancy
mkdir -p gpurun_out
python scratch/spl_sweep.py > gpurun_out/r2l_spl_sweep.txt 2>&1; cat gpurun_out/r2l_spl_sweep.txt
timeout 900 python scratch/beam_repeat_probe3.py 300 256 0 > gpurun_out/r2l_repeat_greedy256.txt 2>&1; tail -3 gpurun_out/r2l_repeat_greedy256.txt
timeout 900 python scratch/beam_repeat_probe3.py 200 45 5 > gpurun_out/r2l_repeat_beam5.txt 2>&1; tail -3 gpurun_out/r2l_repeat_beam5.txt
