#!/bin/bash
# round 2, call S: two-stream pipeline through the public API (encoder of batch i+1 under the decode of batch i)
mkdir -p gpurun_out
timeout 600 python scratch/pipeline_probe.py > gpurun_out/r2s_pipeline_probe.txt 2>&1; cat gpurun_out/r2s_pipeline_probe.txt | tail -6
