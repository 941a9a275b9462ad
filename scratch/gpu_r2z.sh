#!/bin/bash
# round 2, call Z: ncu --set full capture of one decode launch (steps 96..111) of the FINAL kernel, and of the stage-1 fused MLP
mkdir -p gpurun_out
python profiles/run_profile.py --iters 1 --max-len 150 > gpurun_out/prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:decode_persistent -s 6 -c 1 -o gpurun_out/r2z_decode_persistent python profiles/run_profile.py --iters 1 --max-len 150 > gpurun_out/prof_ncu_dp.log 2>&1; echo "decode capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:swin_mlp -s 0 -c 1 -o gpurun_out/r2z_swin_mlp96 python profiles/run_profile.py --iters 1 --max-len 2 > gpurun_out/prof_ncu_mlp.log 2>&1; echo "mlp capture rc=$?"
ls -la gpurun_out/*.ncu-rep
