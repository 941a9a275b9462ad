#!/bin/bash
# round 2, call 3e: ncu --set full of the stage-1 window attention launch and of one stage-3 fc1 GEMM launch (final build)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:window_attn -s 0 -c 1 -o gpurun_out/r3e_window_attn_s1 python profiles/run_profile.py --iters 1 --max-len 2 > gpurun_out/prof_ncu_wa.log 2>&1; echo "attn capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05_kernel<256" -s 4 -c 1 -o gpurun_out/r3e_gemm_s3_fc1 python profiles/run_profile.py --iters 1 --max-len 2 > gpurun_out/prof_ncu_g.log 2>&1; echo "gemm capture rc=$?"
ls -la gpurun_out/r3e_*.ncu-rep
