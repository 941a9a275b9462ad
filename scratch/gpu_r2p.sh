#!/bin/bash
# round 2, call P: degree-5 unclamped GELU polynomial + packed bias add in the fused MLP: kernel tests, encoder tests, bench, MLP launch times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2p_pytest_k.log 2>&1; echo "kernels rc=$?"; tail -3 gpurun_out/r2p_pytest_k.log
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_res18_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2p_pytest_m.log 2>&1; echo "model rc=$?"; tail -3 gpurun_out/r2p_pytest_m.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2p_bench_c2.json 2> gpurun_out/r2p_bench_c2.err; echo "bench c2 rc=$?"; tail -c 300 gpurun_out/r2p_bench_c2.err
python -c "import json;d=json.load(open('gpurun_out/r2p_bench_c2.json'));print(d['value'],d['encoder_ms'],d['decode_ms'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"swin_mlp|gemm" -c 120 --csv --log-file gpurun_out/launches_r2p.csv python profiles/run_profile.py --iters 1 --max-len 2 > gpurun_out/prof_ncu.log 2>&1
grep -o 'swin_mlp_kernel<[0-9]*>.*gpu__time_duration.sum[^0-9]*[0-9.,]*"' gpurun_out/launches_r2p.csv | sed 's/(CUtensorMap_st.*gpu__time/ time/' | tail -4
