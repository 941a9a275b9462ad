#!/bin/bash
# round 2, call K: window attention skips all-padding query tiles; --set full capture of the C = 96 fused MLP launch; parity case re-run
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider -k "window or encoder or swin" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest.log
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -p no:cacheprovider -k "peaked_16" > gpurun_out/r2k_pytest_par.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r2k_pytest_par.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2k_bench_c2.json 2> gpurun_out/r2k_bench_c2.err; echo "bench c2 rc=$?"; tail -c 300 gpurun_out/r2k_bench_c2.err
python -c "import json;d=json.load(open('gpurun_out/r2k_bench_c2.json'));print(d['value'],d['encoder_ms'],d['decode_ms'])"
ncu --set full --clock-control none --import-source on -k regex:swin_mlp -s 4 -c 1 -o gpurun_out/r2k_swin_mlp96 python profiles/run_profile.py --iters 2 --max-len 2 > gpurun_out/prof_ncu2.log 2>&1; echo "full capture rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:window_attn -c 24 --csv --log-file gpurun_out/launches_r2k_attn.csv python profiles/run_profile.py --iters 2 --max-len 2 > gpurun_out/prof_ncu3.log 2>&1
grep -o 'window_attn.*gpu__time_duration.sum[^0-9]*[0-9.,]*"' gpurun_out/launches_r2k_attn.csv | sed 's/(.*gpu__time/ time/' | tail -12
