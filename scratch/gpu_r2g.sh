#!/bin/bash
# round 2, call G: fused Swin MLP kernel (fc1 + GELU + fc2 + residual, hidden tile in TMEM / shared memory): unit tests, encoder tests, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "swin_mlp" > gpurun_out/r2g_pytest_mlp.log 2>&1; echo "mlp pytest rc=$?"; tail -15 gpurun_out/r2g_pytest_mlp.log
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider -k "encoder" > gpurun_out/r2g_pytest_enc.log 2>&1; echo "enc pytest rc=$?"; tail -8 gpurun_out/r2g_pytest_enc.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2g_bench_c2.json 2> gpurun_out/r2g_bench_c2.err; echo "bench c2 rc=$?"; tail -c 300 gpurun_out/r2g_bench_c2.err
python -c "import json;d=json.load(open('gpurun_out/r2g_bench_c2.json'));print(d['value'],d['encoder_ms'],d['decode_ms'])"
