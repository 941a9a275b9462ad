#!/bin/bash
# round 2, call N: 16-row decode clusters (one 512-thread CTA per SM) + 8-row clusters for the remainder
mkdir -p gpurun_out
timeout 600 python scratch/wide_sweep.py 256 > gpurun_out/r2n_wide_sweep0.txt 2>&1; cat gpurun_out/r2n_wide_sweep0.txt | tail -5
timeout 1200 python -m pytest tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2n_pytest.log
timeout 600 python scratch/wide_sweep.py 128 144 160 176 192 208 224 240 248 256 264 > gpurun_out/r2n_wide_sweep.txt 2>&1; cat gpurun_out/r2n_wide_sweep.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2n_bench_c2.json 2> gpurun_out/r2n_bench_c2.err; echo "bench c2 rc=$?"; tail -c 300 gpurun_out/r2n_bench_c2.err
python -c "import json;d=json.load(open('gpurun_out/r2n_bench_c2.json'));print(d['value'],d['encoder_ms'],d['decode_ms'])"
