#!/bin/bash
# round 2, call Q: final GELU coefficients: kernel tests + model tests + parity + bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2q_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2q_bench_c2.json 2> gpurun_out/r2q_bench_c2.err; echo "bench c2 rc=$?"; tail -c 300 gpurun_out/r2q_bench_c2.err
python -c "import json;d=json.load(open('gpurun_out/r2q_bench_c2.json'));print(d['value'],d['e2e']['value'],d['encoder_ms'],d['decode_ms'])"
