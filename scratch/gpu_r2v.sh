#!/bin/bash
# round 2, call V: decode experiment: activation fragments of the vocabulary projection loaded once per step (shared-memory port)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2v_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2v_bench_c2.json 2> gpurun_out/r2v_bench_c2.err; echo "bench c2 rc=$?"; tail -c 300 gpurun_out/r2v_bench_c2.err
python -c "import json;d=json.load(open('gpurun_out/r2v_bench_c2.json'));print(d['value'],d['encoder_ms'],d['decode_ms'])"
python profiles/trace_step.py --batch 256 --step 100 > gpurun_out/r2v_trace_b256.txt 2>&1; cat gpurun_out/r2v_trace_b256.txt
