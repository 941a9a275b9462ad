#!/bin/bash
# round 2, call J: fused MLP with the deeper weight ring at C = 96; full GPU suite on the build; bench; launch list of one generate() call (decode DRAM traffic)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2j_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2j_bench_c2.json 2> gpurun_out/r2j_bench_c2.err; echo "bench c2 rc=$?"; tail -c 300 gpurun_out/r2j_bench_c2.err
python -c "import json;d=json.load(open('gpurun_out/r2j_bench_c2.json'));print(d['value'],d['encoder_ms'],d['decode_ms'])"
python profiles/run_profile.py --iters 2 --max-len 150 > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2j.csv python profiles/run_profile.py --iters 2 --max-len 150 > gpurun_out/prof_ncu.log 2>&1; echo "launch list rc=$?"
grep -o 'swin_mlp_kernel<[0-9]*>.*gpu__time_duration.sum[^0-9]*[0-9.,]*"' gpurun_out/launches_r2j.csv | sed 's/(CUtensorMap_st.*gpu__time/ time/' | tail -4
