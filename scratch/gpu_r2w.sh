#!/bin/bash
# round 2, call W: final build (fc_out activation fragments in registers): full GPU suite, bench (with CPU baseline), launch list of one generate() call (decode DRAM traffic of the final kernel source)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2w_pytest.log
python profiles/run_profile.py --iters 2 --max-len 150 > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2w.csv python profiles/run_profile.py --iters 2 --max-len 150 > gpurun_out/prof_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2w_bench_c2.json 2> gpurun_out/r2w_bench_c2.err; echo "bench c2 rc=$?"; tail -c 300 gpurun_out/r2w_bench_c2.err
python -c "import json;d=json.load(open('gpurun_out/r2w_bench_c2.json'));print(d['value'],d['e2e']['value'],d['encoder_ms'],d['decode_ms'],d['roofline']['frac'])"
timeout 600 python scratch/beam_repeat_probe3.py 200 256 0 > gpurun_out/r2w_repeat_greedy256.txt 2>&1; tail -1 gpurun_out/r2w_repeat_greedy256.txt
