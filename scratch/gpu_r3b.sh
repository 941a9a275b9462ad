#!/bin/bash
# round 2, call 3b: automatic steps-per-launch (one launch when all clusters are co-resident) on top of the device-side early exit: full GPU suite, launch list (decode traffic), bench with CPU baseline, latency probe
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r3b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3b_pytest.log
python profiles/run_profile.py --iters 2 --max-len 150 > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r3b.csv python profiles/run_profile.py --iters 2 --max-len 150 > gpurun_out/prof_ncu.log 2>&1; echo "launch list rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r3b_bench_c2.json 2> gpurun_out/r3b_bench_c2.err; echo "bench c2 rc=$?"; tail -c 300 gpurun_out/r3b_bench_c2.err
python -c "import json;d=json.load(open('gpurun_out/r3b_bench_c2.json'));print(d['value'],d['e2e']['value'],d['encoder_ms'],d['decode_ms'],d['roofline']['frac'],d['p50_ms_per_image_b1'],d['gpu_launches'])"
for c in 3 5; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu > gpurun_out/r3b_bench_c$c.json 2> gpurun_out/r3b_bench_c$c.err; python -c "import json;d=json.load(open('gpurun_out/r3b_bench_c$c.json'));print($c, d['value'],d['ms_per_step'])"; done
timeout 600 python scratch/beam_repeat_probe3.py 100 256 0 > gpurun_out/r3b_repeat_greedy256.txt 2>&1; tail -1 gpurun_out/r3b_repeat_greedy256.txt
