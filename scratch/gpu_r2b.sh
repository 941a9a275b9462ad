#!/bin/bash
# round 2, call B: L2 eviction policies + half-block loads in the decode kernel: decode tests, bench, launch list (DRAM traffic), agreement cases
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_res18_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2b_pytest.log
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2b_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/r2b_parity.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err; echo "bench c2 rc=$?"; tail -c 300 gpurun_out/r2b_bench_c2.err
python bench.py --config 3 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2b_bench_c3.json 2> gpurun_out/r2b_bench_c3.err; echo "bench c3 rc=$?"
python profiles/run_profile.py --iters 2 --max-len 150 > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:decode_persistent -c 20 --csv --log-file gpurun_out/launches_r2b.csv python profiles/run_profile.py --iters 2 --max-len 150 > gpurun_out/prof_ncu.log 2>&1; echo "launch list rc=$?"
