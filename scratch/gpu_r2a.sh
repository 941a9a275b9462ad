#!/bin/bash
# round 2, call A: full GPU test suite, the four bench configs, ncu launch list + GEMM tensor-pipe metrics
mkdir -p gpurun_out
rm -f gpurun_out/parity_r2.json
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_c2.json 2> gpurun_out/r2a_bench_c2.err; echo "bench c2 rc=$?"; tail -c 600 gpurun_out/r2a_bench_c2.err
for c in 3 4 5; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2a_bench_c$c.json 2> gpurun_out/r2a_bench_c$c.err; echo "bench c$c rc=$?"; tail -c 400 gpurun_out/r2a_bench_c$c.err
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_bench_ref.json 2>&1; echo "bench ref rc=$?"
python profiles/gemm_shapes.py --ncu > gpurun_out/gemm_order.txt 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:gemm --csv --log-file gpurun_out/gemm_tensor.csv python profiles/gemm_shapes.py --ncu > gpurun_out/gemm_ncu.log 2>&1; echo "gemm ncu rc=$?"
python profiles/run_profile.py --iters 2 --max-len 150 > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2a.csv python profiles/run_profile.py --iters 2 --max-len 150 > gpurun_out/prof_ncu.log 2>&1; echo "launch list rc=$?"
