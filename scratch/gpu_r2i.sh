#!/bin/bash
# round 2, call I: fused Swin MLP kernel, second version (biases in shared memory, residual rows prefetched to L2 and loaded before the accumulator wait, staging hand-off by mbarrier)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -p no:cacheprovider -k "swin_mlp" > gpurun_out/r2i_pytest_mlp.log 2>&1; echo "mlp pytest rc=$?"; tail -5 gpurun_out/r2i_pytest_mlp.log
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -x -p no:cacheprovider -k "encoder" > gpurun_out/r2i_pytest_enc.log 2>&1; echo "enc pytest rc=$?"; tail -4 gpurun_out/r2i_pytest_enc.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2i_bench_c2.json 2> gpurun_out/r2i_bench_c2.err; echo "bench c2 rc=$?"; tail -c 300 gpurun_out/r2i_bench_c2.err
python -c "import json;d=json.load(open('gpurun_out/r2i_bench_c2.json'));print(d['value'],d['encoder_ms'],d['decode_ms'])"
python profiles/run_profile.py --iters 2 --max-len 2 > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:swin_mlp -c 8 --csv --log-file gpurun_out/launches_r2i.csv python profiles/run_profile.py --iters 2 --max-len 2 > gpurun_out/prof_ncu.log 2>&1; echo "launch list rc=$?"
grep -o 'swin_mlp_kernel<[0-9]*>.*gpu__time_duration.sum[^0-9]*[0-9.,]*"' gpurun_out/launches_r2i.csv | sed 's/(CUtensorMap_st.*gpu__time/ time/' 
