#!/bin/bash
# round 2, call Y: encoder replayed as a CUDA graph at every batch size up to 1024 (host-side launch cost): tests + bench at N GPUs
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 1500 python -m pytest tests/test_model_gpu.py tests/test_res18_gpu.py tests/test_parity_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2y_pytest.log
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2y_bench_c2.json 2> gpurun_out/r2y_bench_c2.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2y_bench_c2.err
  python -c "import json;d=json.load(open('gpurun_out/r2y_bench_c2.json'));print(d['value'],d['e2e']['value'],d['encoder_ms'],d['decode_ms'],d['gpu_launches'])"
else
  for c in 2 4; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $c --steps 10 --warmup 3 --no-cpu > gpurun_out/r2y_bench_c${c}_${N}gpu.json 2> gpurun_out/r2y_bench_c${c}_${N}gpu.err; echo "c$c N=$N rc=$?"
    python -c "import json;d=json.loads(open('gpurun_out/r2y_bench_c${c}_${N}gpu.json').read().strip().splitlines()[-1]);print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'])"
  done
fi
