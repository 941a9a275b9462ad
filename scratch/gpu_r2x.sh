#!/bin/bash
# round 2, call X: multi-GPU bench lines of every BASELINE config on the final build (N = $1 GPUs)
N=$1
mkdir -p gpurun_out
for c in 2 3 4 5; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $c --steps 6 --warmup 3 --no-cpu > gpurun_out/r2x_bench_c${c}_${N}gpu.json 2> gpurun_out/r2x_bench_c${c}_${N}gpu.err; echo "c$c N=$N rc=$?"
  python -c "import json;d=json.loads(open('gpurun_out/r2x_bench_c${c}_${N}gpu.json').read().strip().splitlines()[-1]);print(d['n_gpus'],d['value'],d['unit'],d['ms_per_step'],d['e2e']['value'])"
done
