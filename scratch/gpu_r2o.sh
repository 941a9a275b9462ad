#!/bin/bash
# round 2, call O: full GPU suite + smoke + bench line (with the CPU baseline) + reference arm on the final build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2o_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2o_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2o_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2o_bench_c2.json 2> gpurun_out/r2o_bench_c2.err; echo "bench c2 rc=$?"; tail -c 300 gpurun_out/r2o_bench_c2.err
python -c "import json;d=json.load(open('gpurun_out/r2o_bench_c2.json'));print(d['value'],d['e2e']['value'],d['encoder_ms'],d['decode_ms'],d['roofline']['frac'],d['roofline']['traffic'],d['cpu_baseline'])"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2o_bench_ref.json 2> gpurun_out/r2o_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/r2o_bench_ref.json | head -c 600
for c in 3 4 5; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu > gpurun_out/r2o_bench_c$c.json 2> gpurun_out/r2o_bench_c$c.err; echo "bench c$c rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2o_bench_c$c.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'])"
done
