#!/bin/bash
# round 2, call 3c: last check of the committed tree: smoke, bench (default flags), reference arm
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3c_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r3c_smoke.log
timeout 900 python bench.py > gpurun_out/r3c_bench_default.json 2> gpurun_out/r3c_bench_default.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r3c_bench_default.err
python -c "import json;d=json.load(open('gpurun_out/r3c_bench_default.json'));print(d['value'],d['e2e']['value'],d['steps'],d['roofline']['frac'],d['roofline']['traffic'],d['roofline']['kernel'][:90],d['gpu_launches'],d['clocks'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r3c_bench_ref.json 2>/dev/null; echo "ref rc=$?"; head -c 300 gpurun_out/r3c_bench_ref.json
