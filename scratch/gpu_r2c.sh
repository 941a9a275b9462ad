#!/bin/bash
# round 2, call C: state check of the restored build (GPU suite, bench c2), compute-sanitizer logs (memcheck / racecheck / synccheck)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2c_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2c_bench_c2.json 2> gpurun_out/r2c_bench_c2.err; echo "bench c2 rc=$?"; tail -c 300 gpurun_out/r2c_bench_c2.err
for tool in memcheck synccheck racecheck; do
  SAN_B=9 SAN_T=8 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python profiles/sanitize_case.py > gpurun_out/r2c_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"; tail -4 gpurun_out/r2c_sanitizer_$tool.log
done
