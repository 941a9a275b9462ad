#!/bin/bash
# round 2, call 3d: the two test edits (multi-wave bound, early exit across clusters)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_parity_gpu.py -m gpu -q -x -p no:cacheprovider -k "stops_on_the_device or early_exit" > gpurun_out/r3d_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r3d_pytest.log
