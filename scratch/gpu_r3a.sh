#!/bin/bash
# round 2, call 3a: device-side early exit of the persistent decode kernel: model + parity tests, bench, early-exit latency probe
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_model_gpu.py tests/test_parity_gpu.py tests/test_res18_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r3a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r3a_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r3a_bench_c2.json 2> gpurun_out/r3a_bench_c2.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r3a_bench_c2.err
python -c "import json;d=json.load(open('gpurun_out/r3a_bench_c2.json'));print(d['value'],d['e2e']['value'],d['encoder_ms'],d['decode_ms'],d['p50_ms_per_image_b1'])"
python - <<'PY' > gpurun_out/r3a_early_exit_latency.txt 2>&1
import torch, time, sys
sys.path.insert(0,'.')
from handwritten_math_ocr_api_b200 import FormulaRecognitionModel
from handwritten_math_ocr_api_b200.layout import ModelConfig
from handwritten_math_ocr_api_b200.synthetic import synth_images, synth_state_dict
cfg=ModelConfig(); m=FormulaRecognitionModel(cfg.vocab_size); m.load_state_dict(synth_state_dict(cfg, seed=0))
imgs=synth_images(64, seed=7).cuda()
for B in (1, 8, 64):
    x=imgs[:B]
    for _ in range(3): tok,steps,_=m.generate(x, max_len=150)
    ts=[]
    for _ in range(20):
        torch.cuda.synchronize(); t0=time.perf_counter(); tok,steps,_=m.generate(x, max_len=150); torch.cuda.synchronize(); ts.append(time.perf_counter()-t0)
    ts.sort()
    print(f"B={B}: steps {steps}, steps run on the device {m.last_decode_steps()}, p50 wall {ts[len(ts)//2]*1e3:.3f} ms, timings {m.last_timings_ms()}")
PY
cat gpurun_out/r3a_early_exit_latency.txt
