"""Decode time of the B=256 x T=150 workload against steps_per_launch (launch boundaries cost a cluster drain + refill)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from handwritten_math_ocr_api_b200 import FormulaRecognitionModel
from handwritten_math_ocr_api_b200.layout import ModelConfig
from handwritten_math_ocr_api_b200.synthetic import synth_images, synth_state_dict
cfg = ModelConfig()
m = FormulaRecognitionModel(cfg.vocab_size)
m.load_state_dict(synth_state_dict(cfg, seed=0, eos_bias_sigma=0.0))
imgs = synth_images(8, seed=1234).cuda().repeat(32, 1, 1, 1).contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for spl in (16, 25, 32, 50, 75, 150, 16):
    m.set_option("steps_per_launch", spl)
    for _ in range(2):
        m.generate(imgs, max_len=150)
    ts = []
    for _ in range(6):
        flush.zero_()
        m.generate(imgs, max_len=150)
        torch.cuda.synchronize()
        ts.append(m.last_timings_ms())
    enc = sum(t[0] for t in ts) / len(ts); dec = sum(t[1] for t in ts) / len(ts)
    print(f"steps_per_launch {spl:4d}: encoder {enc:.3f} ms  decode {dec:.3f} ms", flush=True)
