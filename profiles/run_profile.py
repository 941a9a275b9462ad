"""Short profiling workload: B=256 encoder + a few greedy decode steps (ncu launch lists / captures).

    python profiles/run_profile.py [--batch 256] [--max-len 12] [--iters 2]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from handwritten_math_ocr_api_b200 import FormulaRecognitionModel  # noqa: E402
from handwritten_math_ocr_api_b200.layout import ModelConfig  # noqa: E402
from handwritten_math_ocr_api_b200.synthetic import synth_images, synth_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--max-len", type=int, default=12)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--start-len", type=int, default=0, help="unused placeholder")
a = ap.parse_args()
cfg = ModelConfig()
model = FormulaRecognitionModel(cfg.vocab_size)
model.load_state_dict(synth_state_dict(cfg, seed=0, eos_bias_sigma=0.0))
imgs = synth_images(min(a.batch, 8), seed=1234).cuda()
imgs = imgs.repeat((a.batch + imgs.shape[0] - 1) // imgs.shape[0], 1, 1, 1)[: a.batch].contiguous()
for _ in range(a.iters):
    tokens, steps, _ = model.generate(imgs, max_len=a.max_len)
torch.cuda.synchronize()
print("ok", tokens.shape, steps, model.last_timings_ms())
