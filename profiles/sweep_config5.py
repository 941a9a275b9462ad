"""BASELINE.json config 5: throughput sweep, batch 1..1024 per GPU x max LaTeX length 64 / 128 / 256 (one GPU; the
job is data-parallel, N GPUs = N x these numbers minus the token gather).  max_seq_len is raised to 256 before the
model is built, so the 8-block variant of the decode kernel runs for every T.

    python profiles/sweep_config5.py [--batches 1,8,64,256,1024] [--lens 64,128,256]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from handwritten_math_ocr_api_b200 import FormulaRecognitionModel  # noqa: E402
from handwritten_math_ocr_api_b200.config import Config  # noqa: E402
from handwritten_math_ocr_api_b200.layout import ModelConfig  # noqa: E402
from handwritten_math_ocr_api_b200.synthetic import synth_images, synth_state_dict  # noqa: E402


class Config256(Config):
    max_seq_len = 256


ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="1,8,64,256,512,1024")
ap.add_argument("--lens", default="64,128,256")
a = ap.parse_args()
cfg = ModelConfig(max_seq_len=256)
m = FormulaRecognitionModel(cfg.vocab_size, config=Config256())
m.load_state_dict(synth_state_dict(cfg, seed=0, eos_bias_sigma=0.0))
base = synth_images(8, seed=1234).cuda()
print("| batch | T | encoder ms | decode ms | us/step | images/s | tokens/s |")
print("|---|---|---|---|---|---|---|")
for B in [int(x) for x in a.batches.split(",")]:
    imgs = base.repeat((B + 7) // 8, 1, 1, 1)[:B].contiguous()
    for T in [int(x) for x in a.lens.split(",")]:
        for _ in range(2):
            m.generate(imgs, max_len=T)
        torch.cuda.synchronize()
        best = None
        for _ in range(3):
            m.generate(imgs, max_len=T)
            e, d = m.last_timings_ms()
            best = (e, d) if best is None or e + d < sum(best) else best
        e, d = best
        print(f"| {B} | {T} | {e:.2f} | {d:.2f} | {d / T * 1e3:.1f} | {B / (e + d) * 1e3:.0f} | {B * T / (e + d) * 1e3:.0f} |",
              flush=True)
