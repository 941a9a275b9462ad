"""Beam-search throughput (BASELINE.json config 3: beam 5, 64 images per GPU, T = 150).

    python profiles/beam_bench.py [--batch 64] [--beam 5] [--max-len 150]
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
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--beam", type=int, default=5)
ap.add_argument("--max-len", type=int, default=150)
a = ap.parse_args()
cfg = ModelConfig()
m = FormulaRecognitionModel(cfg.vocab_size)
m.load_state_dict(synth_state_dict(cfg, seed=0, eos_bias_sigma=0.0))
base = synth_images(8, seed=1234).cuda()
imgs = base.repeat((a.batch + 7) // 8, 1, 1, 1)[: a.batch].contiguous()
for beam in sorted({1, a.beam}):
    for _ in range(2):
        m.generate(imgs, max_len=a.max_len, beam_size=beam)
    torch.cuda.synchronize()
    best = None
    for _ in range(3):
        m.generate(imgs, max_len=a.max_len, beam_size=beam)
        e, d = m.last_timings_ms()
        best = (e, d) if best is None or d < best[1] else best
    e, d = best
    print(f"B={a.batch} beam={beam}: encoder {e:.3f} ms, decode {d:.3f} ms ({d / a.max_len * 1e3:.1f} us/step), "
          f"{a.batch / (e + d) * 1e3:.1f} images/s, {a.batch * a.max_len / (e + d) * 1e3:.0f} emitted tokens/s, "
          f"{a.batch * beam * a.max_len / (e + d) * 1e3:.0f} hypothesis tokens/s", flush=True)
