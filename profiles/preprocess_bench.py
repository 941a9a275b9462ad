"""N2 measurement: the reference's preprocessing (PIL + torchvision on the host) against hmocr_preprocess_image_u8.

    python profiles/preprocess_bench.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from PIL import Image

from handwritten_math_ocr_api_b200 import FormulaRecognitionModel
from handwritten_math_ocr_api_b200.layout import ModelConfig
from handwritten_math_ocr_api_b200.preprocess import preprocess_batch_gpu, preprocess_image
from handwritten_math_ocr_api_b200.synthetic import synth_state_dict

cfg = ModelConfig()
m = FormulaRecognitionModel(cfg.vocab_size)
m.load_state_dict(synth_state_dict(cfg, seed=0, eos_bias_sigma=0.0))
rng = np.random.default_rng(0)
for (h, w, mode) in [(96, 320, "L"), (200, 800, "RGB"), (480, 1600, "RGB"), (1200, 4000, "RGB")]:
    imgs = [Image.fromarray(rng.integers(0, 256, (h, w) if mode == "L" else (h, w, 3), dtype=np.uint8), mode=mode)
            for _ in range(32)]
    t0 = time.perf_counter()
    ref = torch.cat([preprocess_image(i) for i in imgs])
    t_cpu = (time.perf_counter() - t0) / len(imgs)
    preprocess_batch_gpu(m, imgs[:4]); torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = preprocess_batch_gpu(m, imgs); torch.cuda.synchronize()
    t_gpu = (time.perf_counter() - t0) / len(imgs)
    arrs = [np.ascontiguousarray(np.asarray(i)) for i in imgs]
    t0 = time.perf_counter()
    got2 = preprocess_batch_gpu(m, arrs); torch.cuda.synchronize()
    t_gpu_arr = (time.perf_counter() - t0) / len(imgs)
    print(f"{h}x{w} {mode}: host PIL+torchvision {t_cpu * 1e3:.3f} ms/image, GPU (from PIL images, incl. H2D) {t_gpu * 1e3:.3f} ms, "
          f"GPU (from uint8 arrays) {t_gpu_arr * 1e3:.3f} ms, identical: {torch.equal(got.cpu(), ref) and torch.equal(got2.cpu(), ref)}")
