import os, sys
sys.path.insert(0, '/root/repo')
import torch
from handwritten_math_ocr_api_b200 import FormulaRecognitionModel
from handwritten_math_ocr_api_b200.layout import ModelConfig
from handwritten_math_ocr_api_b200.synthetic import synth_images, synth_state_dict
cfg = ModelConfig()
m = FormulaRecognitionModel(cfg.vocab_size)
m.load_state_dict(synth_state_dict(cfg, seed=0))
imgs = synth_images(3, seed=1234).cuda()
tok, steps, _ = m.generate(imgs, max_len=int(os.environ.get("T", "6")))
print("greedy ok", tok.shape, steps)
tok, steps, _, sc = m.generate(imgs, max_len=int(os.environ.get("T", "6")), beam_size=3)
print("beam ok", tok.shape, steps, sc.tolist())
