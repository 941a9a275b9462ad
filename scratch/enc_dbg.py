import sys, os
sys.path.insert(0, '/root/repo')
import torch
from handwritten_math_ocr_api_b200 import FormulaRecognitionModel
from handwritten_math_ocr_api_b200.layout import ModelConfig
from handwritten_math_ocr_api_b200.synthetic import synth_images, synth_state_dict
cfg = ModelConfig()
m = FormulaRecognitionModel(cfg.vocab_size)
m.load_state_dict(synth_state_dict(cfg, seed=0, eos_bias_sigma=0.0))
imgs = synth_images(8, seed=1234).cuda().repeat(32, 1, 1, 1).contiguous()
for dbg in (0, 3, 7):
    m.set_option("gemm_dbg", dbg)
    for _ in range(3): m.encoder(imgs)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): m.encoder(imgs)
    b.record(); torch.cuda.synchronize()
    print("gemm_dbg", dbg, "encoder ms", a.elapsed_time(b) / 5)
