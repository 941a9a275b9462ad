"""One-off: implicit-GEMM convolutions vs the im2col path at the bench batch (256) and an odd batch (37)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from handwritten_math_ocr_api_b200.model_res18trans import FormulaRecognitionModel
from handwritten_math_ocr_api_b200.layout import ModelConfig
from handwritten_math_ocr_api_b200.synthetic import synth_images, synth_state_dict_res18, synth_pos_table

cfg = ModelConfig()
m = FormulaRecognitionModel(cfg.vocab_size)
m.load_state_dict(synth_state_dict_res18(cfg, seed=0))
pos = synth_pos_table(cfg.d_model, seed=3)
for B in (256, 37):
    imgs = synth_images(B, seed=5).cuda()
    m.set_option("conv_impl", 1); a = m.encoder(imgs, pos).clone()
    m.set_option("conv_impl", 0); b = m.encoder(imgs, pos)
    print("B", B, "max |implicit - im2col|", (a - b).abs().max().item(), "feature std", a.std().item())
