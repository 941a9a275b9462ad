"""Small workload for compute-sanitizer (one tool per gpurun call, B200_PROFILING.md):

    compute-sanitizer --tool memcheck|racecheck|synccheck python profiles/sanitize_case.py

Covers every hand-synchronised kernel on the smallest shapes that still exercise its protocol: the encoder of 9 images
(patch embed, LayerNorm, the tcgen05 GEMM with the fp16 / GELU / residual epilogues, the CTA-pair GEMM, window attention
with and without shift, patch merging), the teacher-forced decoder (LayerNorm-fused epilogue, prefill attention), the
persistent cluster decode kernel greedy (9 rows = one full + one partial cluster, 8 steps in two launches) and beam 3.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from handwritten_math_ocr_api_b200 import FormulaRecognitionModel
from handwritten_math_ocr_api_b200.layout import ModelConfig
from handwritten_math_ocr_api_b200.synthetic import synth_images, synth_state_dict

B = int(os.environ.get("SAN_B", "9"))
T = int(os.environ.get("SAN_T", "8"))
cfg = ModelConfig()
m = FormulaRecognitionModel(cfg.vocab_size)
m.load_state_dict(synth_state_dict(cfg, seed=0))
imgs = synth_images(B, seed=1234).cuda()
enc = m.encoder(imgs)
torch.cuda.synchronize()
print("encoder ok", tuple(enc.shape), float(enc.abs().max()))
tgt = torch.randint(4, cfg.vocab_size, (2, 6), device="cuda")
lg = m.decoder(enc[:2], tgt)
torch.cuda.synchronize()
print("decoder ok", tuple(lg.shape))
m.set_option("steps_per_launch", 5)
tok, steps, lp = m.generate(encoder_out=enc, max_len=T, return_logprobs=True)
torch.cuda.synchronize()
print("greedy ok", tuple(tok.shape), steps)
tok, steps, _, sc = m.generate(encoder_out=enc, max_len=T, beam_size=3)
torch.cuda.synchronize()
print("beam ok", tuple(tok.shape), steps, [round(x, 3) for x in sc.tolist()])
