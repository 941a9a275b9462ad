import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from handwritten_math_ocr_api_b200 import FormulaRecognitionModel
from handwritten_math_ocr_api_b200.layout import ModelConfig
from handwritten_math_ocr_api_b200.synthetic import synth_images, synth_state_dict
cfg = ModelConfig()
m = FormulaRecognitionModel(cfg.vocab_size)
m.load_state_dict(synth_state_dict(cfg, seed=0))
N = int(sys.argv[1]); NIMG = int(sys.argv[2])
imgs = synth_images(8, seed=99).cuda().repeat(9, 1, 1, 1)[:NIMG].contiguous()
feats = m.encoder(imgs)
per_image = [collections.Counter() for _ in range(NIMG)]
for rep in range(N):
    tok, n, _, sc = m.generate(encoder_out=feats, max_len=70, beam_size=5)
    for i, v in enumerate(sc.tolist()):
        per_image[i][v] += 1
multi = [(i, dict(c)) for i, c in enumerate(per_image) if len(c) > 1]
print("images with more than one score value:", len(multi), "of", NIMG)
for i, c in multi[:12]:
    print(i, c)
