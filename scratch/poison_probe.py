import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from handwritten_math_ocr_api_b200 import FormulaRecognitionModel
from handwritten_math_ocr_api_b200.layout import ModelConfig
from handwritten_math_ocr_api_b200.synthetic import synth_images, synth_state_dict
cfg = ModelConfig()
m = FormulaRecognitionModel(cfg.vocab_size)
m.load_state_dict(synth_state_dict(cfg, seed=0))
imgs = synth_images(8, seed=99).cuda().repeat(32, 1, 1, 1).contiguous()
feats = m.encoder(imgs)
nan_runs = 0; dev = 0; ref = None
for rep in range(int(sys.argv[1])):
    tok, n, lp = m.generate(encoder_out=feats, max_len=70, return_logprobs=True)
    sc = lp.sum(1)
    if torch.isnan(sc).any():
        nan_runs += 1
        if nan_runs <= 3: print("rep", rep, "NaN scores at images", torch.isnan(sc).nonzero().flatten().tolist()[:10])
    if ref is None: ref = sc.clone()
    else: dev += int((sc != ref).sum())
print("runs with NaN:", nan_runs, "deviating pairs:", dev)
