"""One-off robustness sweep: greedy tokens / log-probs of an image must not depend on the batch it is decoded in
(rows of a cluster are independent; padding rows, partial clusters and multi-wave launches included)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from handwritten_math_ocr_api_b200 import FormulaRecognitionModel
from handwritten_math_ocr_api_b200.layout import ModelConfig
from handwritten_math_ocr_api_b200.synthetic import synth_images, synth_state_dict

cfg = ModelConfig()
m = FormulaRecognitionModel(cfg.vocab_size)
m.load_state_dict(synth_state_dict(cfg, seed=0))
imgs = synth_images(310, seed=99).cuda()
ref_tok, ref_steps, ref_lp = m.generate(imgs[:8], max_len=150, return_logprobs=True)
full = {}
bad = 0
for B in (1, 2, 3, 5, 7, 8, 9, 15, 16, 17, 31, 33, 63, 65, 100, 255, 256, 257, 264, 265, 300, 310):
    for L in (1, 2, 16, 17, 33, 150):
        tok, steps, lp = m.generate(imgs[:B], max_len=L, return_logprobs=True)
        n = min(B, 8)
        cols = min(tok.shape[1], ref_tok.shape[1])
        # a shorter max_len is a prefix of the longer decode (until early exit pads)
        a, b = tok[:n, :cols], ref_tok[:n, :cols]
        same = True
        for r in range(n):
            ra, rb = a[r].tolist(), b[r].tolist()
            # compare up to the first eos of the reference row
            if m.eos_id in rb:
                k = rb.index(m.eos_id) + 1
                ra, rb = ra[:k], rb[:k]
            same &= ra == rb
        if not same:
            bad += 1
            print("MISMATCH B", B, "L", L)
print("fuzz done, mismatching (B, L) combinations:", bad)
