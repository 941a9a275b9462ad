import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from handwritten_math_ocr_api_b200 import FormulaRecognitionModel
from handwritten_math_ocr_api_b200.layout import ModelConfig
from handwritten_math_ocr_api_b200.synthetic import synth_images, synth_state_dict
cfg = ModelConfig()
m = FormulaRecognitionModel(cfg.vocab_size)
m.load_state_dict(synth_state_dict(cfg, seed=0))
NIMG = int(sys.argv[2]) if len(sys.argv) > 2 else 45
imgs = synth_images(8, seed=99).cuda().repeat(9, 1, 1, 1)[:NIMG].contiguous()
feats = m.encoder(imgs)
ref = None
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
    out = m.generate(encoder_out=feats, max_len=70, beam_size=5)
    tok, n, _, sc = out
    if ref is None:
        ref = (tok.clone(), sc.clone(), n)
    else:
        d = (sc - ref[1]).abs()
        same_tok = torch.equal(tok, ref[0])
        if d.max() > 0 or not same_tok or n != ref[2]:
            bad = (d > 0).nonzero().flatten().tolist()
            print(f"rep {rep}: steps {n} vs {ref[2]}, tokens equal {same_tok}, score diffs at images {bad}: {[(float(sc[i]), float(ref[1][i])) for i in bad[:6]]}")
print("done")
