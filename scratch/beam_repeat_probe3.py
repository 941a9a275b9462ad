import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from handwritten_math_ocr_api_b200 import FormulaRecognitionModel
from handwritten_math_ocr_api_b200.layout import ModelConfig
from handwritten_math_ocr_api_b200.synthetic import synth_images, synth_state_dict
cfg = ModelConfig()
m = FormulaRecognitionModel(cfg.vocab_size)
m.load_state_dict(synth_state_dict(cfg, seed=0))
N = int(sys.argv[1]); NIMG = int(sys.argv[2]); BEAM = int(sys.argv[3]); SPL = int(sys.argv[4]) if len(sys.argv) > 4 else 0
GREEDY = BEAM == 0
if GREEDY:
    BEAM = 1
elif BEAM == 1:
    m.set_option("force_beam_kernel", 1)
if SPL:
    m.set_option("steps_per_launch", SPL)
FLAGS = int(sys.argv[5]) if len(sys.argv) > 5 else 0
if FLAGS:
    m.set_option("dbg_flags", FLAGS)
imgs = synth_images(8, seed=99).cuda().repeat((NIMG + 7) // 8, 1, 1, 1)[:NIMG].contiguous()
feats = m.encoder(imgs)
per_image = [collections.Counter() for _ in range(NIMG)]
for rep in range(N):
    out = m.generate(encoder_out=feats, max_len=70, beam_size=BEAM, return_logprobs=(BEAM == 1))
    sc = out[3] if BEAM > 1 else out[2].sum(1)
    for i, v in enumerate(sc.tolist()):
        per_image[i][v] += 1
multi = [(i, dict(c)) for i, c in enumerate(per_image) if len(c) > 1]
print(f"{'greedy kernel ' if GREEDY else ''}beam {BEAM} images {NIMG} spl {SPL} flags {FLAGS}: images with more than one score value: {len(multi)} of {NIMG}; deviating runs: {sum(sum(c.values()) - max(c.values()) for _, c in [(i, per_image[i]) for i in range(NIMG)])}")
