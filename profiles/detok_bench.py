"""N3 measurement: ids -> strings for one batch (B=256, T=150): the reference's per-id Python loop
(src/inference.py:29-40, here on ids already copied to the host) against the device detokeniser + one join per row."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from handwritten_math_ocr_api_b200 import FormulaRecognitionModel
from handwritten_math_ocr_api_b200.inference import ids_to_strings, packed_to_strings
from handwritten_math_ocr_api_b200.layout import ModelConfig
from handwritten_math_ocr_api_b200.synthetic import synth_state_dict

cfg = ModelConfig()
m = FormulaRecognitionModel(cfg.vocab_size)
m.load_state_dict(synth_state_dict(cfg, seed=0, eos_bias_sigma=0.0))
idx2char = {i: f"t{i}" for i in range(cfg.vocab_size)}
idx2char[m.pad_id], idx2char[m.sos_id], idx2char[m.eos_id] = "<pad>", "<sos>", "<eos>"
g = torch.Generator().manual_seed(0)
tok = torch.randint(3, cfg.vocab_size, (256, 151), generator=g)
tok[:, 0] = m.sos_id
for b in range(256):
    tok[b, 40 + (b * 7) % 110] = m.eos_id
dev = tok.cuda()
for _ in range(3):
    m.pack_tokens(dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    a = ids_to_strings(dev.cpu().tolist(), idx2char)
t_ref = (time.perf_counter() - t0) / 10
t0 = time.perf_counter()
for _ in range(10):
    p, n = m.pack_tokens(dev)
    b = packed_to_strings(p.cpu().tolist(), n.cpu().tolist(), idx2char)
t_new = (time.perf_counter() - t0) / 10
# the reference as written also pays one .item() sync per id (src/inference.py:33 on a CUDA tensor)
t0 = time.perf_counter()
c = []
for seq in dev[:16]:
    toks = []
    for idx in seq:
        t = idx2char[idx.item()]
        if t in ("<sos>", "<pad>"):
            continue
        if t == "<eos>":
            break
        toks.append(t)
    c.append(' '.join(toks))
t_item = (time.perf_counter() - t0) / 16 * 256
print(f"B=256 T=150: reference loop on host ids {t_ref * 1e3:.2f} ms, as written (.item() per id, extrapolated from 16 rows) "
      f"{t_item * 1e3:.1f} ms, device detokeniser + join {t_new * 1e3:.2f} ms; identical: {a == b and c == a[:16]}")
