"""Mirror of ``/root/reference/app/src/im2latex.py`` (what ``POST /predict`` calls):
``predict(model, image_tensor, vocab, idx2char, device) -> (formula, confidence)``.

The reference re-runs the whole Swin encoder for every generated token and syncs 3-4 times per
step (im2latex.py:25-45); here it is one ``generate`` call that also returns the chosen token's
log-softmax, from which the confidence is rebuilt with the reference's exact bookkeeping:
sum of ``log(p + 1e-10)`` over every step INCLUDING the eos step, divided by the number of
non-eos tokens (im2latex.py:33-39,50,55).  ``predict_batch`` is the tensor-batched form of the
``/predict/batch`` Python loop (app/src/main.py:546-550).
"""
from __future__ import annotations

import math
import re
from typing import Dict, List, Tuple

import torch

from .config import config

UNABLE = r"\text{Unable to detect a formula from the image. Please verify the model.}"


def tokens_to_latex(token_ids, idx2char) -> str:
    """app/src/utils.py:17-20."""
    specials = (config.sos_token, config.eos_token, config.pad_token)
    ids = [t for t in token_ids if t in idx2char and idx2char[t] not in specials]
    return ' '.join(idx2char[t] for t in ids)


def clean_latex_output(latex_str: str) -> str:
    """app/src/utils.py:22-27."""
    latex_str = re.sub(r'\\begin\s+\{', r'\\begin{', latex_str)
    latex_str = re.sub(r'\\end\s+\{', r'\\end{', latex_str)
    latex_str = re.sub(r'\{(\s+)([a-zA-Z]+)(\s+)\}', r'{\2}', latex_str)
    latex_str = re.sub(r'\\\s+\\', r'\\\\', latex_str)
    return latex_str


def load_model(model_path: str, vocab, device="cuda"):
    """im2latex.py:7-13 un-pickles a whole reference module; wrap it into the engine."""
    from .model_swin import FormulaRecognitionModel
    ref = torch.load(model_path, map_location="cpu", weights_only=False)
    sd = ref["model_state_dict"] if isinstance(ref, dict) else ref.state_dict()
    m = FormulaRecognitionModel(len(vocab), device=device)
    m.load_state_dict(sd)
    return m


def _finish(row_tokens: List[int], row_logp: List[float], eos: int, idx2char) -> Tuple[str, float]:
    out_tokens: List[int] = []
    lp_sum = 0.0
    for tok, lp in zip(row_tokens, row_logp):
        lp_sum += math.log(math.exp(lp) + 1e-10)      # log(softmax + 1e-10), im2latex.py:37
        if tok == eos:
            break
        out_tokens.append(tok)
    if not out_tokens:
        return UNABLE, 0.0
    formula = clean_latex_output(tokens_to_latex(out_tokens, idx2char))
    return formula, float(torch.exp(torch.tensor(lp_sum / len(out_tokens))).item())


def predict_batch(model, image_tensors: torch.Tensor, vocab: Dict[str, int], idx2char: Dict[int, str],
                  device: str = "cuda") -> List[Tuple[str, float]]:
    model = model.to(device)
    model.eval()
    tokens, steps, logp = model.generate(image_tensors, max_len=config.max_seq_len, return_logprobs=True)
    toks, lps = tokens[:, 1:].cpu().tolist(), logp.cpu().tolist()
    eos = vocab[config.eos_token]
    return [_finish(t, l, eos, idx2char) for t, l in zip(toks, lps)]


def predict(model, image_tensor: torch.Tensor, vocab: dict, idx2char: dict, device: str = "cuda"):
    """-> (cleaned_formula, confidence_score), im2latex.py:15-56."""
    return predict_batch(model, image_tensor, vocab, idx2char, device)[0]
