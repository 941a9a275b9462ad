"""Mirror of ``/root/reference/src/predict.py`` (the B=1 script): ``preprocess_image`` and
``predict(image_tensor, max_len=150)``.

The reference binds ``model`` / ``vocab`` / ``idx2token`` as module globals loaded at import time
(predict.py:13-29); here they are bound explicitly with ``bind``.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .config import config

model = None
vocab: Optional[Dict[str, int]] = None
idx2token: Optional[Dict[int, str]] = None


def bind(model_, vocab_, idx2token_):
    global model, vocab, idx2token
    model, vocab, idx2token = model_, vocab_, idx2token_


def preprocess_image(image_path):
    """predict.py:36-46: Grayscale -> Resize(96,320) -> ToTensor -> Normalize(0.5,0.5) (host side, PIL)."""
    from PIL import Image
    from torchvision import transforms
    transform = transforms.Compose([
        transforms.Grayscale(num_output_channels=1),
        transforms.Resize((config.img_h, config.img_w)),
        transforms.ToTensor(),
        transforms.Normalize(mean=[0.5], std=[0.5]),
    ])
    return transform(Image.open(image_path).convert("RGB")).unsqueeze(0)


def predict(image_tensor: torch.Tensor, max_len: int = 150) -> List[str]:
    """predict.py:49-67.  Returns ``[idx2token[i] for i in output_seq[1:-1]]`` - note the reference
    drops the last token even when the loop ended by length rather than eos."""
    if model is None:
        raise RuntimeError("call handwritten_math_ocr_api_b200.predict.bind(model, vocab, idx2token) first")
    if image_tensor.shape[0] != 1:
        raise ValueError("predict.py decodes one image at a time")
    eos = vocab[config.eos_token]
    tokens, steps, _ = model.generate(image_tensor, max_len=max_len)
    seq = tokens[0].cpu().tolist()
    # the B=1 loop breaks at the first eos
    if eos in seq[1:]:
        seq = seq[: seq.index(eos, 1) + 1]
    return [idx2token[i] for i in seq[1:-1]]
