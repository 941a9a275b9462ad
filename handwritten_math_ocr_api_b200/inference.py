"""Mirror of ``/root/reference/src/inference.py``: same ``predict`` signature and results, but the
150-iteration Python loop (encoder once + full-prefix decoder call per step + a D2H sync per step)
is ONE ``model.generate`` call; ids come back in a single D2H copy.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from .config import config as _default_config


def ids_to_strings(sequences, idx2char: Dict[int, str], config=_default_config) -> List[str]:
    """src/inference.py:29-40: skip sos/pad, stop at eos, join with spaces."""
    skip = (config.sos_token, config.pad_token)
    results = []
    for seq in sequences:
        tokens = []
        for idx in seq:
            token = idx2char[int(idx)]
            if token in skip:
                continue
            if token == config.eos_token:
                break
            tokens.append(token)
        results.append(' '.join(tokens))
    return results


def packed_to_strings(packed, lengths, idx2char: Dict[int, str]) -> List[str]:
    """Strings from the device-side detokeniser (``model.pack_tokens``): one join per sequence."""
    return [' '.join(idx2char[i] for i in row[:n]) for row, n in zip(packed, lengths)]


def predict(images, model, vocab, idx2char, device=None, beam_size=3, config=_default_config, use_beam=False):
    """``predict(images, model, vocab, idx2char, device, beam_size=3)`` (src/inference.py:7).

    The reference ignores ``beam_size`` ("Using greedy search for simplicity", :6) - so does this
    function unless ``use_beam=True``.
    """
    model.eval()
    if device is not None and torch.device(device).type != "cuda":
        raise RuntimeError("the B200 engine has no CPU path (device must be 'cuda')")
    if vocab[config.sos_token] != model.sos_id or vocab[config.eos_token] != model.eos_id:
        raise ValueError("vocab special-token ids differ from the ids the engine was built with")
    if vocab[config.pad_token] != model.pad_id:
        raise ValueError("vocab pad id differs from the id the engine was built with")
    out = model.generate(images, max_len=config.max_seq_len, beam_size=beam_size if use_beam else 1)
    packed, lengths = model.pack_tokens(out[0])            # skip sos / pad, stop at eos: on the device
    return packed_to_strings(packed.cpu().tolist(), lengths.cpu().tolist(), idx2char)
