"""Evaluation harness: a working version of ``/root/reference/src/test_model.py``.

The reference script cannot run as written (it passes ``mode=`` to a ``predict`` that has no such parameter,
``src/test_model.py:69`` vs ``src/inference.py:7``, and labels every row of a batch with the image ids of the first
batch, ``:79``).  The METRICS are kept exactly as the reference computes them (``calculate_metrics``, ``:47-58``):

* ``is_correct`` - exact string match of the space-joined token sequence;
* ``cer`` - ``1 - difflib.SequenceMatcher(None, pred, truth).ratio()`` over characters (the Levenshtein value the
  function computes first is always overwritten by this one).

Ground truth of a sample = ``captions[i][1 : lengths[i] - 1]`` mapped through ``idx2char`` and joined with spaces
(``:74-75``).  Predictions come from ``inference.predict`` = ONE ``generate`` call per batch.
"""
from __future__ import annotations

from difflib import SequenceMatcher
from typing import Dict, Iterable, List, Optional, Tuple

from . import inference
from .config import config as _default_config


def calculate_metrics(pred: str, truth: str) -> Tuple[bool, float]:
    """``src/test_model.py:47-58``."""
    cer = 1 - SequenceMatcher(None, pred, truth).ratio()
    return pred == truth, cer


def truth_string(caption, length: int, idx2char: Dict[int, str]) -> str:
    """``src/test_model.py:74-75``: drop the leading sos and the trailing eos of the padded caption."""
    return ' '.join(idx2char[int(i)] for i in caption[1:int(length) - 1])


def summarize(rows: List[dict]) -> dict:
    """``src/test_model.py:96-100``."""
    n = len(rows)
    return {
        'accuracy': sum(r['is_correct'] for r in rows) / n if n else 0.0,
        'avg_cer': sum(r['cer'] for r in rows) / n if n else 0.0,
        'total_samples': n,
    }


def evaluate_model(model, batches: Iterable, vocab, idx2char, device=None, use_beam: bool = False,
                   beam_size: Optional[int] = None, image_ids: Optional[Iterable] = None, config=_default_config):
    """``evaluate_model`` (``src/test_model.py:60-87``) over an iterable of ``(images, captions, lengths)`` batches
    (what ``data_loader.get_test_loader`` yields).  Returns ``(rows, summary)``; ``rows`` are dicts with the
    reference's columns (``image_id`` is the running sample index unless ``image_ids`` supplies one per sample)."""
    model.eval()
    ids = iter(image_ids) if image_ids is not None else None
    rows: List[dict] = []
    for images, captions, lengths in batches:
        preds = inference.predict(images, model, vocab, idx2char, device,
                                  beam_size=beam_size if beam_size is not None else config.beam_size,
                                  config=config, use_beam=use_beam)
        caps = captions.tolist() if hasattr(captions, "tolist") else captions
        lens = lengths.tolist() if hasattr(lengths, "tolist") else lengths
        for i, pred in enumerate(preds):
            truth = truth_string(caps[i], lens[i], idx2char)
            correct, cer = calculate_metrics(pred, truth)
            rows.append({
                'image_id': next(ids) if ids is not None else len(rows),
                'prediction': pred,
                'ground_truth': truth,
                'is_correct': correct,
                'cer': cer,
            })
    return rows, summarize(rows)
