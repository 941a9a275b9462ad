"""B200-native (sm_100a) engine for the image-to-LaTeX hot path of
PTD504/handwritten-math-ocr-api: Swin-T encoder + 8-layer Transformer decoder, greedy decode.

Host side mirrors the reference's Python surface (``model_swin``, ``inference``, ``predict``,
``im2latex``); the arithmetic runs in hand-written CUDA kernels behind the C ABI of
``include/hmocr.h`` (``libhmocr.so``, built in-tree by ``build.py``).  No CPU fallback.
"""
from .config import Config, config  # noqa: F401
from .model_swin import DecoderTransformer, EncoderSwin, FormulaRecognitionModel  # noqa: F401

__all__ = ["FormulaRecognitionModel", "EncoderSwin", "DecoderTransformer", "Config", "config"]
