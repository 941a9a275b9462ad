"""Names the reference's pickled ``model.pth`` refers to (``/root/reference/app/src/im2latex.py:11`` un-pickles a WHOLE
``model_swin.FormulaRecognitionModel``; pickle resolves classes by module path).  These are skeletons: they only have
to exist and be ``nn.Module``s so that ``torch.load`` can rebuild the object tree and ``state_dict()`` can walk it;
the arithmetic runs in the engine (``handwritten_math_ocr_api_b200.model_swin``), which ``im2latex.load_model`` builds
from that state dict.
"""
import torch.nn as nn

from handwritten_math_ocr_api_b200.model_swin import FormulaRecognitionModel as EngineModel  # noqa: F401


class EncoderSwin(nn.Module):            # /root/reference/app/src/model_swin.py:13
    pass


class DecoderTransformer(nn.Module):     # /root/reference/app/src/model_swin.py:49
    pass


class FormulaRecognitionModel(nn.Module):    # /root/reference/app/src/model_swin.py:91
    pass
