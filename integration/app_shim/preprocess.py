"""``app/src/preprocess.py``: ``preprocess_image(image) -> f32 [1,1,96,320]``.  Before a model is loaded (or when the
model object cannot run kernels) this is the reference's own PIL / torchvision transform; afterwards the same
tensor, bit for bit, is computed on the GPU (Pillow's integer luma + antialiased bilinear resample + ToTensor +
Normalize in ``hmocr_preprocess_image_u8``) and stays on the device for ``predict``."""
from handwritten_math_ocr_api_b200 import preprocess as _pp

_model = None


def bind_model(model) -> None:
    global _model
    _model = model if hasattr(model, "_eng") else None


def preprocess_image(image):
    if _model is not None:
        return _pp.preprocess_image_gpu(_model, image)
    return _pp.preprocess_image(image)
