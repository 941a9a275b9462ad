"""Mirror of ``/root/reference/app/src/preprocess.py`` (and ``src/predict.py:36-46``).

``preprocess_image(image)`` is the reference transform, unchanged and on the host (PIL): Grayscale -> Resize(96, 320)
-> ToTensor -> Normalize(0.5, 0.5).  ``preprocess_image_gpu`` / ``preprocess_batch_gpu`` run the SAME transform on the
GPU, bit for bit (``hmocr_preprocess_image_u8``: Pillow's integer luma, its two-pass antialiased bilinear resample with
22-bit fixed-point coefficients, torchvision's ToTensor + Normalize), for images of any size in mode "L" or "RGB".
``preprocess_u8`` is the tail alone: grayscale uint8 images that are already 96 x 320 go through
``hmocr_preprocess_u8``, so a batch crosses PCIe as 30 KB per image instead of 120 KB.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .config import config


def preprocess_image(image) -> torch.Tensor:
    """app/src/preprocess.py:6-16: PIL image -> f32 [1,1,96,320] in [-1,1] (host side)."""
    from torchvision import transforms
    transform = transforms.Compose([
        transforms.Grayscale(num_output_channels=1),
        transforms.Resize((config.img_h, config.img_w)),
        transforms.ToTensor(),
        transforms.Normalize(mean=[0.5], std=[0.5]),
    ])
    return transform(image).unsqueeze(0)


def preprocess_u8(model, images_u8: torch.Tensor) -> torch.Tensor:
    """uint8 [B,96,320] (host or device) -> f32 [B,1,96,320] on the model's device."""
    if images_u8.dtype != torch.uint8 or images_u8.dim() != 3 or tuple(images_u8.shape[1:]) != (config.img_h, config.img_w):
        raise ValueError(f"images_u8 must be uint8 [B,{config.img_h},{config.img_w}], got {images_u8.dtype} {tuple(images_u8.shape)}")
    x = images_u8.to(model.device).contiguous()
    out = torch.empty(x.shape[0], 1, config.img_h, config.img_w, dtype=torch.float32, device=model.device)
    with torch.cuda.device(model.device):
        _lib.check(model._eng.lib.hmocr_preprocess_u8(model._handle(), C.c_void_p(x.data_ptr()), x.shape[0],
                                                      C.c_void_p(out.data_ptr()),
                                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "hmocr_preprocess_u8")
    return out


def _as_u8_array(image):
    """PIL image (any mode; everything but "L" goes through ``convert("RGB")`` like ``src/predict.py:45``) or a
    numpy / torch uint8 array ``[H, W]`` / ``[H, W, 3]`` -> contiguous numpy uint8 array."""
    import numpy as np
    if hasattr(image, "mode") and hasattr(image, "convert"):          # PIL.Image
        if image.mode not in ("L", "RGB"):
            image = image.convert("RGB")
        arr = np.asarray(image)
    elif isinstance(image, torch.Tensor):
        arr = image.cpu().numpy()
    else:
        arr = np.asarray(image)
    if arr.dtype != np.uint8 or arr.ndim not in (2, 3) or (arr.ndim == 3 and arr.shape[2] != 3):
        raise ValueError(f"image must be uint8 [H,W] or [H,W,3], got {arr.dtype} {arr.shape}")
    return np.ascontiguousarray(arr)


def preprocess_batch_gpu(model, images) -> torch.Tensor:
    """The reference transform of every image (sizes may differ) on the GPU -> f32 ``[B,1,96,320]`` on the device,
    bit-identical to ``torch.cat([preprocess_image(i) for i in images])``."""
    arrs = [_as_u8_array(i) for i in images]
    out = torch.empty(len(arrs), 1, config.img_h, config.img_w, dtype=torch.float32, device=model.device)
    with torch.cuda.device(model.device):
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)      # the current stream OF the model's device
        for b, a in enumerate(arrs):
            _lib.check(model._eng.lib.hmocr_preprocess_image_u8(model._handle(), C.c_void_p(a.ctypes.data),
                                                                1 if a.ndim == 2 else 3, a.shape[0], a.shape[1],
                                                                C.c_void_p(out[b].data_ptr()), stream),
                       "hmocr_preprocess_image_u8")
    return out


def preprocess_image_gpu(model, image) -> torch.Tensor:
    """``preprocess_image`` (app/src/preprocess.py:6-16) on the GPU: one image -> f32 ``[1,1,96,320]`` (device)."""
    return preprocess_batch_gpu(model, [image])


def preprocess_dataloader_gpu(model, grays) -> torch.Tensor:
    """The loader's route (``src/data_loader.py:31-35``: ``cv2.resize(img, (320, 96))`` + ToTensor + Normalize) on the
    GPU for a list of grayscale uint8 arrays of any size -> f32 ``[B,1,96,320]`` (device), bit-identical to OpenCV."""
    arrs = [_as_u8_array(g) for g in grays]
    if any(a.ndim != 2 for a in arrs):
        raise ValueError("the loader reads images with cv2.IMREAD_GRAYSCALE: uint8 [H,W] arrays expected")
    out = torch.empty(len(arrs), 1, config.img_h, config.img_w, dtype=torch.float32, device=model.device)
    with torch.cuda.device(model.device):
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)      # the current stream OF the model's device
        for b, a in enumerate(arrs):
            _lib.check(model._eng.lib.hmocr_preprocess_cv2_u8(model._handle(), C.c_void_p(a.ctypes.data), a.shape[0],
                                                              a.shape[1], C.c_void_p(out[b].data_ptr()), stream),
                       "hmocr_preprocess_cv2_u8")
    return out
