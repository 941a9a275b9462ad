"""Mirror of ``/root/reference/app/src/preprocess.py`` (and ``src/predict.py:36-46``).

``preprocess_image(image)`` is the reference transform, unchanged and on the host (PIL): Grayscale -> Resize(96, 320)
-> ToTensor -> Normalize(0.5, 0.5).  ``preprocess_u8`` moves its tail to the GPU: grayscale uint8 images that are
already 96 x 320 go through ``hmocr_preprocess_u8`` (bit-identical to ToTensor + Normalize), so a batch crosses PCIe
as 30 KB per image instead of 120 KB.
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
