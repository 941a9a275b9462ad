"""TEST INFRASTRUCTURE - CPU restatement of the reference's image preprocessing (SURVEY.md §8 f, row N2).

    /root/reference/app/src/preprocess.py:6-16 and /root/reference/src/predict.py:36-46:
        Grayscale(1) -> Resize((96, 320)) -> ToTensor() -> Normalize([0.5], [0.5])        on a PIL image

The arithmetic lives in two third-party dependencies that are not under /root/reference (requirements.txt pins
neither; the versions in this image are Pillow 12.2.0 and torchvision 0.26.0):

* ``Grayscale`` = ``PIL.Image.convert("L")``: for RGB input ``L = (R*19595 + G*38470 + B*7471 + 0x8000) >> 16``
  (Pillow ``src/libImaging/Convert.c``, ``rgb2l`` / ``L24``); mode "L" input is passed through.
* ``Resize`` on a PIL image = ``Image.resize((w, h), BILINEAR)`` = ``ImagingResample`` (Pillow
  ``src/libImaging/Resample.c``): a separable triangle filter whose support grows with the down-scaling factor
  (antialiasing), coefficients normalised in double precision and rounded to 22 fractional bits
  (``precompute_coeffs``, ``normalize_coeffs_8bpc``), a horizontal pass and then a vertical pass, each accumulating
  in int32 from ``1 << 21`` and clipping ``>> 22`` to uint8 (``ImagingResampleHorizontal_8bpc`` / ``Vertical``).
* ``ToTensor`` + ``Normalize``: ``((u8 / 255) - 0.5) / 0.5`` in float32, each step rounded (torchvision
  ``functional.to_tensor`` / ``normalize``).

The training / evaluation loader takes another route (/root/reference/src/data_loader.py:31-35):
``cv2.imread(IMREAD_GRAYSCALE) -> cv2.resize(img, (320, 96))`` (INTER_LINEAR, no antialiasing) and then the same
ToTensor + Normalize (Grayscale and Resize are no-ops on a 96 x 320 "L" image).  ``cv2_resize_linear`` restates
OpenCV 4.13's 8-bit bilinear resize (modules/imgproc/src/resize.cpp: ``resizeGeneric_`` with ``HResizeLinear`` /
``VResizeLinear<uchar,int,short,...>``): source position ``(d + 0.5) * scale - 0.5`` rounded to float, 11-bit
coefficients ``cvRound(w * 2048)``, horizontal taps clamped with the weight moved to the inner pixel, vertical taps
clamped by row index only, ``(((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2``; an exact 2x
down-scale in both axes is rerouted to INTER_AREA, ``(a + b + c + d + 2) >> 2``.

Pinned by tests/test_oracle.py::test_preprocess_oracle_matches_pil_and_torchvision and
::test_dataloader_preprocess_oracle_matches_cv2 (random sizes, bit-exact against the real libraries).  Only tests/ may
import this module; the product paths are ``hmocr_preprocess_image_u8`` / ``hmocr_preprocess_cv2_u8`` (CUDA).
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def pil_gray(rgb: np.ndarray) -> np.ndarray:
    """uint8 [H, W, 3] -> uint8 [H, W]   (Convert.c: rgb2l)."""
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c: precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle, support 1) filter over the
    whole axis.  Returns (bounds int [out, 2] = (xmin, count), kk int32 [out, ksize])."""
    in0, in1 = np.float32(0.0), np.float32(in_size)
    scale = float(in1 - in0) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int64)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = float(in0) + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = []
        ww = 0.0
        for x in range(xmax):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            w = 1.0 - a if a < 1.0 else 0.0
            k.append(w)
            ww += w
        for x in range(xmax):
            v = k[x] / ww if ww != 0.0 else k[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis(img: np.ndarray, out_size: int) -> np.ndarray:
    """One 8bpc pass along the LAST axis."""
    bounds, kk = precompute_coeffs(img.shape[-1], out_size)
    out = np.empty(img.shape[:-1] + (out_size,), dtype=np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        xmin, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = (1 << (PRECISION_BITS - 1)) + (src[..., xmin:xmin + n] * kk[xx, :n].astype(np.int64)).sum(-1)
        out[..., xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_resize_bilinear(gray: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """uint8 [H, W] -> uint8 [out_h, out_w]   (ImagingResample: horizontal pass first, then vertical; a pass whose
    size does not change is skipped)."""
    x = gray
    if x.shape[1] != out_w:
        x = _resample_axis(x, out_w)
    if x.shape[0] != out_h:
        x = _resample_axis(np.ascontiguousarray(x.T), out_h).T
    return np.ascontiguousarray(x)


def to_tensor_normalize(u8: np.ndarray) -> np.ndarray:
    x = u8.astype(np.float32) / np.float32(255.0)
    return ((x - np.float32(0.5)) / np.float32(0.5)).astype(np.float32)


def reference_preprocess(img: np.ndarray, out_h: int = 96, out_w: int = 320) -> np.ndarray:
    """uint8 [H, W] (mode L) or [H, W, 3] (mode RGB) -> float32 [1, 1, out_h, out_w]."""
    gray = pil_gray(img) if img.ndim == 3 else img
    return to_tensor_normalize(pil_resize_bilinear(gray, out_h, out_w))[None, None]


def cv2_linear_coeffs(ssize: int, dsize: int, clamp_weight: bool):
    """resize.cpp: per destination index the two source indices and the 11-bit weights.  ``clamp_weight`` is the
    horizontal rule (a tap outside the image gives its weight to the border pixel: fx = 0); the vertical pass only
    clips the row indices."""
    scale = 1.0 / (dsize / ssize)
    i0 = np.zeros(dsize, np.int64)
    i1 = np.zeros(dsize, np.int64)
    w = np.zeros((dsize, 2), np.int64)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(math.floor(f))
        f = np.float32(f - np.float32(s))
        if clamp_weight:
            if s < 0:
                f, s = np.float32(0), 0
            if s >= ssize - 1:
                f, s = np.float32(0), ssize - 1
            i0[d], i1[d] = s, min(s + 1, ssize - 1)
        else:
            i0[d], i1[d] = min(max(s, 0), ssize - 1), min(max(s + 1, 0), ssize - 1)
        w[d, 0] = int(np.rint((np.float32(1.0) - f) * np.float32(2048)))
        w[d, 1] = int(np.rint(f * np.float32(2048)))
    return i0, i1, w


def cv2_resize_linear(gray: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """uint8 [H, W] -> uint8 [out_h, out_w] = cv2.resize(gray, (out_w, out_h)) (INTER_LINEAR)."""
    sh, sw = gray.shape
    if sw == 2 * out_w and sh == 2 * out_h:
        s = gray.astype(np.int64)
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    x0, x1, xa = cv2_linear_coeffs(sw, out_w, True)
    y0, y1, ya = cv2_linear_coeffs(sh, out_h, False)
    s = gray.astype(np.int64)
    rows = s[:, x0] * xa[:, 0] + s[:, x1] * xa[:, 1]
    out = (((ya[:, 0:1] * (rows[y0] >> 4)) >> 16) + ((ya[:, 1:2] * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def reference_preprocess_dataloader(gray: np.ndarray, out_h: int = 96, out_w: int = 320) -> np.ndarray:
    """src/data_loader.py:31-35 + the transform: uint8 [H, W] -> float32 [1, out_h, out_w]."""
    return to_tensor_normalize(cv2_resize_linear(gray, out_h, out_w))[None]
