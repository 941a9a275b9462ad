"""Helpers for the ``-m gpu`` tests: call libhmocr's kernel-level C ABI on torch CUDA tensors."""
import ctypes as C

import torch

from handwritten_math_ocr_api_b200 import _lib


def P(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def S():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def gemm(a16, w16, bias=None, act=0, residual=None, out_f32=True, out_f16=False, ln=None, force_bn=0, lda=None):
    lib = _lib.load()
    M, K = a16.shape
    N = w16.shape[0]
    o32 = torch.empty(M, N, dtype=torch.float32, device="cuda") if out_f32 else None
    o16 = torch.empty(M, N, dtype=torch.float16, device="cuda") if out_f16 else None
    g, b = (ln if ln is not None else (None, None))
    rc = lib.hmocr_gemm_f16(P(a16), lda or a16.stride(0), M, K, P(w16), N, P(bias), act, P(residual),
                             N if residual is not None else 0, P(o32), N, P(o16), N, P(g), P(b), force_bn, S())
    _lib.check(rc, "hmocr_gemm_f16")
    torch.cuda.synchronize()
    return o32, o16


def gemm_ref(a16, w16, bias=None, act=0, residual=None, ln=None):
    y = a16.float() @ w16.float().t()
    if bias is not None:
        y = y + bias
    if act == 1:
        y = torch.nn.functional.gelu(y)
    elif act == 2:
        y = torch.relu(y)
    if residual is not None:
        y = y + residual
    if ln is not None:
        y = torch.nn.functional.layer_norm(y, (y.shape[-1],), ln[0], ln[1], 1e-5)
    return y
