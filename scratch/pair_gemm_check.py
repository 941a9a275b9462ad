"""Correctness + timing of the CTA-pair GEMM (force_bn 2192 / 2256) against the 1-CTA kernel and torch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from gpu_utils import gemm, gemm_ref

torch.manual_seed(0)
ok = True
for (M, K, N, fb) in [(256, 64, 192, 2192), (512, 384, 384, 2192), (300, 384, 1152, 2192), (1024, 1536, 384, 2192),
                      (768, 768, 768, 2256), (7680, 768, 2304, 2256), (515, 384, 1536, 2256)]:
    a = torch.randn(M, K, device="cuda").half()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    b = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda")
    for kind in ("f16_gelu", "f32_res"):
        if kind == "f16_gelu":
            _, o = gemm(a, w, bias=b, act=1, out_f32=False, out_f16=True, force_bn=fb)
            _, o1 = gemm(a, w, bias=b, act=1, out_f32=False, out_f16=True)
            ref = gemm_ref(a, w, bias=b, act=1)
            o, o1 = o.float(), o1.float()
        else:
            o, _ = gemm(a, w, bias=b, residual=r, force_bn=fb)
            o1, _ = gemm(a, w, bias=b, residual=r)
            ref = gemm_ref(a, w, bias=b, residual=r)
        e = (o - ref).abs().max().item(); e1 = (o - o1).abs().max().item()
        good = e < 2e-2 and e1 < 5e-3
        ok &= good
        print(f"M={M} K={K} N={N} bn={fb} {kind}: err vs torch {e:.2e}, vs 1-CTA {e1:.2e} {'ok' if good else 'FAIL'}")
print("ALL OK" if ok else "FAILED")

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

import ctypes as C
from handwritten_math_ocr_api_b200 import _lib
lib = _lib.load()
P = lambda x: None if x is None else C.c_void_p(x.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, M, K, N, out16, act, res in [("s3 qkv", 30720, 384, 1152, True, 0, False), ("s3 proj", 30720, 384, 384, False, 0, True),
                                       ("s3 fc1", 30720, 384, 1536, True, 1, False), ("s3 fc2", 30720, 1536, 384, False, 0, True),
                                       ("s4 qkv", 7680, 768, 2304, True, 0, False), ("s4 proj", 7680, 768, 768, False, 0, True),
                                       ("s4 fc1", 7680, 768, 3072, True, 1, False), ("s4 fc2", 7680, 3072, 768, False, 0, True),
                                       ("s2 qkv", 122880, 192, 576, True, 0, False), ("s2 fc2", 122880, 768, 192, False, 0, True)]:
    a = torch.randn(M, K, device="cuda").half(); w = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    b = torch.randn(N, device="cuda"); r = torch.randn(M, N, device="cuda") if res else None
    o32 = None if out16 else torch.empty(M, N, device="cuda"); o16 = torch.empty(M, N, device="cuda", dtype=torch.float16) if out16 else None
    line = f"{name:8s} M={M} K={K} N={N}: "
    for fb in (0, 2192, 2256):
        if fb and N % (fb - 2000): line += f"  bn{fb}: n/a   "; continue
        def run():
            rc = lib.hmocr_gemm_f16(P(a), K, M, K, P(w), N, P(b), act, P(r), N if res else 0, P(o32), N, P(o16), N, None, None, fb, st)
            assert rc == 0, lib.hmocr_last_error()
        line += f"  bn{fb}: {t(run):6.1f} us"
    print(line)
