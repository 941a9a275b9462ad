"""Per-shape timing of the tcgen05 GEMM on the Swin-T shapes of the bench workload (B=256, 96x320).

For every nn.Linear of the encoder: time in isolation (inputs larger than L2 for the stage-1/2 shapes), the HBM
floor of its algorithmic bytes, and the same launch with parts of the epilogue switched off (gemm_dbg: 1 = no
stores, 2 = no residual loads, 4 = main loop + TMEM load only) to see what bounds it.

    python profiles/gemm_shapes.py [--bn N]
"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from handwritten_math_ocr_api_b200 import FormulaRecognitionModel, _lib
from handwritten_math_ocr_api_b200.layout import ModelConfig
from handwritten_math_ocr_api_b200.synthetic import synth_state_dict

HBM = 6549.1e9


def P(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bn", type=int, default=0)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--flags", type=int, nargs="*", default=[0, 1, 3, 7])
    ap.add_argument("--only", default="", help="substring filter on the gemm name (for an ncu capture)")
    ap.add_argument("--ncu", action="store_true", help="ONE launch per shape after one warm-up (for an ncu metrics pass; "
                                                       "prints the launch order so the csv rows can be named)")
    a = ap.parse_args()
    cfg = ModelConfig()
    m = FormulaRecognitionModel(cfg.vocab_size)
    m.load_state_dict(synth_state_dict(cfg, seed=0, eos_bias_sigma=0.0))
    lib = _lib.load()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    shapes = []
    tok = a.batch * 24 * 80
    for s, c in enumerate((96, 192, 384, 768)):
        M = tok >> (2 * s)
        shapes += [(f"s{s + 1} qkv", M, c, 3 * c, "f16", 0, False), (f"s{s + 1} proj", M, c, c, "f32", 0, True),
                   (f"s{s + 1} fc1", M, c, 4 * c, "f16", 1, False), (f"s{s + 1} fc2", M, 4 * c, c, "f32", 0, True)]
        if s < 3:
            shapes.append((f"s{s + 1} merge", M // 4, 4 * c, 2 * c, "f32", 0, False))
    shapes.append(("fc_out 768->256", a.batch * 30, 768, 256, "f32", 0, False))
    print(f"{'gemm':16s} {'M':>7s} {'K':>5s} {'N':>5s} {'floor_us':>9s} " + " ".join(f"dbg{f}_us".rjust(9) for f in a.flags))
    total = {f: 0.0 for f in a.flags}
    floor_total = 0.0
    for name, M, K, N, out, act, res in shapes:
        if a.only and a.only not in name:
            continue
        A = torch.randn(M, K, device="cuda").half()
        W = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
        bias = torch.randn(N, device="cuda")
        R = torch.randn(M, N, device="cuda") if res else None
        o32 = torch.empty(M, N, device="cuda") if out == "f32" else None
        o16 = torch.empty(M, N, device="cuda", dtype=torch.float16) if out == "f16" else None
        nbytes = M * K * 2 + N * K * 2 + M * N * (4 if out == "f32" else 2) + (M * N * 4 if res else 0)
        floor = nbytes / HBM * 1e6
        floor_total += floor
        line = f"{name:16s} {M:7d} {K:5d} {N:5d} {floor:9.1f} "
        if a.ncu:
            def run1():
                rc = lib.hmocr_gemm_f16(P(A), K, M, K, P(W), N, P(bias), act, P(R), N if res else 0, P(o32), N, P(o16), N,
                                        None, None, a.bn, st)
                _lib.check(rc, "gemm")
            run1(); run1()
            torch.cuda.synchronize()
            print(f"NCU_ORDER {name}|{M}|{K}|{N}|{2.0 * M * K * N:.0f}")
            continue
        for f in a.flags:
            m.set_option("gemm_dbg", f)

            def run():
                rc = lib.hmocr_gemm_f16(P(A), K, M, K, P(W), N, P(bias), act, P(R), N if res else 0, P(o32), N, P(o16), N,
                                        None, None, a.bn, st)
                _lib.check(rc, "gemm")
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 100
            mult = {"s3": 3, "s1": 1, "s2": 1, "s4": 1, "fc": 1}[name[:2]] if "merge" not in name else 1
            total[f] += us * (2 if "merge" not in name and "fc_out" not in name else 1) * mult
            line += " ".join([f"{us:9.1f}"]) + " "
        print(line)
    m.set_option("gemm_dbg", 0)
    print(f"encoder GEMM total (x blocks per stage): floor {floor_total:.0f} us (one of each)  " +
          "  ".join(f"dbg{f}: {total[f]:.0f} us" for f in a.flags))


if __name__ == "__main__":
    main()
