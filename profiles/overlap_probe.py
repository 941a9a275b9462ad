"""Throughput of two engines fed from two host threads (each its own CUDA stream) vs one engine: how much of the
encoder of batch i+1 hides in the SMs the decode kernel of batch i leaves idle (it occupies 128 of 148).

    python profiles/overlap_probe.py [--batch 256] [--iters 12] [--spl 150]
"""
import argparse
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from handwritten_math_ocr_api_b200 import FormulaRecognitionModel
from handwritten_math_ocr_api_b200.layout import ModelConfig
from handwritten_math_ocr_api_b200.synthetic import synth_images, synth_state_dict


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--spl", type=int, default=0)
    ap.add_argument("--engines", type=int, default=2)
    a = ap.parse_args()
    cfg = ModelConfig()
    sd = synth_state_dict(cfg, seed=0, eos_bias_sigma=0.0)
    models = []
    for _ in range(a.engines):
        m = FormulaRecognitionModel(cfg.vocab_size)
        m.load_state_dict(sd)
        if a.spl:
            m.set_option("steps_per_launch", a.spl)
        models.append(m)
    imgs = synth_images(8, seed=1234).cuda().repeat(a.batch // 8, 1, 1, 1).contiguous()

    last = {}

    def worker(m, n, stream):
        with torch.cuda.stream(stream):
            for _ in range(n):
                last[id(m)] = m.generate_device(imgs, max_len=150)[0]
            stream.synchronize()

    for n_eng in range(1, a.engines + 1):
        streams = [torch.cuda.Stream() for _ in range(n_eng)]
        for m, s in zip(models, streams):
            worker(m, 2, s)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=worker, args=(m, a.iters, s)) for m, s in zip(models[:n_eng], streams)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        n = n_eng * a.iters * a.batch
        ref = last[id(models[0])]
        same = all(torch.equal(last[id(m)], ref) for m in models[:n_eng])
        print(f"engines={n_eng}: tokens identical across engines: {same};  {n} images in {dt * 1e3:.1f} ms -> {n / dt:.0f} img/s ({dt * 1e3 / (n_eng * a.iters):.2f} ms/batch)")


if __name__ == "__main__":
    main()
