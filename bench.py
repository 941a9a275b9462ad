#!/usr/bin/env python
"""Headline benchmark: greedy image-to-LaTeX throughput (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch: 256 synthetic rasterised stroke images per
GPU -> Swin-T encoder -> 150-step KV-cached greedy decode -> token ids (N > 1: + the all-gather
of the ids, the path's only collective).  Prints ONE JSON line (rank 0).

  value      images/s, whole job, images already resident in HBM when the timed region starts
  e2e        images/s through the public API with HOST buffers (pinned images in, ids out)
  roofline   decode phase against the measured HBM copy bandwidth (algorithmic bytes of
             SURVEY.md 8d / DESIGN.md); `encoder` sub-object: tensor roofline of the encoder
  cpu_baseline  the oracle port of the reference greedy loop (src/inference.py) on the host cores

`--impl reference` times only that CPU path (the reference is pure Python over torch/torchvision
and cannot be pip-installed as a package; its algorithm is restated in oracle/, checked against
the real reference by oracle/make_golden.py).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

T_MAX = 150
CPU_SAMPLE_IMAGES = 8                # bounded CPU sample: one batch of 8 images x 150 full-prefix greedy steps (~3 s on 16 cores)
ENC_GFLOP_PER_IMAGE = 6.515          # reference-executed; 5.498 if padded rows are skipped
ENC_GFLOP_MINIMAL = 5.498


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        try:
            d = json.load(open(f))
            p.update({k: d[k] for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in d})
            p["source"] = "measured"
        except Exception:
            pass
    return p


def decode_algorithmic_bytes(batch: int, T: int, layers=8, d=256, S=30, weight_params=7_630_547) -> float:
    """bf16 KV: per sequence per step cross K+V read L*S*2*d*2 B, self cache read L*2*d*2 B per
    cached position, append L*2*d*2 B; decoder weights (bf16) once per step per device."""
    cross = layers * S * 2 * d * 2
    per_pos = layers * 2 * d * 2
    per_seq = sum(cross + per_pos * t + per_pos for t in range(1, T + 1))
    return batch * per_seq + weight_params * 2 * T


def decode_traffic_from_ncu(batch: int, T: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of the decode kernel launches of ONE step, from the
    committed ncu capture (profiles/decode_traffic.json); None if no capture matches this workload."""
    f = os.path.join(ROOT, "profiles", "decode_traffic.json")
    try:
        d = json.load(open(f))
        if d.get("batch") == batch and d.get("max_len") == T:
            return d["dram_bytes_per_step"]
    except Exception:
        pass
    return None


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU baseline (oracle port of the reference loop)
# --------------------------------------------------------------------------------------------------
def cpu_reference_sample(n_images: int, max_len: int, reps: int, warmup: int):
    from oracle import decode as odec
    from oracle.arch import ModelConfig
    from oracle.synth import synth_images, synth_state_dict
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = ModelConfig()
    sd = synth_state_dict(cfg, seed=0, eos_bias_sigma=0.0)
    imgs = synth_images(n_images, seed=1234)
    times = []
    with torch.no_grad():
        for i in range(warmup + reps):
            t0 = time.perf_counter()
            ys = odec.greedy_batched(imgs, sd, cfg, max_len=max_len)     # src/inference.py:7-25 restated
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    steps = ys.shape[1] - 1
    return {"times": times, "images": n_images, "tokens": n_images * steps, "cores": torch.get_num_threads()}


def run_reference(args, rank):
    if rank != 0:
        return
    n_img, max_len = CPU_SAMPLE_IMAGES, T_MAX
    r = cpu_reference_sample(n_img, max_len, reps=args.steps, warmup=min(args.warmup, 1))
    mean = sum(r["times"]) / len(r["times"])
    ips = n_img / mean
    sample = f"{n_img} images x {max_len} greedy steps per step (oracle port of src/inference.py, torch eager fp32)"
    line = {
        "impl": "reference", "metric": "images_per_sec", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": mean * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "tokens_per_sec": r["tokens"] / mean,
        "config": {"workload": "swin_t+8L-decoder greedy, T=150, V=5075 (BASELINE.json configs[1]) - CPU sample",
                   "batch_per_step": n_img, "max_len": max_len},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": r["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist
    from handwritten_math_ocr_api_b200 import FormulaRecognitionModel, _lib
    from handwritten_math_ocr_api_b200.parallel import gather_tokens, gather_tokens_device
    from handwritten_math_ocr_api_b200.layout import ModelConfig
    from handwritten_math_ocr_api_b200.preprocess import preprocess_u8
    from handwritten_math_ocr_api_b200.synthetic import synth_images, synth_images_u8, synth_state_dict

    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T = args.batch, args.max_len
    cfg = ModelConfig()
    sd = synth_state_dict(cfg, seed=0, eos_bias_sigma=0.0)     # never emits eos: exactly B*T tokens per step
    model = FormulaRecognitionModel(cfg.vocab_size, device=dev)
    model.load_state_dict(sd)
    host_u8 = synth_images_u8(B, seed=1234 + rank).pin_memory()      # the rendered strokes as they come: uint8 96x320
    host_imgs = synth_images(B, seed=1234 + rank)                    # = ToTensor + Normalize(0.5, 0.5) of host_u8
    dev_imgs = host_imgs.to(dev)
    lib = _lib.load()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2

    def step_device():
        if world == 1:
            tokens, steps, _ = model.generate(dev_imgs, max_len=T)
            return tokens, steps
        # stream-ordered: decode -> the path's only collective (NCCL all-gather of the ids) -> one host read
        tok, st, _ = model.generate_device(dev_imgs, max_len=T)
        all_tok, all_steps = gather_tokens_device(tok, st)
        return all_tok, int(all_steps.max().item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.hmocr_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    enc_ms, dec_ms = [], []
    barrier()
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                         # evict L2 between timed iterations (not timed)
        ev[i][0].record()
        tokens, steps = step_device()
        ev[i][1].record()
        e_ms, d_ms = model.last_timings_ms()          # CUDA events recorded by the library on this stream
        enc_ms.append(e_ms); dec_ms.append(d_ms)
    barrier()
    wall = time.perf_counter() - wall0
    launches = lib.hmocr_launch_count() - launches0
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    t = torch.tensor([sum(step_ms) / len(step_ms), sum(enc_ms) / len(enc_ms), sum(dec_ms) / len(dec_ms)],
                     dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, ms_enc, ms_dec = t.tolist()
    assert steps == T, f"workload must run exactly {T} steps, ran {steps}"

    # ---- end to end through the public API with host buffers -------------------------------------
    import ctypes as C
    tok_host = torch.empty(B, T + 1, dtype=torch.int64).pin_memory()
    steps_host = torch.zeros(1, dtype=torch.int32).pin_memory()

    all_host = torch.empty(world * B, T + 1, dtype=torch.int64).pin_memory() if world > 1 else None

    def step_e2e():
        if world == 1:
            _lib.check(lib.hmocr_generate_host_u8(model._handle(), C.c_void_p(host_u8.data_ptr()), B, T, 1,
                                                  C.c_void_p(tok_host.data_ptr()), None, C.c_void_p(steps_host.data_ptr()),
                                                  None, C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                       "generate_host_u8")
            return
        # host images in, gathered ids of the whole job out (every rank reads them), all stream-ordered
        x = preprocess_u8(model, host_u8.to(dev, non_blocking=True))
        tok, st, _ = model.generate_device(x, max_len=T)
        all_tok, all_steps = gather_tokens_device(tok, st)
        all_host.copy_(all_tok, non_blocking=True)
        steps_host.copy_(all_steps.max().to(torch.int32).reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    step_e2e()
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - e0) / args.steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = te.item()

    # ---- B=1 latency (p50 per-image latency of BASELINE.json's metric) -----------------------------
    lat = []
    one = dev_imgs[:1].contiguous()
    for i in range(2 + 5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        model.generate(one, max_len=T)
        torch.cuda.synchronize()
        if i >= 2:
            lat.append((time.perf_counter() - t0) * 1e3)

    if rank == 0:
        pk = peaks()
        total_imgs = B * world
        value = total_imgs / (ms_step * 1e-3)
        dec_bytes = decode_algorithmic_bytes(B, T)
        ach = dec_bytes / (ms_dec * 1e-3) / 1e9
        enc_tf = ENC_GFLOP_PER_IMAGE * 1e9 * B / (ms_enc * 1e-3) / 1e12
        line = {
            "metric": "images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
            "tokens_per_sec": total_imgs * T / (ms_step * 1e-3),
            "p50_ms_per_image_b1": statistics.median(lat), "ms_per_image_at_batch": ms_step / B,
            "encoder_ms": ms_enc, "decode_ms": ms_dec, "wall_ms_per_step": wall * 1e3 / args.steps,
            "config": {"workload": "swin_t+8L-decoder greedy decode, B=256/GPU, T=150, V=5075, d_model=256 "
                                   "(BASELINE.json configs[1])",
                       "batch_per_gpu": B, "max_len": T, "l2_flush_between_steps": True,
                       "weights": "synthetic seed 0 (oracle/synth.py, eos never emitted)",
                       "parallelism": f"dp{world}"},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": ach / pk["hbm_gbs"], "traffic": decode_traffic_from_ncu(B, T),
                         "peak_source": pk["source"],
                         "kernel": "decode_persistent_kernel: the decode phase of one step = ceil(150/16) launches of "
                                   "the persistent cluster kernel (+ memory K/V projection); bytes and CUDA-event time "
                                   "are summed over them; 2-byte (fp16) KV caches",
                         "algorithmic_bytes_per_step": dec_bytes,
                         "encoder": {"bound": "tensor", "achieved": enc_tf, "peak": pk["bf16_tflops_sustained"],
                                     "unit": "TFLOP/s", "frac": enc_tf / pk["bf16_tflops_sustained"],
                                     "gflop_per_image": ENC_GFLOP_PER_IMAGE,
                                     "gflop_per_image_minimal": ENC_GFLOP_MINIMAL}},
            "e2e": {"value": total_imgs / e2e_s, "unit": "images/s", "h2d_bytes_per_step": B * 96 * 320,
                    "input": "pinned uint8 [B,96,320] rendered strokes; ToTensor + Normalize(0.5,0.5) on the device "
                             "(hmocr_generate_host_u8), token ids back to pinned host memory",
                    "d2h_bytes_per_step": (B if world == 1 else world * B) * (T + 1) * 8 + 4},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            r = cpu_reference_sample(CPU_SAMPLE_IMAGES, T, reps=3, warmup=0)
            m = sum(r["times"]) / len(r["times"])
            line["cpu_baseline"] = {"value": CPU_SAMPLE_IMAGES / m, "unit": "images/s", "cores": r["cores"],
                                    "kind": "port",
                                    "sample": f"{CPU_SAMPLE_IMAGES} images x 150 greedy steps in one batch, oracle port of "
                                              "src/inference.py (full-prefix recompute, torch eager fp32), mean of 3 repetitions"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--max-len", type=int, default=T_MAX)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            # launched without torchrun: re-exec under it
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                   "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
