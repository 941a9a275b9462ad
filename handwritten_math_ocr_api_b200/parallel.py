"""Data-parallel sharding of the hot path over the GPUs of one node.

The path shards by image: every image's encoder features and token sequence
depend on nothing else, weights are replicated, and nothing is exchanged during encode/decode.
The ONLY collective is the gather of the emitted token ids at the end of a batch
(``[B_local, 1+steps]`` int64, ~300 KB per rank at B=256, T=150) - ``gather_tokens``.

One process per GPU (``torchrun``), NCCL on GPUs; the same code runs on gloo/CPU tensors, which is
how ``tests/test_parallel_cpu.py`` covers it at world_size 2.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of ``n`` items owned by ``rank`` (first ``n % world`` ranks get one more)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_tokens(tokens: torch.Tensor, pad_id: int = 0, group=None) -> torch.Tensor:
    """All-gather per-rank token matrices ``[B_r, 1+steps_r]`` into ``[sum B_r, 1+max steps]`` in rank
    (= global image) order.  Ranks may hold different batch sizes and different step counts (each
    rank stops when ITS rows have all emitted eos); shorter rows are padded with ``pad_id`` exactly
    as the single-GPU result pads finished rows."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tokens
    world = dist.get_world_size(group)
    dev = tokens.device
    meta = torch.tensor([tokens.shape[0], tokens.shape[1]], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    sizes = [(int(m[0]), int(m[1])) for m in metas]
    bmax, tmax = max(s[0] for s in sizes), max(s[1] for s in sizes)
    padded = torch.full((bmax, tmax), pad_id, dtype=tokens.dtype, device=dev)
    padded[: tokens.shape[0], : tokens.shape[1]] = tokens
    out = torch.empty(world * bmax, tmax, dtype=tokens.dtype, device=dev)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    rows = [out[r * bmax: r * bmax + sizes[r][0]] for r in range(world)]
    return torch.cat(rows, 0)


def gather_tokens_device(tokens_full: torch.Tensor, steps: torch.Tensor, group=None):
    """Stream-ordered gather for equal per-rank batches: ``tokens_full`` is the full-width
    ``[B, 1+max_len]`` buffer of ``generate_device`` (pad past each rank's own step count), ``steps`` its
    int32 ``[1]`` device counter.  Two collectives, no host synchronisation, no metadata exchange:
    returns ``(tokens [world*B, 1+max_len], steps [world])`` still on the device."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tokens_full, steps
    world = dist.get_world_size(group)
    out = torch.empty(world * tokens_full.shape[0], tokens_full.shape[1], dtype=tokens_full.dtype,
                      device=tokens_full.device)
    all_steps = torch.empty(world, dtype=steps.dtype, device=steps.device)
    nvtx = tokens_full.is_cuda
    if nvtx:
        torch.cuda.nvtx.range_push("hmocr.gather_tokens")        # NVTX range per phase
    dist.all_gather_into_tensor(out, tokens_full.contiguous(), group=group)
    dist.all_gather_into_tensor(all_steps, steps, group=group)
    if nvtx:
        torch.cuda.nvtx.range_pop()
    return out, all_steps


def generate_sharded(model, images: torch.Tensor, max_len: Optional[int] = None, group=None) -> torch.Tensor:
    """Every rank passes the same full batch (or any object with ``shape[0]``); each decodes its
    contiguous shard with ``model.generate`` and all ranks return the full ``[N, 1+steps]`` ids."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(images.shape[0], rank, world)
    tokens = model.generate(images[lo:hi], max_len=max_len)[0]
    return gather_tokens(tokens, pad_id=getattr(model, "pad_id", 0), group=group)
