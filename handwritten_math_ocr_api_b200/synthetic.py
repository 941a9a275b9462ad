"""Synthetic workload: checkpoint, stroke images and vocabulary (data definition only - no model
arithmetic lives here; used by bench.py, smoke(), the tests and the golden-vector generator).

The reference ships neither a checkpoint nor a dataset, and torch's CPU
``normal_`` is not guaranteed bit-identical across CPU generations, so "a fixed random-init
seed" is realised here as a platform-independent generator: every parameter element is
``splitmix64(fnv1a(name) ^ seed, index)`` mapped to a uniform float.  The same bytes come out in
the build container (where the golden vectors are made from the real reference) and on the GPU
box (where the engine and the oracle are compared).

Every parameter is non-trivial on purpose (LayerNorm gains != 1, all biases != 0, a large
relative-position table) so that a kernel that drops a term cannot pass parity.

Images follow SURVEY.md section 8(d): uint8 canvas 96x320, white background, 3-12 random-walk
polylines of 8-40 points, thickness 2-3 px, black ink, then ToTensor + Normalize(0.5, 0.5)
(``/root/reference/src/predict.py:36-41``) => float32 [B,1,96,320] in [-1, 1].
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from .layout import (IMG_H, IMG_W, RES18_MEM_TOKENS, WINDOW, ModelConfig, state_dict_layout,
                     state_dict_layout_res18)

_M64 = (1 << 64) - 1


def _fnv1a64(s: str) -> int:
    h = 0xCBF29CE484222325
    for b in s.encode():
        h = ((h ^ b) * 0x100000001B3) & _M64
    return h


def _uniform(name: str, n: int, seed: int) -> np.ndarray:
    """n deterministic floats in (-1, 1), float64, from integer arithmetic only."""
    base = np.uint64(_fnv1a64(name) ^ ((seed * 0x9E3779B97F4A7C15) & _M64))
    with np.errstate(over="ignore"):
        z = base + (np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    v = (z >> np.uint64(40)).astype(np.float64)              # 24 random bits
    return (v + 0.5) / float(1 << 23) - 1.0


def relative_position_index() -> np.ndarray:
    """((dh+6)*13 + (dw+6)) for every (query, key) pair of a 7x7 window, flattened [2401]
    (torchvision swin_transformer.py:272-284)."""
    idx = np.arange(WINDOW * WINDOW)
    r, c = idx // WINDOW, idx % WINDOW
    dh = r[:, None] - r[None, :] + WINDOW - 1
    dw = c[:, None] - c[None, :] + WINDOW - 1
    return (dh * (2 * WINDOW - 1) + dw).reshape(-1).astype(np.int64)


def _scale_for(name: str, shape, cfg: ModelConfig, fc_gain: float) -> tuple:
    """(kind, scale, offset): value = offset + scale * u, u ~ U(-1,1)."""
    s3 = math.sqrt(3.0)
    leaf = name.rsplit(".", 1)[-1]
    if "relative_position_bias_table" in name:
        return 1.0, 0.0
    if "norm" in name and leaf == "weight":
        return 0.2, 1.0
    if "norm" in name and leaf == "bias":
        return 0.1, 0.0
    if name.endswith("embedding.weight"):
        return s3, 0.0
    if name.endswith("pos_encoder.weight"):
        return s3, 0.0
    if leaf == "bias" or leaf == "in_proj_bias":
        return 0.1, 0.0
    # weights: std = gain / sqrt(fan_in)
    fan_in = int(np.prod(shape[1:]))
    gain = 1.0
    if name.endswith("attn.proj.weight") or name.endswith("mlp.3.weight"):
        gain = 0.5
    # decoder sub-layer outputs are kept small so the token/position embedding (not a constant
    # cross-attention term) drives the next token: greedy decodes wander instead of repeating
    if name.endswith("self_attn.out_proj.weight") or name.endswith("linear2.weight"):
        gain = 0.2
    if name.endswith("multihead_attn.out_proj.weight"):
        gain = 0.35
    if name.endswith("fc_out.weight"):
        gain = fc_gain
    return gain * s3 / math.sqrt(fan_in), 0.0


def next_token_permutation(vocab_size: int) -> np.ndarray:
    """The fixed map prev-token -> favoured next token of the ``peak`` checkpoints: 4 + (2 v + 7) mod (V - 4)
    (a permutation of the non-special ids when V - 4 is odd, as for V = 5075)."""
    return 4 + (np.arange(vocab_size, dtype=np.int64) * 2 + 7) % (vocab_size - 4)


def synth_state_dict(cfg: ModelConfig | None = None, seed: int = 0, fc_gain: float = 4.0,
                     eos_bias_sigma: float = 2.5, peak: float = 0.0) -> Dict[str, torch.Tensor]:
    """Reference-layout state dict (517 entries, ``encoder.features.*`` aliasing
    ``encoder.swin.features.*``) filled with the deterministic synthetic values.

    ``eos_bias_sigma`` lifts ``fc_out.bias[eos]`` by that many logit standard deviations so
    greedy decodes terminate at varied lengths (0 => essentially never emits EOS, the fixed
    150-step benchmark workload of SURVEY.md section 8d).

    ``peak`` > 0 gives the logits the PEAKED shape a trained model has.  Scaling ``fc_out``
    (``fc_gain``) cannot do that: it scales the margins and the arithmetic error alike, so the rate of near-tie argmax
    flips does not change (measured: 96.2 / 96.6 / 96.6 % identical sequences at fc_gain 4 / 16 / 32).  Instead a
    readout of the previous token is added: ``fc_out.weight[next(v)] += peak * embedding[v] / |embedding[v]|``.  The
    final hidden state still carries the fed token's embedding (the decoder's sub-layer outputs are small), so the
    favoured token's logit rises by about ``0.7 * peak * |e|``; at peak 8 the oracle's top-1 probability has median
    0.97 (random init: 0.30) while the image still decides the uncertain steps."""
    cfg = cfg or ModelConfig()
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, dtype in state_dict_layout(cfg):
        if name.startswith("encoder.features."):
            sd[name] = sd["encoder.swin.features." + name[len("encoder.features."):]]
            continue
        if name.endswith("relative_position_index"):
            sd[name] = torch.from_numpy(relative_position_index())
            continue
        if name == "decoder.tgt_mask":
            m = torch.triu(torch.full((cfg.max_seq_len, cfg.max_seq_len), float("-inf")), diagonal=1)
            sd[name] = m
            continue
        n = int(np.prod(shape))
        scale, offset = _scale_for(name, shape, cfg, fc_gain)
        vals = (offset + scale * _uniform(name, n, seed)).astype(np.float32).reshape(shape)
        sd[name] = torch.from_numpy(vals)
    if eos_bias_sigma:
        sd["decoder.fc_out.bias"][cfg.eos] += float(eos_bias_sigma) * fc_gain
    if peak:
        emb = sd["decoder.embedding.weight"].numpy().astype(np.float64)
        unit = emb / np.sqrt((emb * emb).sum(1, keepdims=True))
        w = sd["decoder.fc_out.weight"].numpy().astype(np.float64)
        np.add.at(w, next_token_permutation(cfg.vocab_size), float(peak) * unit)
        sd["decoder.fc_out.weight"] = torch.from_numpy(w.astype(np.float32))
    return sd


def synth_state_dict_res18(cfg: ModelConfig | None = None, seed: int = 0, fc_gain: float = 4.0,
                           eos_bias_sigma: float = 2.5) -> Dict[str, torch.Tensor]:
    """Synthetic checkpoint in the 367-entry layout of /root/reference/src/model_res18trans.py.
    Conv weights are He-scaled, BatchNorm running statistics non-trivial (mean ~ U(-0.2,0.2), var ~ U(0.5,1.5))."""
    cfg = cfg or ModelConfig()
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, dtype in state_dict_layout_res18(cfg):
        leaf = name.rsplit(".", 1)[-1]
        if name == "decoder.tgt_mask":
            sd[name] = torch.triu(torch.full((cfg.max_seq_len, cfg.max_seq_len), float("-inf")), diagonal=1)
            continue
        if leaf == "num_batches_tracked":
            sd[name] = torch.tensor(1000, dtype=torch.int64)
            continue
        n = int(np.prod(shape))
        u = _uniform(name, n, seed)
        if leaf == "running_var":
            vals = 1.0 + 0.5 * u
        elif leaf == "running_mean":
            vals = 0.2 * u
        elif len(shape) == 4:                                   # conv: He-uniform on fan_in
            fan_in = int(np.prod(shape[1:]))
            vals = math.sqrt(6.0 / fan_in) * u
        elif ".bn" in name or "downsample.1" in name or name.startswith("encoder.features.1."):
            vals = (1.0 + 0.2 * u) if leaf == "weight" else 0.1 * u
        else:
            key = name.replace("decoder.transformer_decoder.", "decoder.decoder.")
            scale, offset = _scale_for(key, shape, cfg, fc_gain)
            vals = offset + scale * u
        sd[name] = torch.from_numpy(np.asarray(vals, dtype=np.float32).reshape(shape))
    if eos_bias_sigma:
        sd["decoder.fc_out.bias"][cfg.eos] += float(eos_bias_sigma) * fc_gain
    return sd


def synth_pos_table(d_model: int = 256, seed: int = 0) -> torch.Tensor:
    """Stand-in for the nn.Embedding(10, d_model) the reference re-creates (N(0,1)-initialised) on EVERY encoder
    call (src/model_res18trans.py:57-59): the table is an explicit input of the engine."""
    u = _uniform("res18trans.pos_embed", RES18_MEM_TOKENS * d_model, seed)
    return torch.from_numpy((math.sqrt(3.0) * u).astype(np.float32).reshape(RES18_MEM_TOKENS, d_model))


def state_dict_checksum(sd: Dict[str, torch.Tensor]) -> str:
    """Order-independent 64-bit checksum of the parameter BYTES (to prove two boxes hold the
    same checkpoint)."""
    acc = 0
    for k in sorted(sd):
        if k.startswith("encoder.features."):
            continue
        a = sd[k].detach().cpu().contiguous().numpy().view(np.uint8)
        h = _fnv1a64(k)
        # 64-bit folded sum of the bytes viewed as uint64 words (padded)
        pad = (-a.size) % 8
        w = np.concatenate([a.reshape(-1), np.zeros(pad, np.uint8)]).view(np.uint64)
        with np.errstate(over="ignore"):
            s = int(np.bitwise_xor.reduce(w * (np.arange(1, w.size + 1, dtype=np.uint64) | np.uint64(1))))
        acc = (acc + ((h ^ s) * 0x9E3779B97F4A7C15)) & _M64
    return f"{acc:016x}"


# ---------------------------------------------------------------------------------------------
# images
# ---------------------------------------------------------------------------------------------

def _draw_segment(canvas: np.ndarray, x0, y0, x1, y1, thick: float) -> None:
    h, w = canvas.shape
    r = thick / 2.0
    xa, xb = int(math.floor(min(x0, x1) - r - 1)), int(math.ceil(max(x0, x1) + r + 1))
    ya, yb = int(math.floor(min(y0, y1) - r - 1)), int(math.ceil(max(y0, y1) + r + 1))
    xa, xb, ya, yb = max(xa, 0), min(xb, w - 1), max(ya, 0), min(yb, h - 1)
    if xa > xb or ya > yb:
        return
    ys, xs = np.mgrid[ya:yb + 1, xa:xb + 1].astype(np.float64)
    dx, dy = x1 - x0, y1 - y0
    L2 = dx * dx + dy * dy
    if L2 == 0.0:
        t = np.zeros_like(xs)
    else:
        t = np.clip(((xs - x0) * dx + (ys - y0) * dy) / L2, 0.0, 1.0)
    d2 = (xs - (x0 + t * dx)) ** 2 + (ys - (y0 + t * dy)) ** 2
    canvas[ya:yb + 1, xa:xb + 1][d2 <= r * r] = 0


def synth_stroke_image_u8(rs: np.random.RandomState) -> np.ndarray:
    canvas = np.full((IMG_H, IMG_W), 255, np.uint8)
    for _ in range(rs.randint(3, 13)):
        npts = rs.randint(8, 41)
        thick = float(rs.randint(2, 4))
        x, y = rs.uniform(8, IMG_W - 8), rs.uniform(8, IMG_H - 8)
        ang = rs.uniform(0, 2 * math.pi)
        for _ in range(npts - 1):
            ang += rs.normal(0.0, 0.6)
            step = rs.uniform(2.0, 7.0)
            nx = float(np.clip(x + step * math.cos(ang), 1, IMG_W - 2))
            ny = float(np.clip(y + step * math.sin(ang), 1, IMG_H - 2))
            _draw_segment(canvas, x, y, nx, ny, thick)
            x, y = nx, ny
    return canvas


def synth_images(batch: int, seed: int = 1234) -> torch.Tensor:
    """float32 [B,1,96,320] in [-1,1]; image i depends only on (seed, i)."""
    out = np.empty((batch, 1, IMG_H, IMG_W), np.float32)
    for i in range(batch):
        rs = np.random.RandomState((seed * 1000003 + i) % (2 ** 31 - 1))
        u8 = synth_stroke_image_u8(rs)
        out[i, 0] = (u8.astype(np.float32) / 255.0 - 0.5) / 0.5
    return torch.from_numpy(out)


def synth_images_u8(batch: int, seed: int = 1234) -> torch.Tensor:
    """The same images before the reference transform: uint8 [B,96,320] (255 = paper, 0 = ink)."""
    out = np.empty((batch, IMG_H, IMG_W), np.uint8)
    for i in range(batch):
        rs = np.random.RandomState((seed * 1000003 + i) % (2 ** 31 - 1))
        out[i] = synth_stroke_image_u8(rs)
    return torch.from_numpy(out)


def synth_vocab(vocab_size: int):
    """vocab / idx2char in the reference's ``vocab.json`` shape
    (``/root/reference/app/src/utils.py:10-15``, specials first: ``src/config.py:43-47``)."""
    toks = ["<pad>", "<sos>", "<eos>", "<unk>"] + [f"t{i}" for i in range(4, vocab_size)]
    vocab = {t: i for i, t in enumerate(toks)}
    idx2char = {i: t for i, t in enumerate(toks)}
    return vocab, idx2char
