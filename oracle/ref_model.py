"""Plain-PyTorch fp32 restatement of the reference model's arithmetic (oracle side).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Works directly on a reference-layout
state dict; no nn.Module, no torchvision import, so it runs on the GPU box where
``/root/reference`` does not exist.  Checked against the real reference (torchvision 0.26.0 /
torch 2.11.0 modules driven by ``/root/reference/src/model_swin.py``) by
``oracle/make_golden.py`` and against the committed outputs of that run by ``tests/``.

What each function follows
* ``patch_embed``        torchvision swin_transformer.py:556-562 (+ 1-channel conv,
                         /root/reference/src/model_swin.py:19-34)
* ``window_attention``   swin_transformer.py:116-228 (pad AFTER norm1, shift suppression,
                         roll, 7x7 partition, scaled q, bias table gather :49-56, -100 region
                         mask :193-209, reverse/unroll/crop)
* ``swin_block``         swin_transformer.py:452-455 (pre-LN residual, MLP = Linear/GELU(erf)/Linear)
* ``patch_merging``      swin_transformer.py:35-43, 76-86
* ``encoder_forward``    /root/reference/src/model_swin.py:39-46 (features only, NO final swin.norm,
                         view [B,30,768], Linear 768->256)
* ``decoder_forward``    /root/reference/src/model_swin.py:72-88 + torch nn.TransformerDecoderLayer
                         (post-LN, ReLU, eps 1e-5; MHA packed in_proj q|k|v; causal -inf mask)
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .arch import DEPTHS, HEADS, WINDOW, ModelConfig

SD = Dict[str, torch.Tensor]
EPS = 1e-5


def _ln(x, sd: SD, prefix: str):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], EPS)


def _lin(x, sd: SD, prefix: str, bias: bool = True):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"] if bias else None)


# ---------------------------------------------------------------------------------------------
# Swin-T encoder
# ---------------------------------------------------------------------------------------------

def patch_embed(images: torch.Tensor, sd: SD, fp: str) -> torch.Tensor:
    """[B,1,96,320] -> [B,24,80,96]: non-overlapping 4x4 patches as a 16->96 linear map, then LN."""
    B, _, H, W = images.shape
    p = images.reshape(B, H // 4, 4, W // 4, 4).permute(0, 1, 3, 2, 4).reshape(B, H // 4, W // 4, 16)
    w = sd[fp + "0.0.weight"].reshape(-1, 16)
    x = p @ w.t() + sd[fp + "0.0.bias"]
    return _ln(x, sd, fp + "0.2")


def window_geometry(H: int, W: int, shift: int):
    """Index plan of one (shifted) window attention over an HxW token grid.

    Returns (Hp, Wp, sh, sw, src, valid, region) where, for window-major position p
    (window row-major, token row-major inside the window), ``src[p]`` is the flat index of the
    source token in the un-padded HxW grid (or 0 with ``valid[p]`` False for a padded position)
    and ``region[p]`` is the shift-mask region id of that position (all zero if unshifted)."""
    ws = WINDOW
    Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
    sh = shift if Hp > ws else 0                 # swin_transformer.py:158-163
    sw = shift if Wp > ws else 0
    nwh, nww = Hp // ws, Wp // ws
    wr, wc, i, j = torch.meshgrid(torch.arange(nwh), torch.arange(nww), torch.arange(ws), torch.arange(ws),
                                  indexing="ij")
    r = wr * ws + i                              # coordinates in the rolled, padded map
    c = wc * ws + j
    pr = (r + sh) % Hp                           # roll by (-sh,-sw): rolled[r] = padded[r+sh]
    pc = (c + sw) % Wp
    valid = (pr < H) & (pc < W)
    src = torch.where(valid, pr * W + pc, torch.zeros_like(pr))

    def band(x, size, s):
        # slices (0,-ws), (-ws,-s), (-s,None) written in order (swin_transformer.py:196-202);
        # with s == 0 the last slice is the whole axis and overwrites the other two.
        if s == 0:
            return torch.full_like(x, 2)
        return (x >= size - ws).long() + (x >= size - s).long()

    if sh + sw > 0:
        region = band(r, Hp, sh) * 3 + band(c, Wp, sw)
    else:
        region = torch.zeros_like(r)
    flat = lambda t: t.reshape(nwh * nww, ws * ws)
    return Hp, Wp, sh, sw, flat(src), flat(valid), flat(region)


def window_attention(xn: torch.Tensor, sd: SD, bp: str, heads: int, shift: int) -> torch.Tensor:
    """xn: LayerNorm-ed input [B,H,W,C] -> attention branch output [B,H,W,C]."""
    B, H, W, C = xn.shape
    hd = C // heads
    Hp, Wp, sh, sw, src, valid, region = window_geometry(H, W, shift)
    nW, N = src.shape
    dev = xn.device
    src, valid, region = src.to(dev), valid.to(dev), region.to(dev)
    tok = xn.reshape(B, H * W, C)[:, src.reshape(-1), :] * valid.reshape(1, -1, 1).to(xn.dtype)
    tok = tok.reshape(B * nW, N, C)                              # zero rows where padded
    qkv = _lin(tok, sd, bp + "attn.qkv").reshape(B * nW, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)                               # [B*nW, heads, N, N]
    table = sd[bp + "attn.relative_position_bias_table"]
    index = sd[bp + "attn.relative_position_index"]
    bias = table[index].reshape(N, N, heads).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if sh + sw > 0:
        diff = region.unsqueeze(1) != region.unsqueeze(2)        # [nW, N, N]
        mask = torch.where(diff, torch.tensor(-100.0, device=dev), torch.tensor(0.0, device=dev))
        attn = (attn.reshape(B, nW, heads, N, N) + mask[None, :, None]).reshape(B * nW, heads, N, N)
    attn = attn.softmax(-1)
    ctx = (attn @ v).transpose(1, 2).reshape(B * nW, N, C)
    out = _lin(ctx, sd, bp + "attn.proj").reshape(B, nW * N, C)
    # scatter valid positions back to the un-padded grid (window reverse + roll back + crop)
    res = xn.new_zeros(B, H * W, C)
    vflat = valid.reshape(-1)
    res[:, src.reshape(-1)[vflat], :] = out[:, vflat, :]
    return res.reshape(B, H, W, C)


def swin_block(x: torch.Tensor, sd: SD, bp: str, heads: int, shift: int) -> torch.Tensor:
    x = x + window_attention(_ln(x, sd, bp + "norm1"), sd, bp, heads, shift)
    h = F.gelu(_lin(_ln(x, sd, bp + "norm2"), sd, bp + "mlp.0"))
    return x + _lin(h, sd, bp + "mlp.3")


def patch_merging(x: torch.Tensor, sd: SD, mp: str) -> torch.Tensor:
    B, H, W, C = x.shape
    x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
    return F.linear(_ln(x, sd, mp + "norm"), sd[mp + "reduction.weight"])


def encoder_stages(images: torch.Tensor, sd: SD, prefix: str = "encoder.features.") -> List[torch.Tensor]:
    """Residual stream after patch-embed and after each of the 12 blocks / 3 merges (for
    kernel-level parity tests)."""
    outs = []
    x = patch_embed(images, sd, prefix)
    outs.append(x)
    for s in range(4):
        fi = 1 + 2 * s
        for j in range(DEPTHS[s]):
            x = swin_block(x, sd, f"{prefix}{fi}.{j}.", HEADS[s], 0 if j % 2 == 0 else WINDOW // 2)
            outs.append(x)
        if s < 3:
            x = patch_merging(x, sd, f"{prefix}{fi + 1}.")
            outs.append(x)
    return outs


def encoder_forward(images: torch.Tensor, sd: SD) -> torch.Tensor:
    """f32 [B,1,96,320] -> f32 [B,30,d_model]."""
    x = encoder_stages(images, sd)[-1]
    B, H, W, C = x.shape
    return _lin(x.reshape(B, H * W, C), sd, "encoder.projection")


# ---------------------------------------------------------------------------------------------
# Transformer decoder
# ---------------------------------------------------------------------------------------------

def _split_heads(x, nhead):
    B, T, D = x.shape
    return x.reshape(B, T, nhead, D // nhead).transpose(1, 2)          # [B,h,T,hd]


def _mha(q_in, kv_in, sd: SD, ap: str, nhead: int, causal: bool):
    d = q_in.shape[-1]
    w, b = sd[ap + ".in_proj_weight"], sd[ap + ".in_proj_bias"]
    q = F.linear(q_in, w[:d], b[:d])
    k = F.linear(kv_in, w[d:2 * d], b[d:2 * d])
    v = F.linear(kv_in, w[2 * d:], b[2 * d:])
    q, k, v = _split_heads(q, nhead), _split_heads(k, nhead), _split_heads(v, nhead)
    s = (q @ k.transpose(-2, -1)) / math.sqrt(d // nhead)
    if causal:
        T = s.shape[-1]
        s = s + torch.triu(torch.full((T, T), float("-inf"), device=s.device), diagonal=1)
    ctx = (s.softmax(-1) @ v).transpose(1, 2).reshape(q_in.shape)
    return _lin(ctx, sd, ap + ".out_proj")


def decoder_layer(x, mem, sd: SD, lp: str, nhead: int):
    x = _ln(x + _mha(x, x, sd, lp + "self_attn", nhead, True), sd, lp + "norm1")
    x = _ln(x + _mha(x, mem, sd, lp + "multihead_attn", nhead, False), sd, lp + "norm2")
    h = F.relu(_lin(x, sd, lp + "linear1"))
    return _ln(x + _lin(h, sd, lp + "linear2"), sd, lp + "norm3")


def decoder_hidden(enc_out: torch.Tensor, tgt: torch.Tensor, sd: SD, cfg: ModelConfig) -> torch.Tensor:
    T = tgt.shape[1]
    x = sd["decoder.embedding.weight"][tgt] + sd["decoder.pos_encoder.weight"][:T].unsqueeze(0)
    for l in range(cfg.num_layers):
        x = decoder_layer(x, enc_out, sd, f"decoder.decoder.layers.{l}.", cfg.nhead)
    return x


def decoder_forward(enc_out: torch.Tensor, tgt: torch.Tensor, sd: SD, cfg: Optional[ModelConfig] = None) -> torch.Tensor:
    """enc_out f32 [B,S,d], tgt int64 [B,T] -> logits f32 [B,T,V] (all T positions)."""
    cfg = cfg or ModelConfig(vocab_size=sd["decoder.fc_out.weight"].shape[0])
    return _lin(decoder_hidden(enc_out, tgt, sd, cfg), sd, "decoder.fc_out")


def model_forward(images, captions, sd: SD, cfg: Optional[ModelConfig] = None, drop_last: bool = False):
    """``FormulaRecognitionModel.forward``: app flavour feeds ``captions`` whole
    (/root/reference/app/src/model_swin.py:100), training flavour drops the last column
    (/root/reference/src/model_swin.py:100)."""
    enc = encoder_forward(images, sd)
    return decoder_forward(enc, captions[:, :-1] if drop_last else captions, sd, cfg)


# ---------------------------------------------------------------------------------------------
# KV-cached single-step decoder (exactly the same arithmetic, one new position at a time)
# ---------------------------------------------------------------------------------------------

class CachedDecoder:
    """Incremental form of ``decoder_forward``: self-attention K/V kept per layer, encoder-memory
    K/V projected once.  SURVEY.md 7.2-1: identical tokens, |dlogit| ~2e-6 vs full recompute."""

    def __init__(self, enc_out: torch.Tensor, sd: SD, cfg: ModelConfig):
        self.sd, self.cfg = sd, cfg
        d, h = cfg.d_model, cfg.nhead
        self.mem_kv: List[Tuple[torch.Tensor, torch.Tensor]] = []
        for l in range(cfg.num_layers):
            ap = f"decoder.decoder.layers.{l}.multihead_attn"
            w, b = sd[ap + ".in_proj_weight"], sd[ap + ".in_proj_bias"]
            k = _split_heads(F.linear(enc_out, w[d:2 * d], b[d:2 * d]), h)
            v = _split_heads(F.linear(enc_out, w[2 * d:], b[2 * d:]), h)
            self.mem_kv.append((k, v))
        self.self_kv: List[Optional[Tuple[torch.Tensor, torch.Tensor]]] = [None] * cfg.num_layers
        self.t = 0

    def reorder(self, parent: torch.Tensor) -> None:
        """Beam search: row i continues hypothesis ``parent[i]``."""
        self.self_kv = [None if kv is None else (kv[0][parent], kv[1][parent]) for kv in self.self_kv]

    def step(self, tok: torch.Tensor, mem_index: Optional[torch.Tensor] = None) -> torch.Tensor:
        """tok int64 [B] at position self.t -> logits [B,V]."""
        sd, cfg = self.sd, self.cfg
        d, h = cfg.d_model, cfg.nhead
        scale = 1.0 / math.sqrt(d // h)
        x = (sd["decoder.embedding.weight"][tok] + sd["decoder.pos_encoder.weight"][self.t]).unsqueeze(1)
        for l in range(cfg.num_layers):
            lp = f"decoder.decoder.layers.{l}."
            w, b = sd[lp + "self_attn.in_proj_weight"], sd[lp + "self_attn.in_proj_bias"]
            qkv = F.linear(x, w, b)
            q, k, v = (_split_heads(t, h) for t in qkv.split(d, dim=-1))
            if self.self_kv[l] is not None:
                k = torch.cat([self.self_kv[l][0], k], 2)
                v = torch.cat([self.self_kv[l][1], v], 2)
            self.self_kv[l] = (k, v)
            ctx = ((q @ k.transpose(-2, -1) * scale).softmax(-1) @ v).transpose(1, 2).reshape(x.shape)
            x = _ln(x + _lin(ctx, sd, lp + "self_attn.out_proj"), sd, lp + "norm1")
            w, b = sd[lp + "multihead_attn.in_proj_weight"], sd[lp + "multihead_attn.in_proj_bias"]
            q = _split_heads(F.linear(x, w[:d], b[:d]), h)
            mk, mv = self.mem_kv[l]
            if mem_index is not None:
                mk, mv = mk[mem_index], mv[mem_index]
            ctx = ((q @ mk.transpose(-2, -1) * scale).softmax(-1) @ mv).transpose(1, 2).reshape(x.shape)
            x = _ln(x + _lin(ctx, sd, lp + "multihead_attn.out_proj"), sd, lp + "norm2")
            hdn = F.relu(_lin(x, sd, lp + "linear1"))
            x = _ln(x + _lin(hdn, sd, lp + "linear2"), sd, lp + "norm3")
        self.t += 1
        return _lin(x[:, 0], sd, "decoder.fc_out")
