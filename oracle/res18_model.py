"""Plain-PyTorch fp32 restatement of /root/reference/src/model_res18trans.py (BASELINE.json config 4).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Works on the reference-layout state dict (367 entries).

What each function follows
* ``trunk``            ``EncoderCNN.features`` = torchvision resnet18 children[:-2] with a 1-channel conv1
                       (src/model_res18trans.py:16-32): conv7x7/2 -> BN -> ReLU -> maxpool3x3/2 -> 4 x 2 BasicBlocks
                       (torchvision resnet.py BasicBlock: conv3x3-BN-ReLU-conv3x3-BN, + identity / conv1x1-BN, ReLU);
                       BatchNorm in eval mode (running statistics, eps 1e-5)
* ``encoder_forward``  src/model_res18trans.py:48-64: AdaptiveAvgPool2d((1,None)) -> permute -> Linear 512->256 ->
                       + positional table -> permute to [10, B, d] -> 8 x TransformerEncoderLayer(batch_first=True)
                       -> permute back.  NOTE: with batch_first=True the layers see batch = 10 and
                       sequence = B, i.e. self-attention runs ACROSS THE IMAGES OF THE BATCH; and the reference
                       draws a fresh N(0,1) positional table on every call - here it is an explicit argument.
* ``decoder_forward``  src/model_res18trans.py:88-103 = the Swin model's decoder with the module renamed
                       ``transformer_decoder``
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

from .arch import ModelConfig
from . import ref_model as swin_oracle

SD = Dict[str, torch.Tensor]
EPS = 1e-5
DEC_PREFIX = "decoder.transformer_decoder.layers."


def _bn(x, sd: SD, p: str):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=EPS)


def _basic_block(x, sd: SD, p: str, stride: int):
    out = F.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"], None, stride, 1), sd, p + "bn1"))
    out = _bn(F.conv2d(out, sd[p + "conv2.weight"], None, 1, 1), sd, p + "bn2")
    if p + "downsample.0.weight" in sd:
        x = _bn(F.conv2d(x, sd[p + "downsample.0.weight"], None, stride, 0), sd, p + "downsample.1")
    return F.relu(out + x)


def trunk(images: torch.Tensor, sd: SD) -> torch.Tensor:
    """[B,1,96,320] -> [B,512,3,10]."""
    f = "encoder.features."
    x = F.relu(_bn(F.conv2d(images, sd[f + "0.weight"], None, 2, 3), sd, f + "1"))
    x = F.max_pool2d(x, 3, 2, 1)
    for s, stride in enumerate((1, 2, 2, 2)):
        x = _basic_block(x, sd, f"{f}{4 + s}.0.", stride)
        x = _basic_block(x, sd, f"{f}{4 + s}.1.", 1)
    return x


def _mha(x, sd: SD, p: str, nhead: int):
    """Self-attention of a [N, S, d] batch-first tensor (no mask)."""
    N, S, d = x.shape
    hd = d // nhead
    qkv = F.linear(x, sd[p + "in_proj_weight"], sd[p + "in_proj_bias"])
    q, k, v = (t.reshape(N, S, nhead, hd).transpose(1, 2) for t in qkv.split(d, dim=-1))
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), dim=-1)
    ctx = (att @ v).transpose(1, 2).reshape(N, S, d)
    return F.linear(ctx, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def encoder_layers(x: torch.Tensor, sd: SD, cfg: ModelConfig) -> torch.Tensor:
    """x [10, B, d] as the reference feeds it: post-LN TransformerEncoderLayer, ReLU, batch = dim 0."""
    for l in range(cfg.num_layers):
        p = f"encoder.transformer_encoder.layers.{l}."
        x = F.layer_norm(x + _mha(x, sd, p + "self_attn.", cfg.nhead), (cfg.d_model,), sd[p + "norm1.weight"],
                         sd[p + "norm1.bias"], EPS)
        h = F.linear(F.relu(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"])), sd[p + "linear2.weight"],
                     sd[p + "linear2.bias"])
        x = F.layer_norm(x + h, (cfg.d_model,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], EPS)
    return x


def encoder_forward(images: torch.Tensor, sd: SD, cfg: ModelConfig, pos_table: torch.Tensor) -> torch.Tensor:
    """-> [B, 10, d]."""
    x = trunk(images, sd)                                  # [B,512,3,10]
    x = x.mean(dim=2, keepdim=True)                        # AdaptiveAvgPool2d((1, None))
    x = x.permute(0, 3, 2, 1)                              # [B,10,1,512]
    x = F.linear(x, sd["encoder.projection.weight"], sd["encoder.projection.bias"]).squeeze(2)
    x = x + pos_table.unsqueeze(0)                         # [B,10,d]
    x = encoder_layers(x.permute(1, 0, 2), sd, cfg)        # [10,B,d]: attention over B
    return x.permute(1, 0, 2)


def _as_swin_decoder_sd(sd: SD) -> SD:
    return {k.replace("decoder.transformer_decoder.", "decoder.decoder."): v for k, v in sd.items()}


def decoder_forward(enc: torch.Tensor, tgt: torch.Tensor, sd: SD, cfg: ModelConfig) -> torch.Tensor:
    return swin_oracle.decoder_forward(enc, tgt, _as_swin_decoder_sd(sd), cfg)


def greedy_cached(enc: torch.Tensor, sd: SD, cfg: ModelConfig, max_len=None, return_logits=False):
    from . import decode as odec
    return odec.greedy_cached(enc, _as_swin_decoder_sd(sd), cfg, max_len=max_len, return_logits=return_logits)
