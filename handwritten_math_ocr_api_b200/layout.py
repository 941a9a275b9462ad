"""Architecture constants and the 517-entry state-dict layout of the reference model.

Follows ``/root/reference/src/model_swin.py:13-101`` (EncoderSwin / DecoderTransformer /
FormulaRecognitionModel), ``/root/reference/src/config.py:16-40`` (hyper-parameters) and
torchvision ``swin_t`` (``swin_transformer.py:804-837``: patch 4, embed 96, depths [2,2,6,2],
heads [3,6,12,24], window 7).  Pure data definition (no arithmetic); checked entry by entry against
the real reference's ``state_dict()`` (tests/golden/state_dict_manifest.json).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

IMG_H, IMG_W = 96, 320            # src/config.py:17-18
PATCH = 4
EMBED = 96
DEPTHS = (2, 2, 6, 2)
HEADS = (3, 6, 12, 24)
WINDOW = 7
SWIN_HEAD_DIM = 32
MEM_TOKENS = 30                   # 3 x 10 stage-4 grid
SPECIALS = {"<pad>": 0, "<sos>": 1, "<eos>": 2, "<unk>": 3}   # src/config.py:43-47, src/utils.py:101-112


@dataclass
class ModelConfig:
    vocab_size: int = 5075        # images/mlflow-swin.png
    d_model: int = 256            # src/config.py:19
    nhead: int = 8                # src/config.py:20
    dim_feedforward: int = 512    # src/config.py:21
    num_layers: int = 8           # src/config.py:32
    max_seq_len: int = 150        # src/config.py:40
    sos: int = 1
    eos: int = 2
    pad: int = 0

    @property
    def head_dim(self) -> int:
        return self.d_model // self.nhead


def stage_dims() -> List[Tuple[int, int, int, int]]:
    """(H, W, C, heads) of the token grid inside each of the four Swin stages."""
    h, w, c = IMG_H // PATCH, IMG_W // PATCH, EMBED
    out = []
    for s in range(4):
        out.append((h, w, c, HEADS[s]))
        h, w, c = h // 2, w // 2, c * 2
    return out


def _swin_feature_keys() -> List[Tuple[str, Tuple[int, ...], str]]:
    keys: List[Tuple[str, Tuple[int, ...], str]] = []
    f32, i64 = "float32", "int64"
    keys += [("0.0.weight", (EMBED, 1, PATCH, PATCH), f32), ("0.0.bias", (EMBED,), f32),
             ("0.2.weight", (EMBED,), f32), ("0.2.bias", (EMBED,), f32)]
    c = EMBED
    for s in range(4):
        fi = 1 + 2 * s
        for j in range(DEPTHS[s]):
            p = f"{fi}.{j}."
            keys += [
                (p + "norm1.weight", (c,), f32), (p + "norm1.bias", (c,), f32),
                (p + "attn.relative_position_bias_table", ((2 * WINDOW - 1) ** 2, HEADS[s]), f32),
                (p + "attn.relative_position_index", (WINDOW ** 4,), i64),
                (p + "attn.qkv.weight", (3 * c, c), f32), (p + "attn.qkv.bias", (3 * c,), f32),
                (p + "attn.proj.weight", (c, c), f32), (p + "attn.proj.bias", (c,), f32),
                (p + "norm2.weight", (c,), f32), (p + "norm2.bias", (c,), f32),
                (p + "mlp.0.weight", (4 * c, c), f32), (p + "mlp.0.bias", (4 * c,), f32),
                (p + "mlp.3.weight", (c, 4 * c), f32), (p + "mlp.3.bias", (c,), f32),
            ]
        if s < 3:
            p = f"{fi + 1}."
            keys += [(p + "reduction.weight", (2 * c, 4 * c), f32),
                     (p + "norm.weight", (4 * c,), f32), (p + "norm.bias", (4 * c,), f32)]
            c *= 2
    return keys


def state_dict_layout(cfg: ModelConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """The 517 (name, shape, dtype) entries of ``FormulaRecognitionModel(V).state_dict()``
    in the reference's own order."""
    f32 = "float32"
    feats = _swin_feature_keys()
    out: List[Tuple[str, Tuple[int, ...], str]] = []
    out += [("encoder.swin.features." + k, s, d) for k, s, d in feats]
    out += [("encoder.swin.norm.weight", (768,), f32), ("encoder.swin.norm.bias", (768,), f32),
            ("encoder.swin.head.weight", (1000, 768), f32), ("encoder.swin.head.bias", (1000,), f32)]
    out += [("encoder.features." + k, s, d) for k, s, d in feats]      # same storage, second name
    d, ff, v = cfg.d_model, cfg.dim_feedforward, cfg.vocab_size
    out += [("encoder.projection.weight", (d, 768), f32), ("encoder.projection.bias", (d,), f32),
            ("decoder.tgt_mask", (cfg.max_seq_len, cfg.max_seq_len), f32),
            ("decoder.embedding.weight", (v, d), f32),
            ("decoder.pos_encoder.weight", (cfg.max_seq_len, d), f32)]
    for l in range(cfg.num_layers):
        p = f"decoder.decoder.layers.{l}."
        for att in ("self_attn", "multihead_attn"):
            out += [(p + att + ".in_proj_weight", (3 * d, d), f32), (p + att + ".in_proj_bias", (3 * d,), f32),
                    (p + att + ".out_proj.weight", (d, d), f32), (p + att + ".out_proj.bias", (d,), f32)]
        out += [(p + "linear1.weight", (ff, d), f32), (p + "linear1.bias", (ff,), f32),
                (p + "linear2.weight", (d, ff), f32), (p + "linear2.bias", (d,), f32)]
        for n in ("norm1", "norm2", "norm3"):
            out += [(p + n + ".weight", (d,), f32), (p + n + ".bias", (d,), f32)]
    out += [("decoder.fc_out.weight", (v, d), f32), ("decoder.fc_out.bias", (v,), f32)]
    return out


# ------------------------------------------------------------------------------------------------
# ResNet-18 + TransformerEncoder variant (/root/reference/src/model_res18trans.py, BASELINE.json config 4)
# ------------------------------------------------------------------------------------------------
RES18_MEM_TOKENS = 10             # [B,512,3,10] -> AdaptiveAvgPool (1,None) -> 10 memory tokens
RES18_STAGES = ((64, 1), (128, 2), (256, 2), (512, 2))     # (channels, stride of the first block) of layer1..4


def _bn_keys(p: str, c: int):
    f32 = "float32"
    return [(p + ".weight", (c,), f32), (p + ".bias", (c,), f32), (p + ".running_mean", (c,), f32),
            (p + ".running_var", (c,), f32), (p + ".num_batches_tracked", (), "int64")]


def state_dict_layout_res18(cfg: ModelConfig, enc_layers: int = 0) -> List[Tuple[str, Tuple[int, ...], str]]:
    """The 367 (name, shape, dtype) entries of model_res18trans.FormulaRecognitionModel(V).state_dict()
    in the reference's own order (checked against the live reference by oracle/make_golden_res18.py)."""
    f32 = "float32"
    out: List[Tuple[str, Tuple[int, ...], str]] = []
    f = "encoder.features."
    out += [(f + "0.weight", (64, 1, 7, 7), f32)] + _bn_keys(f + "1", 64)
    cin = 64
    for s, (c, stride) in enumerate(RES18_STAGES):
        for j in range(2):
            p = f"{f}{4 + s}.{j}."
            out += [(p + "conv1.weight", (c, cin if j == 0 else c, 3, 3), f32)] + _bn_keys(p + "bn1", c)
            out += [(p + "conv2.weight", (c, c, 3, 3), f32)] + _bn_keys(p + "bn2", c)
            if j == 0 and (stride != 1 or cin != c):
                out += [(p + "downsample.0.weight", (c, cin, 1, 1), f32)] + _bn_keys(p + "downsample.1", c)
        cin = c
    d, ff, v = cfg.d_model, cfg.dim_feedforward, cfg.vocab_size
    out += [("encoder.projection.weight", (d, 512), f32), ("encoder.projection.bias", (d,), f32)]
    for l in range(enc_layers or cfg.num_layers):        # config.res18trans_num_encoder_layers (src/config.py:28)
        p = f"encoder.transformer_encoder.layers.{l}."
        out += [(p + "self_attn.in_proj_weight", (3 * d, d), f32), (p + "self_attn.in_proj_bias", (3 * d,), f32),
                (p + "self_attn.out_proj.weight", (d, d), f32), (p + "self_attn.out_proj.bias", (d,), f32),
                (p + "linear1.weight", (ff, d), f32), (p + "linear1.bias", (ff,), f32),
                (p + "linear2.weight", (d, ff), f32), (p + "linear2.bias", (d,), f32)]
        for n in ("norm1", "norm2"):
            out += [(p + n + ".weight", (d,), f32), (p + n + ".bias", (d,), f32)]
    out += [("decoder.tgt_mask", (cfg.max_seq_len, cfg.max_seq_len), f32),
            ("decoder.embedding.weight", (v, d), f32),
            ("decoder.pos_encoder.weight", (cfg.max_seq_len, d), f32)]
    for l in range(cfg.num_layers):
        p = f"decoder.transformer_decoder.layers.{l}."
        for att in ("self_attn", "multihead_attn"):
            out += [(p + att + ".in_proj_weight", (3 * d, d), f32), (p + att + ".in_proj_bias", (3 * d,), f32),
                    (p + att + ".out_proj.weight", (d, d), f32), (p + att + ".out_proj.bias", (d,), f32)]
        out += [(p + "linear1.weight", (ff, d), f32), (p + "linear1.bias", (ff,), f32),
                (p + "linear2.weight", (d, ff), f32), (p + "linear2.bias", (d,), f32)]
        for n in ("norm1", "norm2", "norm3"):
            out += [(p + n + ".weight", (d,), f32), (p + n + ".bias", (d,), f32)]
    out += [("decoder.fc_out.weight", (v, d), f32), ("decoder.fc_out.bias", (v,), f32)]
    return out
