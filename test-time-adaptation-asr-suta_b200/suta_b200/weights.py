"""Weight sources: random initialisation with HF's magnitudes (no checkpoint offline) or a local HF checkpoint."""
from __future__ import annotations

import math
import os
from typing import Dict

import torch

from .config import ModelConfig


def random_state_dict(cfg: ModelConfig, seed: int = 0, blank_bias: float = 0.0, special_bias: float = 0.0) -> Dict[str, torch.Tensor]:
    """State dict with HF parameter names and the scales of HF/modeling_wav2vec2.py:968-1003 (_init_weights).
    `blank_bias` shifts lm_head.bias[0] so that greedy decoding emits blanks like a trained CTC head does."""
    g = torch.Generator().manual_seed(seed)
    H, I = cfg.hidden_size, cfg.intermediate_size
    sd: Dict[str, torch.Tensor] = {}
    rn = lambda shape, std: torch.randn(shape, generator=g) * std
    ru = lambda shape, k: (torch.rand(shape, generator=g) * 2 - 1) * k
    cin = 1
    for i, (c, k) in enumerate(zip(cfg.conv_dim, cfg.conv_kernel)):
        sd[f"wav2vec2.feature_extractor.conv_layers.{i}.conv.weight"] = rn((c, cin, k), math.sqrt(2.0 / (cin * k)))
        if cfg.conv_bias:
            sd[f"wav2vec2.feature_extractor.conv_layers.{i}.conv.bias"] = ru((c,), math.sqrt(1.0 / (cin * k)))
        cin = c

    def ln(prefix, n):
        sd[prefix + ".weight"], sd[prefix + ".bias"] = torch.ones(n), torch.zeros(n)

    for i in range(len(cfg.conv_dim) if cfg.feat_extract_norm == "layer" else 1):      # LayerNorm per layer / one GroupNorm
        ln(f"wav2vec2.feature_extractor.conv_layers.{i}.layer_norm", cfg.conv_dim[i])
    ln("wav2vec2.feature_projection.layer_norm", cfg.conv_dim[-1])
    kk = math.sqrt(1.0 / cfg.conv_dim[-1])
    sd["wav2vec2.feature_projection.projection.weight"] = ru((H, cfg.conv_dim[-1]), kk)
    sd["wav2vec2.feature_projection.projection.bias"] = ru((H,), kk)
    K, G = cfg.num_conv_pos_embeddings, cfg.num_conv_pos_embedding_groups
    sd["wav2vec2.encoder.pos_conv_embed.conv.weight"] = rn((H, H // G, K), 2 * math.sqrt(1.0 / (K * (H // G))))
    sd["wav2vec2.encoder.pos_conv_embed.conv.bias"] = torch.zeros(H)
    ln("wav2vec2.encoder.layer_norm", H)
    for l in range(cfg.num_hidden_layers):
        p = f"wav2vec2.encoder.layers.{l}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + f"attention.{nm}.weight"], sd[p + f"attention.{nm}.bias"] = rn((H, H), 0.02), torch.zeros(H)
        ln(p + "layer_norm", H)
        sd[p + "feed_forward.intermediate_dense.weight"], sd[p + "feed_forward.intermediate_dense.bias"] = rn((I, H), 0.02), torch.zeros(I)
        sd[p + "feed_forward.output_dense.weight"], sd[p + "feed_forward.output_dense.bias"] = rn((H, I), 0.02), torch.zeros(H)
        ln(p + "final_layer_norm", H)
    sd["lm_head.weight"] = rn((cfg.vocab_size, H), 0.02)
    b = torch.zeros(cfg.vocab_size)
    b[0] = blank_bias
    b[1:4] = special_bias              # <s>, </s>, <unk>: a trained CTC head never emits them
    sd["lm_head.bias"] = b
    return sd


def load_checkpoint(path: str):
    """(ModelConfig, state_dict) from a local HF Wav2Vec2ForCTC directory (config.json + weights)."""
    from transformers import Wav2Vec2Config
    cfg = ModelConfig.from_any(Wav2Vec2Config.from_pretrained(path))
    st = os.path.join(path, "model.safetensors")
    if os.path.exists(st):
        from safetensors.torch import load_file
        sd = load_file(st)
    else:
        sd = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu")
    if "wav2vec2.encoder.pos_conv_embed.conv.weight_g" in sd:      # old weight_norm naming
        sd["wav2vec2.encoder.pos_conv_embed.conv.parametrizations.weight.original0"] = sd.pop("wav2vec2.encoder.pos_conv_embed.conv.weight_g")
        sd["wav2vec2.encoder.pos_conv_embed.conv.parametrizations.weight.original1"] = sd.pop("wav2vec2.encoder.pos_conv_embed.conv.weight_v")
    return cfg, sd
