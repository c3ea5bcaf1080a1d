"""Model configuration (the fields of HF Wav2Vec2Config the hot path reads, HF/configuration_wav2vec2.py:165-211)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple


@dataclass(frozen=True)
class ModelConfig:
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    vocab_size: int = 32
    conv_dim: Tuple[int, ...] = (512,) * 7
    conv_kernel: Tuple[int, ...] = (10, 3, 3, 3, 3, 2, 2)
    conv_stride: Tuple[int, ...] = (5, 2, 2, 2, 2, 2, 2)
    num_conv_pos_embeddings: int = 128
    num_conv_pos_embedding_groups: int = 16
    layer_norm_eps: float = 1e-5
    # the "lv60" family (large-960h-lv60[-self], large-robust; REF/main_SDPL.py:238-241): every conv layer is
    # Conv1d(bias) -> LayerNorm(C) -> GELU (HF/modeling_wav2vec2.py:275-299), the encoder is pre-LN (:612-655,730-803)
    feat_extract_norm: str = "group"
    conv_bias: bool = False
    do_stable_layer_norm: bool = False

    @staticmethod
    def base() -> "ModelConfig":              # facebook/wav2vec2-base-960h architecture
        return ModelConfig()

    @staticmethod
    def large() -> "ModelConfig":             # wav2vec2-large-960h shape (BASELINE.json configs[3])
        return ModelConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096)

    @staticmethod
    def large_lv60() -> "ModelConfig":        # facebook/wav2vec2-large-960h-lv60(-self) architecture
        return ModelConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096,
                           feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True)

    @staticmethod
    def tiny_lv60() -> "ModelConfig":         # the same variant at test size
        return ModelConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                           conv_dim=(64,) * 7, num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=2,
                           feat_extract_norm="layer", conv_bias=True, do_stable_layer_norm=True)

    @staticmethod
    def tiny() -> "ModelConfig":              # small shape for tests; same kernels
        return ModelConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                           conv_dim=(64,) * 7, num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=2)

    @staticmethod
    def from_any(cfg) -> "ModelConfig":
        """Accept a ModelConfig, an HF Wav2Vec2Config or any object with the same attribute names."""
        if isinstance(cfg, ModelConfig):
            cfg.validate()
            return cfg
        out = ModelConfig(**{f: (tuple(getattr(cfg, f)) if isinstance(getattr(cfg, f), (list, tuple)) else getattr(cfg, f))
                             for f in ModelConfig.__dataclass_fields__ if hasattr(cfg, f)})
        out.validate()
        return out

    def validate(self):
        if self.feat_extract_norm not in ("group", "layer"):
            raise ValueError(f"feat_extract_norm must be 'group' or 'layer', got {self.feat_extract_norm!r}")
        if self.conv_bias and self.feat_extract_norm == "group":
            raise NotImplementedError("conv_bias=True is built for the LayerNorm feature extractor only (no HF checkpoint "
                                      "pairs it with the GroupNorm one)")

    def frames(self, n_samples: int) -> int:
        """HF/modeling_wav2vec2.py:1012-1018."""
        L = n_samples
        for k, s in zip(self.conv_kernel, self.conv_stride):
            L = (L - k) // s + 1
        return L

    @property
    def min_samples(self) -> int:
        n = 1
        for k, s in zip(reversed(self.conv_kernel), reversed(self.conv_stride)):
            n = (n - 1) * s + k
        return n
