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

    @staticmethod
    def base() -> "ModelConfig":              # facebook/wav2vec2-base-960h architecture
        return ModelConfig()

    @staticmethod
    def large() -> "ModelConfig":             # wav2vec2-large-960h shape (BASELINE.json configs[3])
        return ModelConfig(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096)

    @staticmethod
    def tiny() -> "ModelConfig":              # small shape for tests; same kernels
        return ModelConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                           conv_dim=(64,) * 7, num_conv_pos_embeddings=16, num_conv_pos_embedding_groups=2)

    @staticmethod
    def from_any(cfg) -> "ModelConfig":
        """Accept a ModelConfig, an HF Wav2Vec2Config or any object with the same attribute names."""
        if isinstance(cfg, ModelConfig):
            return cfg
        if getattr(cfg, "feat_extract_norm", "group") != "group" or getattr(cfg, "do_stable_layer_norm", False):
            raise NotImplementedError("only the group-norm / post-LN wav2vec2 family (base, large-960h) is supported")
        if getattr(cfg, "conv_bias", False):
            raise NotImplementedError("conv_bias=True is not supported")
        return ModelConfig(**{f: (tuple(getattr(cfg, f)) if isinstance(getattr(cfg, f), (list, tuple)) else getattr(cfg, f))
                              for f in ModelConfig.__dataclass_fields__})

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
