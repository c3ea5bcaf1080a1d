"""suta_b200: B200-native single-utterance test-time adaptation (SUTA) for wav2vec2-CTC models.

Host side in Python/PyTorch, arithmetic in hand-written sm_100a CUDA behind a C ABI (include/suta_b200.h).
"""
from .config import ModelConfig  # noqa: F401
from .engine import AdaptHyper, SutaEngine, CHECKPOINT_STEPS  # noqa: F401
