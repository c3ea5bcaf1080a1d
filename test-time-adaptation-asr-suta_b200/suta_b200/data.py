"""Synthetic utterance sets shaped like the corpora the reference evaluates on (no datasets offline).

Replaces REF/data.py + REF/corpus/*.py for benchmarking: LibriSpeech-test-other-shaped durations (SURVEY.md 8d
config 2: 2939 utterances, clip(lognormal(median 5.5 s, sigma 0.6), 2, 35)), audio = 0.1*randn per utterance with a
per-utterance seed, optional REF/data.py:23 extra noise, REF/data.py:19-21 truncation to 600000 samples."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np

SAMPLE_RATE = 16000
MAX_SAMPLES = 600000           # REF/data.py:9,19-21
_WORDS = ["THE", "OF", "AND", "A", "TO", "IN", "HE", "WAS", "THAT", "IT", "HIS", "HER", "WITH", "AS", "HAD", "FOR"]


@dataclass
class Utterance:
    index: int
    n_samples: int
    text: str
    seed: int
    extra_noise: float = 0.0

    @property
    def duration(self) -> float:
        return self.n_samples / SAMPLE_RATE

    def audio(self, with_noise: bool = True) -> np.ndarray:
        """The waveform as the reference's collate function returns it (REF/data.py:15-23).  with_noise=False leaves the
        extra_noise term out: the batched runner adds it on the device (suta_batch_add_noise)."""
        rng = np.random.default_rng(self.seed)
        wav = (0.1 * rng.standard_normal(self.n_samples)).astype(np.float32)
        if with_noise and self.extra_noise > 0:
            wav = wav + (self.extra_noise * rng.standard_normal(self.n_samples)).astype(np.float32)
        return wav


def librispeech_shaped(n_utts: int = 2939, seed: int = 0, extra_noise: float = 0.0) -> List[Utterance]:
    rng = np.random.default_rng(seed)
    dur = np.clip(rng.lognormal(np.log(5.5), 0.6, n_utts), 2.0, 35.0)
    out = []
    for i, d in enumerate(dur):
        n = min(int(round(d * SAMPLE_RATE)), MAX_SAMPLES)
        nw = max(1, int(d * 2.7))
        text = " ".join(_WORDS[j] for j in rng.integers(0, len(_WORDS), nw))
        out.append(Utterance(i, n, text, seed=1000 + i, extra_noise=extra_noise))
    return out


def fixed_length(n_utts: int, seconds: float, seed: int = 0) -> List[Utterance]:
    n = int(round(seconds * SAMPLE_RATE))
    return [Utterance(i, n, "THE", seed=seed * 100003 + i) for i in range(n_utts)]
