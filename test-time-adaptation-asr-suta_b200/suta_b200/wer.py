"""Word error rate with jiwer.wer semantics (corpus-level (S+D+I)/N over whitespace-split words), as used at
REF/main.py:336,353,...,408-417.  jiwer itself is not a dependency here.  Counts are kept as integers so that
shards can be summed exactly across ranks."""
from __future__ import annotations

from typing import Sequence, Tuple


def edit_distance(ref: Sequence[str], hyp: Sequence[str]) -> int:
    if len(ref) < len(hyp):
        ref, hyp = hyp, ref                       # distance is symmetric; keep the inner row short
    prev = list(range(len(hyp) + 1))
    for i, r in enumerate(ref, 1):
        cur = [i] + [0] * len(hyp)
        for j, h in enumerate(hyp, 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (r != h))
        prev = cur
    return prev[-1]


def wer_counts(refs: Sequence[str], hyps: Sequence[str]) -> Tuple[int, int]:
    if len(refs) != len(hyps):
        raise ValueError("refs and hyps differ in length")
    e = n = 0
    for r, h in zip(refs, hyps):
        rw, hw = r.split(), h.split()
        e += edit_distance(rw, hw)
        n += len(rw)
    return e, n


def wer(refs, hyps) -> float:
    if isinstance(refs, str):
        refs, hyps = [refs], [hyps]
    e, n = wer_counts(refs, hyps)
    return e / n if n else float("nan")
