"""Work partitioning: utterances are independent (episodic reset, REF/main.py:327-328), so the list is sharded
across GPUs with no data-path collective, and cut into length-bucketed adaptation batches inside a shard.

Cost model: linear in frames plus the attention term ~ T^2."""
from __future__ import annotations

from typing import List, Sequence


def utterance_cost(frames: int, hidden: int = 768, inter: int = 3072, layers: int = 12, steps: int = 10) -> float:
    lin = 2.0 * frames * (4 * hidden * hidden + 2 * hidden * inter) * layers * (2 * steps + 1)
    att = layers * float(frames) ** 2 * hidden * (12 * steps + 4)
    return lin + att


def shard_lpt(costs: Sequence[float], world_size: int) -> List[List[int]]:
    """Greedy longest-processing-time partition; returns utterance indices per rank (deterministic)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += costs[i]
    return shards


def bucket_batches(frames: Sequence[int], indices: Sequence[int], max_utts: int = 64, max_frames: int = 32768
                   ) -> List[List[int]]:
    """Sort by length and cut into batches of at most max_utts utterances / max_frames packed frames.
    Tokens are packed (no padding), so bucketing only balances attention tiles and workspace size."""
    order = sorted(indices, key=lambda i: (-frames[i], i))
    batches, cur, tot = [], [], 0
    for i in order:
        if cur and (len(cur) >= max_utts or tot + frames[i] > max_frames):
            batches.append(cur)
            cur, tot = [], 0
        cur.append(i)
        tot += frames[i]
    if cur:
        batches.append(cur)
    return batches
