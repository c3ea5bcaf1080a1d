"""CPU: the N>1 path (shard -> adapt -> gather) with world_size 2 over gloo; the per-rank adaptation is replaced by a
deterministic stand-in so that only the sharding and the end-of-run exchange are exercised."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from suta_b200.data import librispeech_shaped
from suta_b200.runner import gather_results
from suta_b200.shard import shard_lpt
from suta_b200.wer import wer_counts


def _fake_transcript(u, step):
    words = u.text.split()
    return " ".join(words[: max(1, len(words) - (u.index + step) % 3)])


def _local(utts, idx, steps):
    texts = {s: {i: _fake_transcript(utts[i], s) for i in idx} for s in steps}
    counts = {s: wer_counts([utts[i].text for i in sorted(idx)], [texts[s][i] for i in sorted(idx)]) for s in steps}
    return dict(texts=texts, wer_counts=counts, wall_s=1.0, audio_s=sum(utts[i].duration for i in idx))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    utts = librispeech_shaped(40)
    steps = [0, 1, 10]
    shard = shard_lpt([u.duration for u in utts], world)[rank]
    out = gather_results(_local(utts, shard, steps), steps)
    if rank == 0:
        q.put((out["wer_counts"], {s: len(d) for s, d in out["texts"].items()}))
    dist.destroy_process_group()


def test_two_rank_gather_equals_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    counts, sizes = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    utts = librispeech_shaped(40)
    ref = _local(utts, list(range(40)), [0, 1, 10])
    assert counts == ref["wer_counts"]
    assert all(n == 40 for n in sizes.values())


def test_single_process_gather_is_identity():
    utts = librispeech_shaped(5)
    loc = _local(utts, list(range(5)), [0, 10])
    out = gather_results(loc, [0, 10])
    assert out["wer_counts"] == loc["wer_counts"] and set(out["wer"]) == {0, 10}
