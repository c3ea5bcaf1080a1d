"""CPU: property tests (hypothesis) of the host logic around the hot path -- partitioning, bucketing, greedy CTC text,
WER -- against the oracle restatement and their defining invariants (size-independent properties)."""
from hypothesis import given, settings, strategies as st

from oracle import suta_oracle as O
from suta_b200.shard import bucket_batches, shard_lpt
from suta_b200.text import CTCVocab
from suta_b200.wer import wer_counts

WORDS = st.lists(st.sampled_from(["A", "BE", "SEE", "DEE", "E'S", "EFF"]), min_size=0, max_size=12).map(" ".join)


@settings(max_examples=200, deadline=None)
@given(st.lists(st.integers(0, 31), min_size=0, max_size=80))
def test_ctc_text_equals_oracle_and_is_idempotent_under_frame_repetition(ids):
    v = CTCVocab()
    col = O.ctc_collapse(ids)
    assert v.ids_to_text(col) == O.ctc_ids_to_text(col)
    # repeating every frame changes nothing (repeats collapse); inserting a blank between two frames only separates repeats
    assert O.ctc_collapse([i for i in ids for _ in range(3)]) == col
    assert all(a != 0 for a in col)                                                 # no blank survives
    assert O.ctc_collapse([x for i in ids for x in (i, 0)]) == [i for i in ids if i != 0]   # blanks separate repeats


@settings(max_examples=200, deadline=None)
@given(st.lists(st.tuples(WORDS.filter(lambda s: len(s.split()) > 0), WORDS), min_size=1, max_size=6))
def test_wer_counts_match_oracle_and_metric_properties(pairs):
    refs, hyps = [r for r, _ in pairs], [h for _, h in pairs]
    e, n = wer_counts(refs, hyps)
    assert (e, n) == O.wer_counts(refs, hyps)
    assert n == sum(len(r.split()) for r in refs)
    assert wer_counts(refs, refs)[0] == 0                                           # identity
    # corpus counts are sums of per-utterance counts (what the multi-GPU all-reduce relies on)
    assert e == sum(wer_counts([r], [h])[0] for r, h in pairs)
    # edit distance is bounded by the longer sequence and at least the length difference
    for r, h in pairs:
        d = wer_counts([r], [h])[0]
        assert abs(len(r.split()) - len(h.split())) <= d <= max(len(r.split()), len(h.split()))


@settings(max_examples=150, deadline=None)
@given(st.lists(st.integers(1, 1749), min_size=1, max_size=300), st.integers(1, 8), st.integers(1, 64), st.integers(1749, 40000))
def test_sharding_and_bucketing_are_partitions_within_limits(frames, world, max_utts, max_frames):
    costs = [float(f) * (1 + f / 500.0) for f in frames]
    shards = shard_lpt(costs, world)
    assert sorted(i for s in shards for i in s) == list(range(len(frames)))         # a partition, nothing lost or duplicated
    loads = [sum(costs[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= max(costs) + 1e-9                              # the LPT guarantee
    assert shard_lpt(costs, world) == shards                                         # deterministic
    for s in shards:
        batches = bucket_batches(frames, s, max_utts, max_frames)
        assert sorted(i for b in batches for i in b) == sorted(s)
        for b in batches:
            assert 1 <= len(b) <= max_utts and (len(b) == 1 or sum(frames[i] for i in b) <= max_frames)
        flat = [frames[i] for b in batches for i in b]
        assert flat == sorted(flat, reverse=True)                                    # length-bucketed: longest first


@settings(max_examples=100, deadline=None)
@given(st.integers(400, 600000))
def test_frame_count_formula_is_monotone_and_matches_the_conv_arithmetic(n):
    from suta_b200 import ModelConfig
    c = ModelConfig.base()
    L = n
    for k, s in zip(c.conv_kernel, c.conv_stride):
        L = (L - k) // s + 1
    assert c.frames(n) == L == O.W2V2Config.base().frames(n)
    assert c.frames(n + 1) >= c.frames(n) and c.frames(n + 320) == c.frames(n) + 1   # one frame per 320 samples (20 ms)
