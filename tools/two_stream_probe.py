"""Experiment: two adaptation batches in flight on two CUDA streams (one engine + one host thread each) vs one after the
other.  Kernels of one batch's dependency chain cannot overlap each other, but the tail of a tensor-bound GEMM of batch A
can overlap the ramp-up of a memory-bound kernel of batch B.  Usage: python tools/two_stream_probe.py [n_batches] [mode]"""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "test-time-adaptation-asr-suta_b200")]
import torch  # noqa: E402

from suta_b200 import AdaptHyper, ModelConfig, SutaEngine  # noqa: E402
from suta_b200.api import reference_multiplicities  # noqa: E402
from suta_b200.data import librispeech_shaped  # noqa: E402
from suta_b200.runner import SutaRunner, adapt_batch  # noqa: E402
from suta_b200.text import CTCVocab  # noqa: E402
from suta_b200.weights import random_state_dict  # noqa: E402
import bench  # noqa: E402

NB = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tf = (sys.argv[2] if len(sys.argv) > 2 else "feature") == "feature"
cfg = ModelConfig.base()
sd = random_state_dict(cfg, 0, 1.75)
mult = reference_multiplicities(cfg, train_feature=True) if tf else None
utts = librispeech_shaped(2939, seed=0)
idx = bench.select_batches(utts, cfg, NB, 64, 36864)
engines = [SutaEngine(cfg, sd, train_feature=tf, trainable_mult=mult) for _ in range(2)]
runners = [SutaRunner(e, 10, AdaptHyper(), vocab=CTCVocab()) for e in engines]
staged = [runners[i % 2].stage(utts, [b])[0] for i, b in enumerate(idx)]
audio = sum(utts[j].duration for b in idx for j in b)
hp, vocab = AdaptHyper(), CTCVocab()


def work(k, items, stream):
    with torch.cuda.stream(stream):
        for b, lens, packed in items:
            engines[k].begin_batch_lengths(lens)
            adapt_batch(engines[k], packed, lens, 10, AdaptHyper(), vocab)


def run(two):
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if two:
        th = [threading.Thread(target=work, args=(k, staged[k::2], streams[k])) for k in range(2)]
        for t in th:
            t.start()
        for t in th:
            t.join()
    else:
        work(0, staged[0::2], streams[0])
        work(1, staged[1::2], streams[1])
    torch.cuda.synchronize()
    return time.perf_counter() - t0


for two in (False, True, False, True):
    dt = run(two)
    print(f"{'two streams' if two else 'sequential '}: {dt * 1e3:.1f} ms for {NB} batches -> {audio / dt:.0f} audio-s/s", flush=True)
