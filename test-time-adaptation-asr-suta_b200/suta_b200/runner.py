"""Batched, sharded SUTA over a list of utterances: the reference's outer loop (REF/main.py:319-417) re-designed
for one process per GPU.  Each rank takes an LPT shard of the utterance list, adapts it in length-bucketed batches on
its own engine (no collective inside the loop), and only transcripts + integer WER counts are exchanged at the end."""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .data import Utterance
from .engine import CHECKPOINT_STEPS, AdaptHyper, SutaEngine
from .shard import bucket_batches, shard_lpt, utterance_cost
from .text import CTCVocab
from .wer import wer_counts


@dataclass
class BatchResult:
    indices: List[int]
    texts: Dict[int, List[str]]                     # checkpoint step (0 = original) -> transcript per utterance
    losses: List[np.ndarray] = field(default_factory=list)
    audio_seconds: float = 0.0


def adapt_batch(engine: SutaEngine, wavs_packed: torch.Tensor, lengths: np.ndarray, steps: int, hp: AdaptHyper,
                vocab: CTCVocab, episodic: bool = True, collect_losses: bool = False,
                sched_gamma: Optional[float] = None, sched_step: int = 1, extra_noise: float = 0.0, noise_seed: int = 0,
                utt_ids: Optional[Sequence[int]] = None) -> Dict[int, List[str]]:
    """One adaptation batch = the body of REF/main.py:319-402 for len(lengths) utterances at once.
    `wavs_packed` is the packed waveform buffer (pinned host or device) laid out by engine.begin_batch_lengths.
    Batches are always episodic: every utterance starts from the pristine parameters (REF/main.py:327-328) -- carrying
    state from utterance to utterance serialises them (that mode lives in api.SutaModel, batch size 1).
    sched_gamma / sched_step: the StepLR of REF/main.py:20-21, restarted for every utterance like the reference's
    scheduler.load_state_dict (:151-153).
    extra_noise: REF/data.py:23, added to the raw waveform ON THE DEVICE (keyed by noise_seed and utt_ids)."""
    if not episodic:
        raise ValueError("adapt_batch adapts independent utterances: episodic only (use the api.* surface for continual mode)")
    engine.set_audio(wavs_packed)
    if extra_noise > 0:
        engine.add_noise(extra_noise, noise_seed, utt_ids)
    engine.reset()
    engine.forward()                                               # vanilla forward, REF/main.py:331-334
    texts = {0: vocab.batch_to_text(engine.decode_ids())}
    losses = []
    lr0 = hp.lr
    try:
        for i in range(steps):                                     # REF/main.py:347-348
            if sched_gamma is not None:
                hp.lr = lr0 * sched_gamma ** (i // sched_step)
            engine.adapt_step(hp)
            if collect_losses:
                losses.append(engine.losses()[0].cpu().numpy().copy())
            if (i + 1) in CHECKPOINT_STEPS:                        # REF/main.py:349-398
                texts[i + 1] = vocab.batch_to_text(engine.decode_ids())
    finally:
        hp.lr = lr0
    if collect_losses:
        texts["losses"] = losses
    return texts


def pack_batch(engine: SutaEngine, utts: Sequence[Utterance], pinned: bool = True, with_noise: bool = True) -> torch.Tensor:
    lens = np.asarray([u.n_samples for u in utts], dtype=np.int32)
    engine.begin_batch_lengths(lens)
    host = torch.zeros(engine.total_samples, dtype=torch.float32)
    if pinned:
        host = host.pin_memory()
    hv = host.numpy()
    for u, o in zip(utts, engine.sample_off):
        hv[o:o + u.n_samples] = u.audio(with_noise)
    return host


class SutaRunner:
    """Dataset-level driver for one rank."""

    def __init__(self, engine: SutaEngine, steps: int = 10, hp: Optional[AdaptHyper] = None, max_utts: int = 64,
                 max_frames: int = 32768, vocab: Optional[CTCVocab] = None, rank: int = 0, world_size: int = 1,
                 sched_gamma: Optional[float] = None, sched_step: int = 1, extra_noise: float = 0.0, noise_seed: int = 0):
        self.engine, self.steps, self.hp = engine, steps, hp or AdaptHyper()
        self.sched_gamma, self.sched_step = sched_gamma, sched_step
        # REF/data.py:23: the runner stages CLEAN waveforms and the engine adds the noise on the device, keyed by the
        # utterance's index in the set (the same noise whichever rank / batch adapts it)
        self.extra_noise, self.noise_seed = extra_noise, noise_seed
        self.max_utts, self.max_frames = max_utts, max_frames
        self.vocab = vocab or CTCVocab()
        self.rank, self.world_size = rank, world_size

    def plan(self, utts: Sequence[Utterance]) -> List[List[int]]:
        cfg = self.engine.cfg
        frames = [cfg.frames(u.n_samples) for u in utts]
        costs = [utterance_cost(f, cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers, self.steps) for f in frames]
        mine = shard_lpt(costs, self.world_size)[self.rank]
        return bucket_batches(frames, mine, self.max_utts, self.max_frames)

    def stage(self, utts: Sequence[Utterance], batches: Optional[List[List[int]]] = None, device: bool = False):
        """Materialise the packed waveform buffer of every batch of this rank's plan (pinned host memory, or device
        memory with device=True): [(indices, lengths, packed)].  Data loading / synthesis is not part of the adaptation."""
        out = []
        for b in (self.plan(utts) if batches is None else batches):
            sel = [utts[i] for i in b]
            packed = pack_batch(self.engine, sel, with_noise=False)
            out.append((b, np.asarray([u.n_samples for u in sel], dtype=np.int32), packed.to(self.engine.device) if device else packed))
        return out

    def run(self, utts: Sequence[Utterance], staged=None) -> Dict[str, object]:
        """Adapt this rank's shard batch by batch (REF/main.py:319-402) and score it (REF/main.py:405-417).
        `staged` = the result of stage(): skips building the waveform buffers."""
        if staged is None:
            staged = self.stage(utts)
        texts: Dict[int, Dict[int, str]] = {}
        t0 = time.time()
        for b, lens, packed in staged:
            self.engine.begin_batch_lengths(lens)
            out = adapt_batch(self.engine, packed, lens, self.steps, self.hp, self.vocab, sched_gamma=self.sched_gamma,
                              sched_step=self.sched_step, extra_noise=self.extra_noise, noise_seed=self.noise_seed, utt_ids=b)
            for step, tl in out.items():
                texts.setdefault(step, {}).update({i: t for i, t in zip(b, tl)})
        torch.cuda.synchronize()
        wall = time.time() - t0
        counts = {}
        for step, d in texts.items():
            idx = sorted(d)
            counts[step] = wer_counts([utts[i].text for i in idx], [d[i] for i in idx])
        return dict(texts=texts, wer_counts=counts, wall_s=wall,
                    audio_s=sum(utts[i].duration for b, _l, _p in staged for i in b), n_batches=len(staged))


def gather_results(local: Dict[str, object], steps_keys: Sequence[int]) -> Dict[str, object]:
    """End-of-run exchange: all_reduce(SUM) of the integer WER counters + gather of transcripts.
    Works with any initialised torch.distributed backend (nccl on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(local, wer={k: (e / n if n else float("nan")) for k, (e, n) in local["wer_counts"].items()})
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    vec = torch.zeros(2 * len(steps_keys), dtype=torch.int64, device=dev)
    for j, k in enumerate(steps_keys):
        e, n = local["wer_counts"].get(k, (0, 0))
        vec[2 * j], vec[2 * j + 1] = e, n
    dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    gathered = [None] * dist.get_world_size()
    dist.all_gather_object(gathered, local["texts"])
    texts: Dict[int, Dict[int, str]] = {}
    for part in gathered:
        for step, d in part.items():
            texts.setdefault(step, {}).update(d)
    v = vec.cpu().tolist()
    counts = {k: (v[2 * j], v[2 * j + 1]) for j, k in enumerate(steps_keys)}
    return dict(texts=texts, wer_counts=counts, wer={k: (e / n if n else float("nan")) for k, (e, n) in counts.items()},
                wall_s=local["wall_s"], audio_s=local["audio_s"])
