"""--train_all at the reference's operating point: one utterance at a time (REF/main.py:319-402 with :96-100).

Times, per utterance of a duration-stratified sample of the LibriSpeech-shaped set: reset + vanilla forward + S x (loss,
backward, AdamW over every weight with the reference's multiplicities, forward) + the greedy decodes -- through this
repo's engine (wav2vec2-base, SUTA_FLAG_TRAIN_ALL) and through the reference's eager fp32 loop on the same GPU
(oracle/hf_reference.py over HF Wav2Vec2ForCTC + torch.optim.AdamW, `collect_params(train_all=True)`).  One JSON line.
The engine leg imports the product only; the oracle is imported by the baseline leg alone (as in bench.py's
`gpu_eager_baseline`), never on the measured product path.

    python tools/train_all_latency.py [--steps 10] [--utts 5] [--no-eager]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "test-time-adaptation-asr-suta_b200")]

from suta_b200 import AdaptHyper, ModelConfig, SutaEngine, api  # noqa: E402
from suta_b200.data import librispeech_shaped  # noqa: E402
from suta_b200.runner import adapt_batch  # noqa: E402
from suta_b200.text import CTCVocab  # noqa: E402
from suta_b200.weights import random_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--utts", type=int, default=5)
    ap.add_argument("--reps", type=int, default=3, help="timed repetitions per utterance (the median is reported)")
    ap.add_argument("--no-eager", action="store_true")
    a = ap.parse_args()
    utts = sorted(librispeech_shaped(2939), key=lambda u: u.n_samples)
    sample = [utts[int((i + 0.5) * len(utts) / a.utts)] for i in range(a.utts)]
    mcfg = ModelConfig.base()
    sd = random_state_dict(mcfg, seed=0, blank_bias=1.75)
    eng = SutaEngine(mcfg, sd, train_all=True, trainable_mult=api.reference_multiplicities(mcfg, train_all=True))
    hp, vocab = AdaptHyper(), CTCVocab()
    ms, secs = [], []
    for j, u in enumerate([sample[len(sample) // 2]] + sample):          # first = warm-up
        wav = u.audio()
        host = torch.zeros((len(wav) + 3) & ~3, dtype=torch.float32).pin_memory()       # the engine's packed layout (16-byte rows)
        host[:len(wav)] = torch.from_numpy(np.ascontiguousarray(wav, dtype=np.float32))
        ts = []
        for _ in range(a.reps if j else 1):              # a single shot is at the mercy of the host's scheduling noise
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            eng.begin_batch_lengths(np.asarray([len(wav)], dtype=np.int32))
            adapt_batch(eng, host, np.asarray([len(wav)]), a.steps, hp, vocab)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        if j:
            ms.append(float(np.median(ts)))
            secs.append(u.duration)
    out = {"what": "train_all, one utterance per step (REF/main.py:96-100), wav2vec2-base, %d-step SUTA" % a.steps,
           "utt_seconds": secs, "ms_per_utt": ms, "audio_s_per_s": sum(secs) / (sum(ms) * 1e-3),
           "rtf_median": float(np.median([m * 1e-3 / s for m, s in zip(ms, secs)])), "params_per_utt": int(eng.n_params),
           "launches": int(eng.launch_count)}
    eng.close()
    if not a.no_eager:
        from oracle import hf_reference as HR            # baseline leg only
        from oracle import suta_oracle as O
        ocfg = O.W2V2Config.base()
        # HF's positional conv is weight_norm-parametrised: g / v from the folded weight of random_state_dict
        pre = "wav2vec2.encoder.pos_conv_embed.conv."
        sd = dict(sd)
        w = sd.pop(pre + "weight")
        sd[pre + "parametrizations.weight.original0"] = w.norm(dim=(0, 1), keepdim=True)
        sd[pre + "parametrizations.weight.original1"] = w.clone()
        loop = HR.ReferenceLoop(ocfg, sd, "cuda", train_all=True)
        t = HR.time_utterances(loop, [sample[len(sample) // 2].audio()] + [u.audio() for u in sample], steps=a.steps, warmup=1)
        out["gpu_eager_reference"] = {"s_per_utt": t, "audio_s_per_s": sum(secs) / sum(t),
                                      "kind": "oracle/hf_reference.py ReferenceLoop(train_all=True) on cuda:0, fp32 eager"}
        out["speedup_vs_eager"] = out["audio_s_per_s"] / out["gpu_eager_reference"]["audio_s_per_s"]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
