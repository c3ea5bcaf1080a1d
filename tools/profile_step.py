"""Small fixed workload for ncu: one adaptation batch (default 24 utterances x 6 s, 2 SUTA steps) so that kernel
replay does not have to save/restore tens of GB.  Usage: python tools/profile_step.py [--utts N] [--seconds S] [--mode ln|feature]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "test-time-adaptation-asr-suta_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

from suta_b200 import AdaptHyper, ModelConfig, SutaEngine  # noqa: E402
from suta_b200.weights import random_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--utts", type=int, default=24)
ap.add_argument("--seconds", type=float, default=6.0)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--mode", default="ln")
ap.add_argument("--model", default="base", help="base | large | large_lv60 | tiny ...")
a = ap.parse_args()
cfg = getattr(ModelConfig, a.model)()
mult = None
if a.mode == "feature":
    sys.path.insert(0, ROOT)
    from suta_b200.api import reference_multiplicities
    mult = reference_multiplicities(cfg, train_feature=True)
eng = SutaEngine(cfg, random_state_dict(cfg, 0, 1.75), train_feature=a.mode == "feature", trainable_mult=mult)
rng = np.random.default_rng(0)
wavs = [(0.1 * rng.standard_normal(int(a.seconds * 16000) + 160 * i)).astype(np.float32) for i in range(a.utts)]
eng.begin_batch(wavs)
eng.reset()
eng.forward()
hp = AdaptHyper()
for _ in range(a.steps):
    eng.adapt_step(hp)
ids = eng.decode_ids()
torch.cuda.synchronize()
print("frames", eng.total_frames, "launches", eng.launch_count, "first ids", ids[0][:8])
