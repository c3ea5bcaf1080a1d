"""Per-segment error of the engine's step-0 gradient against fp32 autograd through the oracle (localises a wrong kernel).
Usage: python tools/grad_segments.py <fixture case>"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "test-time-adaptation-asr-suta_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import e2e_checks as E  # noqa: E402
from oracle import suta_oracle as O  # noqa: E402
from suta_b200 import AdaptHyper  # noqa: E402

case = sys.argv[1]
z, meta = E.load_golden(case)
ocfg, _ = E._cfgs(meta["cfg"])
sd = O.init_weights(ocfg, meta["weight_seed"], blank_bias=meta["blank_bias"], ln_jitter=meta["ln_jitter"])
wav = O.synth_audio(meta["n_samples"], meta["audio_seed"], meta.get("extra_noise", 0.0))
hp = AdaptHyper(**{k: meta["hyper"][k] for k in ("lr", "em_coef", "reweight", "temp", "not_blank")})
tf = bool(meta["train_feature"])
res = E.run_engine(meta["cfg"], sd, [wav], 1, hp, keep_grads=True, train_feature=tf, mult=E._mult(ocfg, tf))[0]
g0 = E.oracle_grad0(ocfg, sd, O.normalize_audio(wav), hp, meta["names"])
for name, off, size in res["segments"]:
    if name in g0:
        ref = res["to_layout"](name, torch.from_numpy(g0[name])).numpy().astype(np.float64)
        got = res["grad0"][off:off + size]
        print(f"{name:70s} rel {np.linalg.norm(got - ref) / (np.linalg.norm(ref) + 1e-30):.4f}  |ref| {np.linalg.norm(ref):.3e}")
