"""--train_all on the longest and shortest utterances the pipeline admits (35 s, the 600 000-sample clamp of REF/data.py:19-21,
0.125 s): many M-blocks / k-blocks in every weight-gradient GEMM, graph chains at T ~ 1 900.  Prints frames, finiteness of
logits and gradients, the losses of 3 steps and the graph-replay count per utterance (product imports only).

    python tools/train_all_long_probe.py
"""
import sys, numpy as np, torch
sys.path[:0]=['.','test-time-adaptation-asr-suta_b200']
from suta_b200 import AdaptHyper, ModelConfig, SutaEngine, api
from suta_b200.weights import random_state_dict
from suta_b200.runner import adapt_batch
from suta_b200.text import CTCVocab
mcfg=ModelConfig.base(); sd=random_state_dict(mcfg,seed=0,blank_bias=1.75)
eng=SutaEngine(mcfg,sd,train_all=True,trainable_mult=api.reference_multiplicities(mcfg,train_all=True))
rng=np.random.default_rng(0)
for n in (560000, 16000*35, 600000, 2000):
    wav=(0.1*rng.standard_normal(n)).astype(np.float32)
    eng.begin_batch([wav])
    host=torch.zeros(eng.total_samples); host[:n]=torch.from_numpy(wav)
    t=adapt_batch(eng,host.pin_memory(),eng.lengths,3,AdaptHyper(),CTCVocab(),collect_losses=True)
    lg=eng.logits(); g=eng.grads()
    print(n, int(eng.total_frames), bool(torch.isfinite(lg).all()), bool(torch.isfinite(g).all()), [float(x[0]) for x in t["losses"]], float(g.abs().max()), eng.graph_replays)
