import sys, os, numpy as np, torch
ROOT='/root/repo'
sys.path[:0]=[ROOT, os.path.join(ROOT,'test-time-adaptation-asr-suta_b200')]
from suta_b200 import AdaptHyper, ModelConfig, SutaEngine
from suta_b200.weights import random_state_dict
from suta_b200.api import reference_multiplicities
cfg=ModelConfig.base()
rng=np.random.default_rng(1)
for mode in ('ln','feature'):
    mult=reference_multiplicities(cfg, train_feature=True) if mode=='feature' else None
    eng=SutaEngine(cfg, random_state_dict(cfg,0,1.75), train_feature=mode=='feature', trainable_mult=mult)
    for name, durs in (('one 0.5 s',[0.5]),('one 35 s',[35.0]),('100 short',list(rng.uniform(2,3,100))),('mixed 40',list(rng.uniform(0.3,20,40))),('two equal 128-frame',[2.5725,2.5725])):
        wavs=[(0.1*rng.standard_normal(int(d*16000))).astype(np.float32) for d in durs]
        eng.begin_batch(wavs); eng.reset(); eng.forward()
        hp=AdaptHyper()
        for _ in range(3): eng.adapt_step(hp)
        ids=eng.decode_ids(); torch.cuda.synchronize()
        lg=eng.tensor('logits') if hasattr(eng,'tensor') else None
        losses=eng.losses()[0].cpu().numpy()
        print(mode, name, 'frames', eng.total_frames, 'loss finite', bool(np.isfinite(losses).all()), 'loss[0]', float(losses[0]), 'ids0', len(ids[0]))
print('ok')
