"""The reference's per-utterance loop over the INSTALLED third-party stack -- TEST / BENCH INFRASTRUCTURE ONLY.

`/root/reference` is a script without packaging and does not travel to the GPU box, but everything its hot path
executes underneath does: HuggingFace `Wav2Vec2ForCTC` (transformers 5.5.0) and `torch.optim`.  This module drives
those real objects -- nn.Modules with requires_grad flags, autograd, torch.optim.AdamW/Adam/SGD with the duplicated
parameter list, state_dict snapshot + restore per utterance -- through a restatement of the ~60 lines of main.py that
sit on top (REF/main.py:8-23 setup_optimizer, :62-103 collect_params, :137-155 snapshot/restore, :172-215
forward_and_adapt, :319-402 the driver loop); the loss functions come from suta_oracle (REF/main.py:26-60).

Used by bench.py's `cpu_baseline`, `--impl reference` (device "cpu": the reference's CPU path on the box's host cores)
and `gpu_eager_baseline` (device "cuda": the reference's eager fp32 GPU path, the bar SURVEY.md 8d names) legs, and
by tests/test_oracle_golden.py, which checks it against the golden vectors the UNMODIFIED reference produced.
The product path (suta_b200) never imports it.
"""
from __future__ import annotations

import time
import warnings
from copy import deepcopy
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import suta_oracle as O


def build_model(cfg: O.W2V2Config, sd: Dict[str, torch.Tensor], device: str = "cpu"):
    """Wav2Vec2ForCTC(cfg) with the given weights, `.eval()` on `device` (REF/main.py:303)."""
    from transformers import Wav2Vec2ForCTC
    model = Wav2Vec2ForCTC(cfg.to_hf()).eval()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [m for m in missing if "masked_spec_embed" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    return model.to(device)


def collect_params(model, bias_only=False, train_feature=False, train_all=False, train_LN=True):
    """REF/main.py:62-103 (same walk, same duplicates; the per-module print is dropped)."""
    params, names = [], []
    trainable = ["bias"] if bias_only else ["weight", "bias"]
    for nm, m in model.named_modules():
        if train_LN and isinstance(m, torch.nn.LayerNorm):
            for pn, p in m.named_parameters():
                if pn in trainable:
                    p.requires_grad = True
                    params.append(p)
                    names.append(f"{nm}.{pn}")
        if train_feature and len(nm.split(".")) > 1 and nm.split(".")[1] in ("feature_extractor", "feature_projection"):
            for pn, p in m.named_parameters():
                p.requires_grad = True
                params.append(p)
                names.append(f"{nm}.{pn}")
        if train_all:
            for pn, p in m.named_parameters():
                p.requires_grad = True
                params.append(p)
                names.append(f"{nm}.{pn}")
    return params, names


class ReferenceLoop:
    """configure_model + collect_params + setup_optimizer + snapshot once; adapt() = one pass of REF/main.py:319-402."""

    def __init__(self, cfg: O.W2V2Config, sd: Dict[str, torch.Tensor], device: str = "cpu", train_feature: bool = False,
                 bias_only: bool = False, opt: str = "AdamW", lr: float = 2e-5, beta: float = 0.9,
                 sched_gamma: Optional[float] = None, train_all: bool = False):
        self.device = device
        self.model = build_model(cfg, sd, device)
        self.model.requires_grad_(False)                                       # configure_model, REF/main.py:167-170
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")                                    # torch warns about duplicate parameters
            self.params, self.names = collect_params(self.model, bias_only, train_feature, train_all)
            kw = dict(betas=(beta, 0.999)) if opt == "Adam" else {}            # REF/main.py:12-18
            self.optimizer = getattr(torch.optim, opt)(self.params, lr=lr, weight_decay=0.0, **kw)
        self.scheduler = (torch.optim.lr_scheduler.StepLR(self.optimizer, step_size=1, gamma=sched_gamma)
                          if sched_gamma is not None else None)                # REF/main.py:20-21
        self.model_state = deepcopy(self.model.state_dict())                   # REF/main.py:137-145
        self.optimizer_state = deepcopy(self.optimizer.state_dict())
        self.scheduler_state = deepcopy(self.scheduler.state_dict()) if self.scheduler is not None else None

    def reset(self):
        """load_model_and_optimizer, REF/main.py:147-155."""
        self.model.load_state_dict(self.model_state, strict=True)
        self.optimizer.load_state_dict(self.optimizer_state)
        if self.scheduler is not None:
            self.scheduler.load_state_dict(self.scheduler_state)

    def forward_and_adapt(self, x, em_coef, reweight, temp, not_blank, div_coef=0.0):
        """REF/main.py:172-215."""
        outputs = self.model(x).logits
        loss = O.suta_loss(outputs, em_coef, reweight, temp, not_blank, div_coef)
        loss.backward()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.optimizer.step()
        if self.scheduler is not None:
            self.scheduler.step()
        self.model.zero_grad()
        with torch.no_grad():
            outputs = self.model(x).logits
        return outputs, float(loss.detach())

    def adapt(self, x: np.ndarray, steps: int = 10, em_coef: float = 0.3, reweight: bool = True, temp: float = 2.5,
              not_blank: bool = True, div_coef: float = 0.0, episodic: bool = True) -> O.AdaptResult:
        """x: normalised waveform [N] (the processor's input_values).  Decodes at the reference's checkpoints."""
        xt = torch.from_numpy(np.asarray(x, dtype=np.float32))[None].to(self.device)
        if episodic:
            self.reset()
        res = O.AdaptResult(logits0=None)
        with torch.no_grad():
            res.logits0 = self.model(xt).logits[0].cpu().numpy()
        res.texts[0] = O.ctc_greedy_decode(res.logits0)
        for i in range(steps):
            out, loss = self.forward_and_adapt(xt, em_coef, reweight, temp, not_blank, div_coef)
            res.losses.append(loss)
            if (i + 1) in O.CHECKPOINT_STEPS or i + 1 == steps:
                res.logits[i + 1] = out[0].cpu().numpy()
            if (i + 1) in O.CHECKPOINT_STEPS:
                res.texts[i + 1] = O.ctc_greedy_decode(res.logits[i + 1])
        msd = self.model.state_dict()
        # (train_all lists the root module's parameters as ".<name>" and wav2vec2.masked_spec_embed, which the state dict of
        #  an eval-mode model may not carry: every real parameter is also listed under its full name)
        res.params = {n: msd[n].detach().cpu().numpy().copy() for n in dict.fromkeys(self.names) if n in msd}
        return res


def time_utterances(loop: ReferenceLoop, wavs: Sequence[np.ndarray], steps: int, warmup: int = 0, **hyper) -> List[float]:
    """Wall seconds per utterance (reset + vanilla forward + `steps` x forward_and_adapt + decodes), after `warmup`
    untimed utterances taken from the front of `wavs`."""
    out = []
    for i, w in enumerate(wavs):
        x = O.normalize_audio(w)
        if loop.device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        loop.adapt(x, steps=steps, **hyper)
        if loop.device != "cpu":
            torch.cuda.synchronize()
        if i >= warmup:
            out.append(time.perf_counter() - t0)
    return out
