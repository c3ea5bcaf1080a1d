"""Drop-in mirror of the reference's adaptation API (REF/main.py:8-215): same names, positional order, defaults and
error behaviour, backed by the CUDA engine instead of torch autograd over HF modules.

    model = SutaModel(cfg, state_dict)                 # replaces Wav2Vec2ForCTC.from_pretrained(...).eval().cuda()
    model = configure_model(model)
    params, names = collect_params(model, bias_only, train_feature, train_all, train_LN)
    optimizer, scheduler = setup_optimizer(params, 'AdamW', lr)
    state = copy_model_and_optimizer(model, optimizer, scheduler)
    model, optimizer, scheduler = load_model_and_optimizer(model, optimizer, *state)
    logits = forward_and_adapt(x, model, optimizer, em_coef, reweight, temp, not_blank, scheduler, div_coef)

`x` is what the reference passes: the processor's normalised input_values, fp32 [B, N].  Unlike the reference (whose
mcc_loss only works for B == 1, REF/main.py:32,39) every row of x is adapted as an independent utterance with its own
parameters -- for B == 1 the semantics are identical.
"""
from __future__ import annotations

import ctypes as C
from copy import deepcopy
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check
from .config import ModelConfig
from .engine import AdaptHyper, SutaEngine, _stream_ptr


# --------------------------------------------------------------------------------------------------
# model / parameter / optimizer objects
# --------------------------------------------------------------------------------------------------
class _Out:
    def __init__(self, logits):
        self.logits = logits


class SutaParam:
    """Handle to one trainable tensor (all utterances of the live batch): what collect_params returns."""

    def __init__(self, model: "SutaModel", name: str, offset: int, size: int):
        self.model, self.name, self.offset, self.size = model, name, offset, size
        self.requires_grad = False

    @property
    def data(self) -> torch.Tensor:            # [B, size] view of the adapted values
        return self.model.engine.params()[:, self.offset:self.offset + self.size]

    @property
    def grad(self) -> torch.Tensor:
        return self.model.engine.grads()[:, self.offset:self.offset + self.size]

    def __repr__(self):
        return f"SutaParam({self.name}, size={self.size})"


def _module_order(cfg: ModelConfig) -> List[Tuple[str, str, List[str]]]:
    """(module name, kind, own parameter leaves) in HF named_modules() order, trainable-relevant modules only."""
    mods = [("wav2vec2.feature_extractor", "container", []), ("wav2vec2.feature_extractor.conv_layers", "container", [])]
    for i in range(len(cfg.conv_dim)):
        b = f"wav2vec2.feature_extractor.conv_layers.{i}"
        mods += [(b, "container", []), (b + ".conv", "conv", ["weight", "bias"] if cfg.conv_bias else ["weight"])]
        if cfg.feat_extract_norm == "layer":       # every conv layer carries a LayerNorm(C) (HF/modeling_wav2vec2.py:288)
            mods.append((b + ".layer_norm", "layernorm", ["weight", "bias"]))
        elif i == 0:
            mods.append((b + ".layer_norm", "groupnorm", ["weight", "bias"]))
    mods += [("wav2vec2.feature_projection", "container", []),
             ("wav2vec2.feature_projection.layer_norm", "layernorm", ["weight", "bias"]),
             ("wav2vec2.feature_projection.projection", "linear", ["weight", "bias"]),
             ("wav2vec2.encoder.layer_norm", "layernorm", ["weight", "bias"])]
    for l in range(cfg.num_hidden_layers):
        b = f"wav2vec2.encoder.layers.{l}"
        mods += [(b + ".layer_norm", "layernorm", ["weight", "bias"]), (b + ".final_layer_norm", "layernorm", ["weight", "bias"])]
    return mods


def _module_tree_all(cfg: ModelConfig) -> List[Tuple[str, str, List[str]]]:
    """EVERY module of HF's Wav2Vec2ForCTC in named_modules() order with its own parameter leaves: what --train_all
    walks (REF/main.py:96-100).  Both families (the module ORDER of the two encoders is the same; the LayerNorm feature
    extractor registers conv, layer_norm, activation per layer, HF:281-289).  tests/test_host.py checks it against the real
    HF module trees."""
    if (cfg.feat_extract_norm == "layer") != bool(cfg.do_stable_layer_norm):
        raise NotImplementedError("--train_all: GroupNorm + post-LN or LayerNorm feature extractor + pre-LN (the two shipped families)")
    mods = [("", "container", []), ("wav2vec2", "container", ["masked_spec_embed"]),
            ("wav2vec2.feature_extractor", "container", []), ("wav2vec2.feature_extractor.conv_layers", "container", [])]
    for i in range(len(cfg.conv_dim)):
        b = f"wav2vec2.feature_extractor.conv_layers.{i}"
        mods += [(b, "container", []), (b + ".conv", "conv", ["weight", "bias"] if cfg.conv_bias else ["weight"])]
        if cfg.feat_extract_norm == "layer":
            mods += [(b + ".layer_norm", "layernorm", ["weight", "bias"]), (b + ".activation", "none", [])]
            continue
        mods.append((b + ".activation", "none", []))
        if i == 0:
            mods.append((b + ".layer_norm", "groupnorm", ["weight", "bias"]))
    e = "wav2vec2.encoder"
    mods += [("wav2vec2.feature_projection", "container", []),
             ("wav2vec2.feature_projection.layer_norm", "layernorm", ["weight", "bias"]),
             ("wav2vec2.feature_projection.projection", "linear", ["weight", "bias"]),
             ("wav2vec2.feature_projection.dropout", "none", []),
             (e, "container", []), (e + ".pos_conv_embed", "container", []), (e + ".pos_conv_embed.conv", "conv", ["bias"]),
             (e + ".pos_conv_embed.conv.parametrizations", "container", []),
             (e + ".pos_conv_embed.conv.parametrizations.weight", "container", ["original0", "original1"]),
             (e + ".pos_conv_embed.conv.parametrizations.weight.0", "none", []),
             (e + ".pos_conv_embed.padding", "none", []), (e + ".pos_conv_embed.activation", "none", []),
             (e + ".layer_norm", "layernorm", ["weight", "bias"]), (e + ".dropout", "none", []), (e + ".layers", "container", [])]
    for l in range(cfg.num_hidden_layers):
        b = f"{e}.layers.{l}"
        mods += [(b, "container", []), (b + ".attention", "container", [])]
        mods += [(b + f".attention.{n}_proj", "linear", ["weight", "bias"]) for n in ("k", "v", "q", "out")]
        mods += [(b + ".dropout", "none", []), (b + ".layer_norm", "layernorm", ["weight", "bias"]),
                 (b + ".feed_forward", "container", []), (b + ".feed_forward.intermediate_dropout", "none", []),
                 (b + ".feed_forward.intermediate_dense", "linear", ["weight", "bias"]),
                 (b + ".feed_forward.intermediate_act_fn", "none", []),
                 (b + ".feed_forward.output_dense", "linear", ["weight", "bias"]),
                 (b + ".feed_forward.output_dropout", "none", []), (b + ".final_layer_norm", "layernorm", ["weight", "bias"])]
    mods += [("dropout", "none", []), ("lm_head", "linear", ["weight", "bias"])]
    return mods


def _walk_all(cfg: ModelConfig, bias_only: bool, train_feature: bool, train_all: bool, train_LN: bool):
    """REF/main.py:62-103 over the full module tree: yields (module name, listed name, full parameter name) in the
    reference's order (module name None marks the visit of a module, for the reference's print)."""
    trainable = ['bias'] if bias_only else ['weight', 'bias']
    tree = _module_tree_all(cfg)

    def below(nm):          # m.named_parameters(): own parameters first, then the sub-modules', pre-order
        for nm2, _k, leaves in tree:
            if nm == "" or nm2 == nm or nm2.startswith(nm + "."):
                for leaf in leaves:
                    full = f"{nm2}.{leaf}" if nm2 else leaf
                    yield (full[len(nm) + 1:] if nm else full), full

    for nm, kind, leaves in tree:
        yield nm, None, None
        if train_LN and kind == "layernorm":
            for leaf in leaves:
                if leaf in trainable:
                    yield nm, f"{nm}.{leaf}", f"{nm}.{leaf}"
        if train_feature and len(nm.split('.')) > 1 and nm.split('.')[1] in ('feature_extractor', 'feature_projection'):
            for rel, full in below(nm):
                yield nm, f"{nm}.{rel}", full
        if train_all:
            for rel, full in below(nm):
                yield nm, f"{nm}.{rel}", full


def reference_multiplicities(cfg, bias_only: bool = False, train_feature: bool = False, train_LN: bool = True,
                             train_all: bool = False) -> Dict[str, int]:
    """How many times REF/main.py:62-103 would list each parameter (0 = not trainable): LayerNorm affine once, and under
    train_feature every parameter of feature_extractor / feature_projection once per enclosing module (recursive
    named_parameters), e.g. conv weights x4, feature_projection.layer_norm x3, projection x2."""
    cfg = ModelConfig.from_any(cfg)
    if train_all:           # every parameter once per enclosing module: up to 7 for an encoder Linear or LayerNorm
        mult = {}
        for _nm, listed, full in _walk_all(cfg, bias_only, train_feature, True, train_LN):
            if listed is not None:
                mult[full] = mult.get(full, 0) + 1
        return mult
    trainable = ['bias'] if bias_only else ['weight', 'bias']
    tree = _module_order(cfg)
    mult: Dict[str, int] = {f"{nm}.{leaf}": 0 for nm, _k, leaves in tree for leaf in leaves}
    for nm, kind, leaves in tree:
        if train_LN and kind == "layernorm":
            for leaf in leaves:
                if leaf in trainable:
                    mult[f"{nm}.{leaf}"] += 1
        if train_feature and len(nm.split('.')) > 1 and nm.split('.')[1] in ('feature_extractor', 'feature_projection'):
            for nm2, _k2, leaves2 in tree:
                if nm2 == nm or nm2.startswith(nm + "."):
                    for leaf in leaves2:
                        mult[f"{nm2}.{leaf}"] += 1
    return mult


class SutaModel:
    """Engine-backed stand-in for the HF model object the reference functions touch."""

    def __init__(self, cfg, state_dict: Dict[str, torch.Tensor], train_feature: bool = False, device=None,
                 pseudo_label: bool = False, train_all: bool = False):
        self.cfg = ModelConfig.from_any(cfg)
        self.engine = SutaEngine(self.cfg, state_dict, train_feature=train_feature, trainable_mult={}, device=device,
                                 pseudo_label=pseudo_label, train_all=train_all)
        self._params = {name: SutaParam(self, name, off, size) for name, off, size in self.engine.segments}
        if train_all:       # listed by the reference's walk, never given a gradient (eval mode: no SpecAugment mask), so
            #                 torch.optim skips it: an empty handle keeps the parameter list and its names identical
            self._params["wav2vec2.masked_spec_embed"] = SutaParam(self, "wav2vec2.masked_spec_embed", 0, 0)
        self._x_ref = None              # strong reference to the bound input: its address cannot be recycled while bound
        self._x_version = -1
        self._logits_valid = False
        self._shape = None
        # model + optimizer state carried from one bound input to the next (REF/main.py:319-348 without --episodic):
        # P = trainable vector, m / v = Adam moments (None = empty optimizer state), steps = optimizer.step() count
        self._carry = dict(P=None, m=None, v=None, steps=0)

    # ---- what REF/main.py calls on `model` ----
    def eval(self):
        return self

    def cuda(self, *a, **k):
        return self

    def requires_grad_(self, flag: bool = False):
        for p in self._params.values():
            p.requires_grad = bool(flag)
        self.engine.set_trainable({n: int(flag) for n in self._params})
        return self

    def zero_grad(self, set_to_none: bool = True):
        return None                     # gradients are re-zeroed at the start of every backward

    def named_parameters(self):
        return list(self._params.items())

    def state_dict(self) -> Dict[str, torch.Tensor]:
        if self.engine.n_utts:
            return {"trainable": self.engine.params()[0].clone()}
        p = self._carry["P"]
        return {"trainable": (self.engine.params0 if p is None else p).clone()}

    def load_state_dict(self, state: Dict[str, torch.Tensor], strict: bool = True):
        if strict and set(state) != {"trainable"}:
            raise RuntimeError(f"unexpected keys in state_dict: {sorted(state)}")
        p = state["trainable"].to(self.engine.device)
        if self.engine.n_utts:
            self.engine.params().copy_(p[None].expand(self.engine.n_utts, -1))
            self.engine.params_written()
        self._carry["P"] = p.clone()
        self._logits_valid = False

    def _clear_optimizer_state(self):
        """optimizer.load_state_dict(<empty state>): both moments and the step count start over (REF/main.py:150)."""
        self._carry.update(m=None, v=None, steps=0)
        eng = self.engine
        if eng.n_utts:
            eng.exp_avg().zero_()
            eng.exp_avg_sq().zero_()
            eng.opt_steps = 0

    def _bind(self, x: torch.Tensor):
        if x.dim() != 2:
            raise ValueError("input_values must be [batch, samples]")
        # identity of the tensor OBJECT (kept alive by _x_ref) + its version counter; never the storage address: the
        # caching allocator hands the block of a deleted input to the next one of the same length (REF/main.py:400-401)
        if x is self._x_ref and x._version == self._x_version:
            return
        B, N = x.shape
        eng = self.engine
        c = self._carry
        if eng.n_utts:                     # the live batch IS the model: carry its state into the next input
            if eng.n_utts > 1 and (c["steps"] or eng.opt_steps):
                raise NotImplementedError("carrying adapted state across inputs (no --episodic) needs batch size 1, like "
                                          "the reference (REF/main.py:32); restore a snapshot first or use B == 1")
            c.update(P=eng.params()[0].clone(), m=eng.exp_avg()[0].clone(), v=eng.exp_avg_sq()[0].clone(), steps=eng.opt_steps)
        eng.begin_batch_lengths(np.full(B, N, dtype=np.int32))
        packed = torch.zeros(eng.total_samples, dtype=torch.float32, device=eng.device)
        xs = x.detach().to(device=eng.device, dtype=torch.float32)
        for u in range(B):
            o = int(eng.sample_off[u])
            packed[o:o + N] = xs[u]
        eng.set_audio(packed, normalized=True)
        # the new input starts from the model's CURRENT parameters and optimizer state (continual mode, the reference's
        # default) -- episodic callers restored the pristine snapshot through load_model_and_optimizer right before
        eng.reset()
        if c["P"] is not None:
            eng.params().copy_(c["P"][None].expand(B, -1))
            eng.params_written()
        if c["m"] is not None:
            eng.exp_avg().copy_(c["m"][None].expand(B, -1))
            eng.exp_avg_sq().copy_(c["v"][None].expand(B, -1))
        eng.opt_steps = c["steps"]
        self._x_ref, self._x_version, self._logits_valid = x, x._version, False
        self._shape = (B, int(eng.frames[0]), self.cfg.vocab_size)

    def __call__(self, x: torch.Tensor) -> _Out:
        self._bind(x)
        if not self._logits_valid:
            self.engine.forward()
            self._logits_valid = True
        return _Out(self.engine.logits().view(*self._shape))


class SutaOptimizer:
    """torch.optim-like object returned by setup_optimizer; the update itself is csrc/optim.cu."""

    def __init__(self, params: List[SutaParam], opt_name: str, lr: float, betas=(0.9, 0.999), weight_decay: float = 0.0):
        if not params:
            raise ValueError("optimizer got an empty parameter list")
        self.model = params[0].model
        mult: Dict[str, int] = {}
        for p in params:                                 # duplicates are kept: REF/main.py:88-94 semantics
            mult[p.name] = mult.get(p.name, 0) + 1
        self.mult = mult
        self.model.engine.set_trainable(mult)
        self.hp = AdaptHyper(opt=opt_name, lr=lr, beta1=betas[0], beta2=betas[1], weight_decay=weight_decay)
        self.model._optimizer = self
        self.param_groups = [dict(lr=lr, betas=betas, weight_decay=weight_decay, params=params)]

    def step(self):
        self.hp.lr = self.param_groups[0]["lr"]
        self.model.engine.optimizer_step(self.hp)
        self.model._logits_valid = False

    def zero_grad(self, set_to_none: bool = True):
        return None

    def state_dict(self):
        return {"state": {}, "param_groups": [{k: v for k, v in self.param_groups[0].items() if k != "params"}]}

    def load_state_dict(self, sd):
        # the reference only ever reloads the pre-adaptation snapshot (empty state): moments and step are cleared
        if sd.get("state"):
            raise NotImplementedError("only pristine (empty-state) optimizer snapshots are supported")
        self.param_groups[0].update(sd["param_groups"][0])
        self.model._clear_optimizer_state()


class _StepLR:
    def __init__(self, optimizer: SutaOptimizer, step_size=1, gamma=0.7):
        self.opt, self.step_size, self.gamma, self.n = optimizer, step_size, gamma, 0
        self.base = optimizer.param_groups[0]["lr"]

    def step(self):
        self.n += 1
        self.opt.param_groups[0]["lr"] = self.base * self.gamma ** (self.n // self.step_size)

    def state_dict(self):
        return {"n": self.n, "base": self.base}

    def load_state_dict(self, sd):
        self.n, self.base = sd["n"], sd["base"]
        self.opt.param_groups[0]["lr"] = self.base * self.gamma ** (self.n // self.step_size)


# --------------------------------------------------------------------------------------------------
# the reference's functions
# --------------------------------------------------------------------------------------------------
def setup_optimizer(params, opt_name='AdamW', lr=1e-4, beta=0.9, weight_decay=0., scheduler=None, step_size=1, gamma=0.7):
    """REF/main.py:8-23."""
    if opt_name not in ("AdamW", "Adam", "SGD"):
        raise AttributeError(f"module 'torch.optim' has no attribute '{opt_name}' supported by suta_b200")
    print(f'[INFO]    optimizer: {opt_name}')
    print(f'[INFO]    scheduler: {scheduler}')
    betas = (beta, 0.999) if opt_name == 'Adam' else (0.9, 0.999)
    optimizer = SutaOptimizer(params, opt_name, lr, betas, weight_decay)
    if scheduler is not None:
        if "StepLR" not in str(scheduler):
            raise NotImplementedError("only torch.optim.lr_scheduler.StepLR is supported")
        return optimizer, _StepLR(optimizer, step_size=step_size, gamma=gamma)
    return optimizer, None


def _as_rows(x: torch.Tensor) -> torch.Tensor:
    if x.shape[-1] != 32:
        raise ValueError("suta_b200 losses are specialised to the 32-symbol CTC vocabulary (class_num=32)")
    return x.detach().to(torch.float32).contiguous().view(-1, 32)


def softmax_entropy(x, dim=2):
    """REF/main.py:26-28: entropy of softmax over the class axis (must be the last axis)."""
    if dim not in (-1, x.dim() - 1):
        raise ValueError("softmax_entropy: only the class (last) axis is supported")
    rows = _as_rows(x)
    out = torch.empty(rows.shape[0], dtype=torch.float32, device=rows.device)
    check(_lib.load().suta_op_softmax_entropy(C.c_void_p(rows.data_ptr()), rows.shape[0], 1.0, C.c_void_p(out.data_ptr()),
                                             _stream_ptr()))
    return out.view(x.shape[:-1])


def mcc_loss(x, reweight=False, dim=2, class_num=32):
    """REF/main.py:30-44 (value only; its gradient is part of the fused kernel used by forward_and_adapt)."""
    if x.dim() != 3 or x.shape[0] != 1:
        raise RuntimeError("mcc_loss expects logits of shape [1, L, 32] (the reference squeezes dim 0)")
    rows = _as_rows(x)
    L = rows.shape[0]
    off = torch.zeros(1, dtype=torch.int64, device=rows.device)
    T = torch.full((1,), L, dtype=torch.int32, device=rows.device)
    loss = torch.empty(3, dtype=torch.float32, device=rows.device)
    p = lambda t: C.c_void_p(t.data_ptr())
    check(_lib.load().suta_op_loss(p(rows), p(off), p(T), 1, 0.0, 1.0, int(bool(reweight)), 0, 0.0, p(loss), None, None,
                                   _stream_ptr()))
    return loss[2] * (32.0 / class_num)


def div_loss(x, non_blank=None, L_thd=64):
    """REF/main.py:46-60: minus the entropy of softmax(mean-over-time logits), blank column dropped whenever
    `non_blank is not None` (value only: one 32-vector; the gradient used by forward_and_adapt is in the fused kernel)."""
    x = x.squeeze(0)
    cls_pred = x.mean(0)[1:] if non_blank is not None else x.mean(0)
    p = torch.softmax(cls_pred, 0)
    return (p * torch.log_softmax(cls_pred, 0)).sum()


def collect_params(model: SutaModel, bias_only=False, train_feature=False, train_all=False, train_LN=True):
    """REF/main.py:62-103: same walk over named_modules(), same duplicates, same names."""
    if train_all:
        if not model.engine.train_all:
            raise _lib.SutaError("model was not created with train_all=True")
        params, names = [], []
        for nm, listed, full in _walk_all(model.cfg, bias_only, train_feature, True, train_LN):
            if listed is None:
                print(nm)
                continue
            p = model._params[full]
            p.requires_grad = True
            params.append(p)
            names.append(listed)
        return params, names
    if train_feature and not model.engine.train_feature:
        raise _lib.SutaError("model was not created with train_feature=True")
    trainable = ['bias'] if bias_only else ['weight', 'bias']
    params, names = [], []
    tree = _module_order(model.cfg)
    for nm, kind, leaves in tree:
        print(nm)
        if train_LN and kind == "layernorm":
            for leaf in leaves:
                if leaf in trainable:
                    p = model._params[f"{nm}.{leaf}"]
                    p.requires_grad = True
                    params.append(p)
                    names.append(f"{nm}.{leaf}")
        if train_feature and len(nm.split('.')) > 1 and nm.split('.')[1] in ('feature_extractor', 'feature_projection'):
            for nm2, _k, leaves2 in tree:
                if nm2 == nm or nm2.startswith(nm + "."):
                    for leaf in leaves2:
                        p = model._params[f"{nm2}.{leaf}"]
                        p.requires_grad = True
                        params.append(p)
                        names.append(f"{nm2}.{leaf}")
    return params, names


def copy_model_and_optimizer(model, optimizer, scheduler):
    """REF/main.py:137-145."""
    model_state = deepcopy(model.state_dict())
    optimizer_state = deepcopy(optimizer.state_dict())
    if scheduler is not None:
        return model_state, optimizer_state, deepcopy(scheduler.state_dict())
    return model_state, optimizer_state, None


def load_model_and_optimizer(model, optimizer, model_state, optimizer_state, scheduler_state, scheduler=None):
    """REF/main.py:147-155.  (The reference reads `scheduler` from a module global, :151; it is an optional
    trailing argument here and falls back to the same-named global of the caller's module.)"""
    model.load_state_dict(model_state, strict=True)
    optimizer.load_state_dict(optimizer_state)
    if scheduler is None:
        scheduler = globals().get("scheduler")
    if scheduler is not None:
        scheduler.load_state_dict(scheduler_state)
        return model, optimizer, scheduler
    return model, optimizer, None


def configure_model(model):
    """REF/main.py:167-170."""
    model.requires_grad_(False)
    return model


def forward_and_adapt(x, model, optimizer, em_coef=0.9, reweight=False, temp=1., not_blank=True, scheduler=None,
                      div_coef=0, repeat_inference=True, skip_short_thd=None, pl_coef=0.0):
    """REF/main.py:172-215: forward, unsupervised loss (entropy + MCC), backward, optimizer step, forward again.
    pl_coef (keyword, extension): REF/main_SDPL.py:176's pseudo-label CTC mix (model built with pseudo_label=True)."""
    outputs = model(x).logits
    hp = optimizer.hp
    hp.em_coef, hp.reweight, hp.temp, hp.not_blank = float(em_coef), bool(reweight), float(temp), bool(not_blank)
    hp.div_coef, hp.pl_coef = float(div_coef), float(pl_coef)
    model.engine.loss_backward(hp)
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    model.zero_grad()
    if repeat_inference:
        outputs = model(x).logits
    return outputs


# --------------------------------------------------------------------------------------------------
# REF/main_SDPL.py's variants of the same functions (the SDPL baseline)
# --------------------------------------------------------------------------------------------------
def sdpl_setup_optimizer(params, opt_name='Adam', lr=1e-4, beta=0.9, weight_decay=0., scheduler=None, step_size=1, gamma=0.85):
    """REF/main_SDPL.py:17-32 (defaults: Adam, StepLR gamma 0.85)."""
    return setup_optimizer(params, opt_name, lr, beta, weight_decay, scheduler, step_size, gamma)


def sdpl_collect_params(model, bias_only=False, train_feature=False):
    """REF/main_SDPL.py:71-104 (no train_all / train_LN switches)."""
    return collect_params(model, bias_only, train_feature, False, True)


def pseudo_labeling_loss(outputs, vocab=None, processor=None):
    """REF/main_SDPL.py:194-209 (value): CTC loss of the logits [1, T, 32] against their own greedy transcript, with the
    reference's log_softmax over the time axis.  `vocab` / `processor` are accepted for signature compatibility: the
    transcript is decoded on the device with the built-in 32-symbol vocabulary."""
    if outputs.dim() != 3 or outputs.shape[0] != 1:
        raise RuntimeError("pseudo_labeling_loss expects logits of shape [1, T, 32]")
    rows = _as_rows(outputs)
    T = rows.shape[0]
    dev = rows.device
    lib = _lib.load()
    off = torch.zeros(1, dtype=torch.int64, device=dev)
    Tt = torch.full((1,), T, dtype=torch.int32, device=dev)
    ids = torch.empty(T, dtype=torch.int32, device=dev); col = torch.empty(T, dtype=torch.int32, device=dev)
    ln = torch.empty(1, dtype=torch.int32, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    check(lib.suta_op_decode(p(rows), p(off), p(Tt), 1, 32, p(ids), p(col), p(ln), _stream_ptr()))
    alpha = torch.empty(T * (2 * T + 1) + 16, dtype=torch.float32, device=dev)
    g = torch.empty(T, 32, dtype=torch.float32, device=dev); d = torch.empty(T, 32, dtype=torch.float32, device=dev)
    loss = torch.empty(4, dtype=torch.float32, device=dev); tl = torch.empty(1, dtype=torch.int32, device=dev)
    check(lib.suta_op_ctc_pseudo_label(p(rows), p(off), p(Tt), 1, p(col), p(ln), p(alpha), p(g), p(loss), p(d), p(tl), _stream_ptr()))
    return loss[3]


def sdpl_forward_and_adapt(x, model, optimizer, em_coef=0.9, reweight=False, temp=1., not_blank=True, scheduler=None, div_coef=0,
                           repeat_inference=True, pl_coef=1, vocab=None, processor=None):
    """REF/main_SDPL.py:143-191: loss * (1 - pl_coef) + pseudo_labeling_loss * pl_coef, then the same update."""
    return forward_and_adapt(x, model, optimizer, em_coef, reweight, temp, not_blank, scheduler, div_coef, repeat_inference,
                             pl_coef=pl_coef)
