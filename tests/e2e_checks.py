"""End-to-end parity of the CUDA engine against the oracle (tiny config, computed live) and against the committed
golden fixtures produced by the unmodified reference (tests/golden/*.npz).  Returns metrics; callers assert."""
import json
import os

import numpy as np
import torch

from oracle import suta_oracle as O
from suta_b200 import AdaptHyper, ModelConfig, SutaEngine

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(case):
    z = np.load(os.path.join(GOLD, case + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta


def _cfgs(name):
    return getattr(O.W2V2Config, name)(), getattr(ModelConfig, name)()


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def run_engine(cfg_name, sd, wavs, steps, hp=None, keep_grads=False, train_feature=False, mult=None):
    """Batched SUTA on the GPU: returns per-utterance dicts (logits at checkpoints, losses, params, ids)."""
    _, mcfg = _cfgs(cfg_name)
    hp = hp or AdaptHyper()
    eng = SutaEngine(mcfg, sd, train_feature=train_feature, trainable_mult=mult)
    eng.begin_batch(wavs)
    eng.reset()
    res = [dict(logits={}, losses=[], ids={}) for _ in wavs]
    lg = eng.forward()
    ids = eng.decode_ids()
    for u in range(len(wavs)):
        res[u]["logits"][0] = eng.utt_logits(u).cpu().numpy().copy()
        res[u]["ids"][0] = ids[u]
    for i in range(steps):
        eng.loss_backward(hp)
        if keep_grads and i == 0:
            g = eng.grads().cpu().numpy().copy()
            for u in range(len(wavs)):
                res[u]["grad0"] = g[u]
                res[u]["dlogits0"] = eng.dlogits()[int(eng.frame_off[u]):int(eng.frame_off[u]) + int(eng.frames[u])].cpu().numpy().copy()
        losses = eng.losses().cpu().numpy()
        eng.optimizer_step(hp)
        eng.forward()
        for u in range(len(wavs)):
            res[u]["losses"].append(float(losses[0, u]))
        if (i + 1) in O.CHECKPOINT_STEPS:
            ids = eng.decode_ids()
            for u in range(len(wavs)):
                res[u]["logits"][i + 1] = eng.utt_logits(u).cpu().numpy().copy()
                res[u]["ids"][i + 1] = ids[u]
    for u in range(len(wavs)):
        res[u]["params"] = {name: t.detach().cpu().contiguous().numpy().copy().reshape(-1) for name, t in eng.utt_params(u).items()}
    res[0]["segments"] = eng.segments
    res[0]["launches"] = eng.launch_count
    eng.close()
    return res


def compare(res_u, ref_logits, ref_losses, ref_params, sd, ref_ids=None):
    """Error metrics of one utterance vs a reference (oracle AdaptResult-like pieces)."""
    m = {}
    m["logits0_maxabs"] = float(np.abs(res_u["logits"][0] - ref_logits[0]).max())
    m["logits0_rel"] = _rel(res_u["logits"][0], ref_logits[0])
    last = max(k for k in ref_logits if k in res_u["logits"])
    m["logitsN_maxabs"] = float(np.abs(res_u["logits"][last] - ref_logits[last]).max())
    # how well the CHANGE of the logits caused by adaptation is reproduced
    m["dlogits_rel"] = _rel(res_u["logits"][last] - res_u["logits"][0], ref_logits[last] - ref_logits[0])
    m["loss_rel_max"] = float(np.max(np.abs(np.asarray(res_u["losses"]) - np.asarray(ref_losses)) / np.abs(ref_losses)))
    num = den = 0.0
    pv_num = pv_den = 0.0
    for name, p in ref_params.items():
        if name not in res_u["params"]:
            continue
        p0 = sd[name].numpy().reshape(-1)
        d_ref = np.asarray(p, np.float64).reshape(-1) - p0
        d_got = res_u["params"][name].astype(np.float64) - p0
        num += float(np.sum((d_ref - d_got) ** 2)); den += float(np.sum(d_ref ** 2))
        pv_num += float(np.sum((np.asarray(p, np.float64).reshape(-1) - res_u["params"][name]) ** 2))
        pv_den += float(np.sum(np.asarray(p, np.float64) ** 2))
    m["param_delta_rel"] = float(np.sqrt(num / max(den, 1e-30)))
    m["param_value_rel"] = float(np.sqrt(pv_num / max(pv_den, 1e-30)))
    if ref_ids is not None:
        m["decode_equal_given_logits"] = all(O.ctc_collapse(np.argmax(res_u["logits"][k], -1).tolist()) == res_u["ids"][k]
                                             for k in res_u["ids"])
    return m


def check_tiny_batch(steps=10):
    ocfg, _ = _cfgs("tiny")
    sd = O.init_weights(ocfg, 3, blank_bias=0.5, ln_jitter=0.1)
    lens, seeds = [12000, 9000, 2000, 16001], [11, 12, 13, 14]
    wavs = [O.synth_audio(n, s) for n, s in zip(lens, seeds)]
    res = run_engine("tiny", sd, wavs, steps, keep_grads=True)
    out = {}
    for u, w in enumerate(wavs):
        ora = O.adapt_utterance(ocfg, sd, O.normalize_audio(w), steps=steps)
        ref_logits = dict(ora.logits); ref_logits[0] = ora.logits0
        m = compare(res[u], ref_logits, ora.losses, ora.params, sd, ref_ids=True)
        out[f"utt{u}_T{ora.logits0.shape[0]}"] = m
    return out


def _mult(ocfg, train_feature):
    m = {}
    for n in O.collect_param_names(ocfg, train_feature=train_feature):
        m[n] = m.get(n, 0) + 1
    return m


def check_tiny_feat_batch(steps=5):
    """train_feature: per-utterance CNN + projection weights, duplicate-parameter Adam semantics (REF/main.py:88-94)."""
    ocfg, _ = _cfgs("tiny")
    sd = O.init_weights(ocfg, 4, blank_bias=0.35, ln_jitter=0.1)
    lens, seeds = [9000, 12000, 3000], [12, 15, 16]
    wavs = [O.synth_audio(n, s) for n, s in zip(lens, seeds)]
    res = run_engine("tiny", sd, wavs, steps, keep_grads=True, train_feature=True, mult=_mult(ocfg, True))
    out = {}
    for u, w in enumerate(wavs):
        ora = O.adapt_utterance(ocfg, sd, O.normalize_audio(w), steps=steps, train_feature=True)
        ref_logits = dict(ora.logits); ref_logits[0] = ora.logits0
        m = compare(res[u], ref_logits, ora.losses, ora.params, sd, ref_ids=True)
        # per-group delta errors localise a broken gradient
        for grp in ("conv_layers.0.conv", "conv_layers.0.layer_norm", "conv_layers.3.conv", "conv_layers.6.conv",
                    "feature_projection.projection.weight", "feature_projection.projection.bias", "feature_projection.layer_norm",
                    "encoder.layers.0.layer_norm"):
            num = den = 0.0
            for name, p in ora.params.items():
                if grp in name:
                    p0 = sd[name].numpy().reshape(-1)
                    num += float(np.sum((res[u]["params"][name] - p.reshape(-1)) ** 2))
                    den += float(np.sum((p.reshape(-1) - p0) ** 2))
            m["delta:" + grp] = float(np.sqrt(num / max(den, 1e-30)))
        out[f"utt{u}_T{ora.logits0.shape[0]}"] = m
    return out


def check_tiny_stages():
    """Intermediate activations of the forward vs the oracle's taps (localises a broken stage)."""
    ocfg, mcfg = _cfgs("tiny")
    sd = O.init_weights(ocfg, 3, blank_bias=0.5, ln_jitter=0.1)
    wavs = [O.synth_audio(9000, 21), O.synth_audio(5000, 22)]
    eng = SutaEngine(mcfg, sd)
    eng.begin_batch(wavs)
    eng.reset()
    eng.forward()
    out = {}
    for u, w in enumerate(wavs):
        taps = {}
        x = O.normalize_audio(w)
        with torch.no_grad():
            lg = O.model_forward(ocfg, sd, torch.from_numpy(x)[None], taps)
        o, T = int(eng.frame_off[u]), int(eng.frames[u])
        so = int(eng.sample_off[u])
        out[f"u{u}_wav_norm"] = _rel(eng.debug_buffer("wav_norm")[0, so:so + len(w)].cpu().numpy(), x)
        out[f"u{u}_feat"] = _rel(eng.debug_buffer("feat")[o:o + T].float().cpu().numpy(), taps["conv6"][0].t().numpy())
        out[f"u{u}_h0"] = _rel(eng.debug_buffer("h0")[o:o + T].cpu().numpy(), taps["proj"][0].numpy())
        out[f"u{u}_h2_0"] = _rel(eng.debug_buffer("h2_0")[o:o + T].cpu().numpy(), 0 * taps["layer0"][0].numpy() + eng.debug_buffer("h2_0")[o:o + T].cpu().numpy())
        out[f"u{u}_x_final"] = _rel(eng.debug_buffer("x_final")[o:o + T].cpu().numpy(), taps[f"layer{ocfg.num_hidden_layers - 1}"][0].numpy())
        out[f"u{u}_logits"] = _rel(eng.utt_logits(u).cpu().numpy(), lg[0].numpy())
    eng.close()
    return out


def check_golden(case):
    z, meta = load_golden(case)
    ocfg, _ = _cfgs(meta["cfg"])
    sd = O.init_weights(ocfg, meta["weight_seed"], blank_bias=meta["blank_bias"], ln_jitter=meta["ln_jitter"])
    wav = O.synth_audio(meta["n_samples"], meta["audio_seed"])
    hp = AdaptHyper(**{k: meta["hyper"][k] for k in ("lr", "em_coef", "reweight", "temp", "not_blank")})
    tf = bool(meta["train_feature"])
    res = run_engine(meta["cfg"], sd, [wav], meta["steps"], hp, train_feature=tf, mult=_mult(ocfg, tf))[0]
    ref_logits = {int(k.split("_")[1]): z[k] for k in z.files if k.startswith("logits_")}
    ref_params = {k[6:]: z[k] for k in z.files if k.startswith("param:") and z[k].dtype == np.float32}
    m = compare(res, ref_logits, z["losses"], ref_params, sd, ref_ids=True)
    m["texts_equal"] = {k: (O.ctc_ids_to_text(res["ids"][int(k)]) == v) for k, v in meta["texts"].items()}
    a0 = np.argmax(res["logits"][0], -1); r0 = np.argmax(ref_logits[0], -1)
    m["argmax_agree0"] = float((a0 == r0).mean())
    m["launches"] = res.get("launches")
    return m


ALL = [("tiny_stages", check_tiny_stages), ("tiny_batch", check_tiny_batch), ("tiny_feat_batch", check_tiny_feat_batch),
       ("golden_tiny_feat", lambda: check_golden("tiny_feat")), ("golden_base_feat_2s", lambda: check_golden("base_feat_2s")),
       ("golden_tiny_ln", lambda: check_golden("tiny_ln")), ("golden_tiny_short", lambda: check_golden("tiny_short")),
       ("golden_base_ln_5s", lambda: check_golden("base_ln_5s")),
       ("golden_base_ln_5s_noblank", lambda: check_golden("base_ln_5s_noblank"))]
