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


def run_engine(cfg_name, sd, wavs, steps, hp=None, keep_grads=False, train_feature=False, mult=None, sched_gamma=None,
               pseudo_label=False, train_all=False):
    """Batched SUTA on the GPU: returns per-utterance dicts (logits at checkpoints, losses, params, ids).
    sched_gamma: StepLR(step_size=1) factor, applied by the caller between steps like REF/main.py:207-208."""
    _, mcfg = _cfgs(cfg_name)
    hp = hp or AdaptHyper()
    lr0 = hp.lr
    eng = SutaEngine(mcfg, sd, train_feature=train_feature, trainable_mult=mult, pseudo_label=pseudo_label, train_all=train_all)
    eng.begin_batch(wavs)
    eng.reset()
    res = [dict(logits={}, losses=[], ids={}, pl_losses=[]) for _ in wavs]
    lg = eng.forward()
    ids = eng.decode_ids()
    for u in range(len(wavs)):
        res[u]["logits"][0] = eng.utt_logits(u).cpu().numpy().copy()
        res[u]["ids"][0] = ids[u]
    for i in range(steps):
        hp.lr = lr0 if sched_gamma is None else lr0 * sched_gamma ** i
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
            res[u]["pl_losses"].append(float(losses[3, u]))
        if (i + 1) in O.CHECKPOINT_STEPS or i + 1 == steps:
            ids = eng.decode_ids()
            for u in range(len(wavs)):
                res[u]["logits"][i + 1] = eng.utt_logits(u).cpu().numpy().copy()
                res[u]["ids"][i + 1] = ids[u]
    for u in range(len(wavs)):
        res[u]["params"] = {name: t.detach().cpu().contiguous().numpy().copy().reshape(-1) for name, t in eng.utt_params(u).items()}
    res[0]["segments"] = eng.segments
    res[0]["to_layout"] = eng.to_engine_layout
    res[0]["launches"] = eng.launch_count
    eng.close()
    return res


def oracle_forward_with(ocfg, sd, params_flat, x):
    """fp32 oracle forward of the model whose trainables are the ENGINE's adapted values: isolates the parity of the
    adaptation (what the parameters do to the logits) from the bf16 rounding noise of the engine's own forward."""
    w = dict(sd)
    for name, flat in params_flat.items():
        w[name] = torch.from_numpy(np.ascontiguousarray(flat)).view(sd[name].shape)
    with torch.no_grad():
        return O.model_forward(ocfg, w, torch.from_numpy(x)[None])[0].numpy()


def oracle_grad0(ocfg, sd, x, hp, names, div_coef=0.0):
    """d loss / d trainables at the pristine weights by fp32 autograd through the oracle forward (step 0)."""
    w = {k: v.clone() for k, v in sd.items()}
    uniq = list(dict.fromkeys(names))
    for n in uniq:
        w[n].requires_grad_(True)
    lg = O.model_forward(ocfg, w, torch.from_numpy(x)[None])
    loss = O.suta_loss(lg, hp.em_coef, hp.reweight, hp.temp, hp.not_blank, div_coef)
    gr = torch.autograd.grad(loss, [w[n] for n in uniq])
    return {n: g.numpy() for n, g in zip(uniq, gr)}


def compare(res_u, ref_logits, ref_losses, ref_params, sd, ref_ids=None, ocfg=None, x=None):
    """Error metrics of one utterance vs a reference (oracle AdaptResult-like pieces).

    What the numbers mean (tools/precision_study.py reproduces them on the CPU by rounding the GEMM operands of the
    oracle to bf16): the engine's forward carries ~0.8 % relative rounding noise, which DECORRELATES when the
    parameters move, so `dlogits_rel` (change of the engine's own logits) is noise-dominated whenever the adaptation
    moves the logits by less than that (LayerNorm-only SUTA: 2e-5 x 10 sign-like Adam steps).  The parity of the
    adaptation itself is `dlogits_via_oracle_rel`: the logit change the ENGINE's adapted parameters produce in the
    fp32 oracle forward, against the reference's logit change."""
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
    # REF/main.py:183-184,190: the entropy term averages over the frames whose argmax is not blank.  If the reference decides
    # that for some frame by a margin below the engine's logit error, the engine may select a different frame set
    # (random-init logits are nearly tied; trained models are peaky) and the gradient of that step differs by 1/n_M of
    # the entropy term: callers loosen the gradient / delta bounds in that case, and only then.
    m["mask_margin"] = float(min(np.abs(lg[:, 0] - lg[:, 1:].max(-1)).min() for lg in ref_logits.values()))
    m["param_moved_rel"] = float(np.sqrt(den / max(pv_den, 1e-30)))      # how far the reference moved: an un-adapted engine has param_delta_rel = 1
    m["logit_change_maxabs"] = float(np.abs(ref_logits[last] - ref_logits[0]).max())
    if ocfg is not None and x is not None:
        lN = oracle_forward_with(ocfg, sd, res_u["params"], x)          # every trainable of the engine (frozen ones are pristine)
        m["dlogits_via_oracle_rel"] = _rel(lN - ref_logits[0], ref_logits[last] - ref_logits[0])
    if ref_ids is not None:
        m["decode_equal_given_logits"] = all(O.ctc_collapse(np.argmax(res_u["logits"][k], -1).tolist()) == res_u["ids"][k]
                                             for k in res_u["ids"])
    return m


def check_tiny_batch(steps=10, cfg_name="tiny"):
    ocfg, _ = _cfgs(cfg_name)
    sd = O.init_weights(ocfg, 3, blank_bias=0.5, ln_jitter=0.1)
    lens, seeds = [12000, 9000, 2000, 16001], [11, 12, 13, 14]
    wavs = [O.synth_audio(n, s) for n, s in zip(lens, seeds)]
    res = run_engine(cfg_name, sd, wavs, steps, keep_grads=True)
    out = {}
    for u, w in enumerate(wavs):
        ora = O.adapt_utterance(ocfg, sd, O.normalize_audio(w), steps=steps)
        ref_logits = dict(ora.logits); ref_logits[0] = ora.logits0
        m = compare(res[u], ref_logits, ora.losses, ora.params, sd, ref_ids=True, ocfg=ocfg, x=O.normalize_audio(w))
        g0 = oracle_grad0(ocfg, sd, O.normalize_audio(w), AdaptHyper(), list(ora.params))
        m["grad0_rel"] = _grad_rel(res[u]["grad0"], res[0]["segments"], g0, res[0]["to_layout"])
        out[f"utt{u}_T{ora.logits0.shape[0]}"] = m
    return out


def check_ragged_batch(cfg_name="tiny", steps=2):
    """Edge shapes in ONE batch (empty-ish and ragged inputs, tile boundaries): the shortest possible utterance
    (T = 1), T = 2, exactly 128 and 129 frames (one GEMM / attention tile and one row more), 256 frames, a long one, and
    40 more short ones so that the batch exceeds one 64-utterance wave of per-utterance tables.  Forward parity for every
    utterance; `steps` adaptation steps must reproduce the oracle's logits for the utterances whose loss is defined."""
    ocfg, mcfg = _cfgs(cfg_name)
    sd = O.init_weights(ocfg, 3, blank_bias=0.5, ln_jitter=0.1)

    def n_for(T):
        n = 400
        while ocfg.frames(n) < T:
            n += 1
        return n
    lens = [n_for(1), n_for(2), n_for(128), n_for(129), n_for(256) + 3, 130001, 3001] + [2000 + 37 * i for i in range(40)]
    wavs = [O.synth_audio(n, 100 + i) for i, n in enumerate(lens)]
    res = run_engine(cfg_name, sd, wavs, steps)
    out = {"frames": [int(r["logits"][0].shape[0]) for r in res[:7]]}
    worst0 = worstN = 0.0
    for u in list(range(7)) + [10, 46]:
        x = O.normalize_audio(wavs[u])
        with torch.no_grad():
            ref0 = O.model_forward(ocfg, sd, torch.from_numpy(x)[None])[0].numpy()
        worst0 = max(worst0, float(np.abs(res[u]["logits"][0] - ref0).max()))
        if u >= 2:                                        # T = 1 / 2: the loss of a one-frame utterance is degenerate
            ora = O.adapt_utterance(ocfg, sd, x, steps=steps, keep_all_logits=True)
            if np.isfinite(ora.losses).all():
                worstN = max(worstN, float(np.abs(res[u]["logits"][steps] - ora.logits[steps]).max()))
                out[f"loss_rel_u{u}"] = float(np.max(np.abs(np.asarray(res[u]["losses"]) - np.asarray(ora.losses)) / np.abs(ora.losses)))
    out["logits0_maxabs"], out["logitsN_maxabs"] = worst0, worstN
    out["all_finite"] = bool(all(np.isfinite(r["logits"][steps]).all() for r in res[2:]))
    return out


def _grad_rel(flat, segments, g0, to_layout):
    num = den = 0.0
    for name, off, size in segments:
        if name in g0:
            ref = to_layout(name, torch.from_numpy(g0[name])).numpy().astype(np.float64)
            num += float(((flat[off:off + size] - ref) ** 2).sum()); den += float((ref ** 2).sum())
    return float(np.sqrt(num / max(den, 1e-300)))


def _mult(ocfg, train_feature, bias_only=False, train_all=False):
    m = {}
    for n in O.collect_param_names(ocfg, bias_only=bias_only, train_feature=train_feature, train_all=train_all):
        m[n] = m.get(n, 0) + 1
    return m


def check_tiny_feat_batch(steps=5, cfg_name="tiny"):
    """train_feature: per-utterance CNN + projection weights, duplicate-parameter Adam semantics (REF/main.py:88-94)."""
    ocfg, _ = _cfgs(cfg_name)
    sd = O.init_weights(ocfg, 4, blank_bias=0.35, ln_jitter=0.1)
    lens, seeds = [9000, 12000, 3000], [12, 15, 16]
    wavs = [O.synth_audio(n, s) for n, s in zip(lens, seeds)]
    res = run_engine(cfg_name, sd, wavs, steps, keep_grads=True, train_feature=True, mult=_mult(ocfg, True))
    out = {}
    for u, w in enumerate(wavs):
        ora = O.adapt_utterance(ocfg, sd, O.normalize_audio(w), steps=steps, train_feature=True)
        ref_logits = dict(ora.logits); ref_logits[0] = ora.logits0
        m = compare(res[u], ref_logits, ora.losses, ora.params, sd, ref_ids=True, ocfg=ocfg, x=O.normalize_audio(w))
        # per-group delta errors localise a broken gradient
        for grp in ("conv_layers.0.conv.weight", "conv_layers.0.conv.bias", "conv_layers.0.layer_norm", "conv_layers.3.conv.weight",
                    "conv_layers.3.conv.bias", "conv_layers.3.layer_norm", "conv_layers.6.conv.weight", "conv_layers.6.layer_norm",
                    "feature_projection.projection.weight", "feature_projection.projection.bias", "feature_projection.layer_norm",
                    "encoder.layers.0.layer_norm"):
            num = den = 0.0
            for name, p in ora.params.items():
                if grp in name:
                    p0 = sd[name].numpy().reshape(-1)
                    num += float(np.sum((res[u]["params"][name] - p.reshape(-1)) ** 2))
                    den += float(np.sum((p.reshape(-1) - p0) ** 2))
            if den > 0:
                m["delta:" + grp] = float(np.sqrt(num / max(den, 1e-30)))
        out[f"utt{u}_T{ora.logits0.shape[0]}"] = m
    return out


def check_tiny_stages(cfg_name="tiny"):
    """Intermediate activations of the forward vs the oracle's taps (localises a broken stage)."""
    ocfg, mcfg = _cfgs(cfg_name)
    stable = ocfg.do_stable_layer_norm
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
        for l in (0, 3, 6):
            out[f"u{u}_conv{l}"] = _rel(_conv_rows(eng, l, u).float().cpu().numpy(), taps[f"conv{l}"][0].t().numpy())
        if stable:      # pre-LN: h0 + pos-conv enters layer 0 directly (buffer h1_0); "hE" holds the stream after the last layer
            out[f"u{u}_h1_0"] = _rel(eng.debug_buffer("h1_0")[o:o + T].cpu().numpy(), taps["pos"][0].numpy())
            out[f"u{u}_h2_0"] = _rel(eng.debug_buffer("h2_0")[o:o + T].cpu().numpy(), taps["pre_ln2_0"][0].numpy())
            out[f"u{u}_hE"] = _rel(eng.debug_buffer("hE")[o:o + T].cpu().numpy(), taps[f"layer{ocfg.num_hidden_layers - 1}"][0].numpy())
        else:
            out[f"u{u}_hE"] = _rel(eng.debug_buffer("hE")[o:o + T].cpu().numpy(), taps["pos"][0].numpy())
            out[f"u{u}_h2_0"] = _rel(eng.debug_buffer("h2_0")[o:o + T].cpu().numpy(), taps["pre_ln2_0"][0].numpy())
            out[f"u{u}_x_final"] = _rel(eng.debug_buffer("x_final")[o:o + T].cpu().numpy(), taps[f"layer{ocfg.num_hidden_layers - 1}"][0].numpy())
        out[f"u{u}_logits"] = _rel(eng.utt_logits(u).cpu().numpy(), lg[0].numpy())
    eng.close()
    return out


def _conv_rows(eng, l, u):
    """Valid rows of utterance u in conv layer l's output (the engine's row offsets follow csrc/engine.cu::plan_batch)."""
    c = eng.cfg
    Ls = []
    for n in eng.lengths:
        L, per = int(n), []
        for k, s in zip(c.conv_kernel, c.conv_stride):
            L = (L - k) // s + 1
            per.append(L)
        Ls.append(per)
    last = l == len(c.conv_dim) - 1
    cnn_bwd = eng.train_feature or c.feat_extract_norm == "layer"
    off = 0
    for v in range(u):
        off += Ls[v][l] if last else (((Ls[v][l] + 1 + 255) & ~255) if cnn_bwd else ((Ls[v][l] + 7) & ~7))
    return eng.debug_buffer(f"conv{l}")[off:off + Ls[u][l]]


def check_golden_sdpl(case):
    """The SDPL baseline (REF/main_SDPL.py): fixtures produced by the unmodified script's functions."""
    z, meta = load_golden(case)
    ocfg, _ = _cfgs(meta["cfg"])
    sd = O.init_weights(ocfg, meta["weight_seed"], blank_bias=meta["blank_bias"], ln_jitter=meta["ln_jitter"],
                        special_bias=meta["special_bias"])
    wav = O.synth_audio(meta["n_samples"], meta["audio_seed"])
    hp = AdaptHyper(lr=meta["lr"], em_coef=meta["em_coef"], reweight=meta["reweight"], temp=meta["temp"], not_blank=meta["not_blank"],
                    opt=meta["opt"], pl_coef=meta["pl_coef"])
    tf = bool(meta["train_feature"])
    res = run_engine(meta["cfg"], sd, [wav], meta["steps"], hp, keep_grads=True, train_feature=tf, mult=_mult(ocfg, tf),
                     pseudo_label=True)[0]
    skip = set(meta.get("zero_gradient_params", []))         # analytically zero gradient: Adam steps on rounding noise
    ref_logits = {int(k.split("_")[1]): z[k] for k in z.files if k.startswith("logits_")}
    ref_params = {k[6:]: z[k] for k in z.files if k.startswith("param:") and z[k].dtype == np.float32 and k[6:] not in skip}
    x = O.normalize_audio(wav)
    m = compare(res, ref_logits, [abs(v) + 1.0 for v in res["losses"]], ref_params, sd, ref_ids=True)   # total loss is not in the fixture
    del m["loss_rel_max"]
    m["pl_loss_rel_max"] = float(np.max(np.abs(np.asarray(res["pl_losses"]) - z["pl_losses"]) / np.abs(z["pl_losses"])))
    # gradient at step 0 vs fp32 autograd through the oracle (the CTC term included)
    w = {k: v.clone() for k, v in sd.items()}
    uniq = [n for n in dict.fromkeys(meta["names"]) if n not in skip]
    for n in uniq:
        w[n].requires_grad_(True)
    lg = O.model_forward(ocfg, w, torch.from_numpy(x)[None])
    loss = O.suta_loss(lg, hp.em_coef, hp.reweight, hp.temp, hp.not_blank, 0.0, hp.pl_coef)
    g0 = {n: g.numpy() for n, g in zip(uniq, torch.autograd.grad(loss, [w[n] for n in uniq]))}
    m["grad0_rel"] = _grad_rel(res["grad0"], res["segments"], g0, res["to_layout"])
    return m


def check_determinism(cfg_name="base", train_feature=False, steps=3):
    """The same batch adapted twice: are gradients, parameters and logits the same BITS?"""
    ocfg, _ = _cfgs(cfg_name)
    sd = O.init_weights(ocfg, 0, blank_bias=1.75)
    wavs = [O.synth_audio(n, s) for n, s in ((80000, 1), (33000, 2), (120000, 3), (48000, 4))]
    runs = [run_engine(cfg_name, sd, wavs, steps, keep_grads=True, train_feature=train_feature, mult=_mult(ocfg, train_feature))
            for _ in range(2)]
    out = dict(grad0=True, params=True, logits=True)
    for a, b in zip(*runs):
        out["grad0"] &= bool(np.array_equal(a["grad0"], b["grad0"]))
        out["params"] &= all(np.array_equal(a["params"][n], b["params"][n]) for n in a["params"])
        out["logits"] &= all(np.array_equal(a["logits"][k], b["logits"][k]) for k in a["logits"])
    return out


def check_golden(case, with_grad0=True):
    z, meta = load_golden(case)
    ocfg, _ = _cfgs(meta["cfg"])
    sd = O.init_weights(ocfg, meta["weight_seed"], blank_bias=meta["blank_bias"], ln_jitter=meta["ln_jitter"])
    wav = O.synth_audio(meta["n_samples"], meta["audio_seed"], meta.get("extra_noise", 0.0))
    hp = AdaptHyper(**{k: meta["hyper"][k] for k in ("lr", "em_coef", "reweight", "temp", "not_blank")},
                    opt=meta.get("opt", "AdamW"), beta1=meta.get("beta", 0.9) if meta.get("opt") == "Adam" else 0.9,
                    div_coef=meta.get("div_coef", 0.0))
    tf, bo, ta = bool(meta["train_feature"]), bool(meta.get("bias_only", False)), bool(meta.get("train_all", False))
    res = run_engine(meta["cfg"], sd, [wav], meta["steps"], hp, keep_grads=True, train_feature=tf, mult=_mult(ocfg, tf, bo, ta),
                     sched_gamma=meta.get("sched_gamma"), train_all=ta)[0]
    hp.lr = meta["hyper"]["lr"]
    ref_logits = {int(k.split("_")[1]): z[k] for k in z.files if k.startswith("logits_")}
    # (train_all: the gradient of a key bias is identically zero -- it shifts every score of a softmax row alike -- so Adam
    #  random-walks on rounding noise there, in the reference as well: tests/golden/make_golden.py)
    ref_params = {k[6:]: z[k] for k in z.files if k.startswith("param:") and z[k].dtype == np.float32 and not k.endswith("k_proj.bias")}
    x = O.normalize_audio(wav)
    m = compare(res, ref_logits, z["losses"], ref_params, sd, ref_ids=True, ocfg=ocfg, x=x)
    # big tensors are stored as (sum, sum|.|, sum .^2, first 4096 values): compare the head and the moments of the DELTA
    big = {k[6:]: z[k] for k in z.files if k.startswith("param:") and z[k].dtype == np.float64}
    num = den = 0.0
    for name, packed in big.items():
        p0 = sd[name].numpy().reshape(-1)[:4096].astype(np.float64)
        d_ref, d_got = packed[3:] - p0, res["params"][name][:4096].astype(np.float64) - p0
        num += float(((d_ref - d_got) ** 2).sum()); den += float((d_ref ** 2).sum())
    if big:
        m["big_param_head_delta_rel"] = float(np.sqrt(num / max(den, 1e-300)))
    if with_grad0:
        names = [n.lstrip(".") for n in meta["names"]]          # (train_all lists the root module's parameters as ".<name>")
        g0 = oracle_grad0(ocfg, sd, x, hp, [n for n in names if n in sd and not n.endswith("k_proj.bias")], meta.get("div_coef", 0.0))
        m["grad0_rel"] = _grad_rel(res["grad0"], res["segments"], g0, res["to_layout"])
        if ta:              # per-tensor gradient errors localise a broken weight-gradient path
            m["grad0_by_tensor"] = {n: _grad_rel(res["grad0"], [sg for sg in res["segments"] if sg[0] == n], g0, res["to_layout"])
                                    for n in g0}
    m["texts_equal"] = {k: (O.ctc_ids_to_text(res["ids"][int(k)]) == v) for k, v in meta["texts"].items() if int(k) in res["ids"]}
    a0 = np.argmax(res["logits"][0], -1); r0 = np.argmax(ref_logits[0], -1)
    m["argmax_agree0"] = float((a0 == r0).mean())
    # smallest gap between the best and the second-best logit over the reference's frames: a frame whose gap is below the
    # engine's logit error can legitimately decode differently (random-init logits are nearly tied)
    srt = np.sort(ref_logits[0], -1)
    m["ref_min_top2_gap"] = float((srt[:, -1] - srt[:, -2]).min())
    m["launches"] = res.get("launches")
    return m


ALL = [("determinism_base_ln", check_determinism), ("determinism_tiny_feat", lambda: check_determinism("tiny", True)),
       ("tiny_stages", check_tiny_stages), ("tiny_batch", check_tiny_batch), ("tiny_feat_batch", check_tiny_feat_batch),
       ("golden_tiny_feat", lambda: check_golden("tiny_feat")), ("golden_base_feat_2s", lambda: check_golden("base_feat_2s")),
       ("golden_tiny_ln", lambda: check_golden("tiny_ln")), ("golden_tiny_short", lambda: check_golden("tiny_short")),
       ("golden_base_ln_5s", lambda: check_golden("base_ln_5s")),
       ("golden_base_ln_5s_noblank", lambda: check_golden("base_ln_5s_noblank"))] + [
       ("golden_tiny_sdpl", lambda: check_golden_sdpl("tiny_sdpl")), ("golden_tiny_sdpl_mix", lambda: check_golden_sdpl("tiny_sdpl_mix"))] + [
       ("golden_" + c, (lambda c=c: check_golden(c))) for c in
       ("tiny_sgd", "tiny_feat_sgd", "tiny_adam_beta", "tiny_steplr", "tiny_bias_only", "tiny_div", "tiny_em_only",
        "tiny_mcc_plain", "tiny_temp1_allframes", "tiny_feat_noise20", "large_ln_2s", "base_feat_5s", "base_ln_30s")]
