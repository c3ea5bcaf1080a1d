"""CPU: the oracle restatement against the golden vectors produced by the unmodified reference (make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import suta_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TINY = ["tiny_ln", "tiny_feat", "tiny_short", "tiny_feat_noise20", "tiny_sgd", "tiny_feat_sgd", "tiny_adam_beta", "tiny_steplr",
        "tiny_bias_only", "tiny_div", "tiny_em_only", "tiny_mcc_plain", "tiny_temp1_allframes",
        "tiny_lv60_ln", "tiny_lv60_short", "tiny_lv60_feat",      # lv60: LayerNorm feature extractor + conv bias + pre-LN encoder
        "tiny_all", "tiny_lv60_all"]                              # --train_all (REF/main.py:96-100), both families
CASES = TINY + ["base_all_2s", "base_ln_5s", "base_ln_5s_noblank", "base_feat_2s", "base_feat_5s", "base_ln_30s", "large_ln_2s", "large_lv60_2s"]


def oracle_kwargs(meta):
    """adapt_utterance arguments of a fixture (fixtures of round 1 predate the optimizer / flag variants)."""
    return dict(steps=meta["steps"], train_feature=meta["train_feature"], bias_only=meta.get("bias_only", False),
                opt=meta.get("opt", "AdamW"), beta=meta.get("beta", 0.9), sched_gamma=meta.get("sched_gamma"),
                div_coef=meta.get("div_coef", 0.0), train_all=meta.get("train_all", False), **meta["hyper"])


def load(case):
    z = np.load(os.path.join(GOLD, case + ".npz"))
    return z, json.loads(bytes(z["meta"]).decode())


@pytest.mark.parametrize("case", CASES)
def test_decode_of_reference_logits_matches_reference_transcripts(case):
    z, meta = load(case)
    for k, text in meta["texts"].items():
        assert O.ctc_greedy_decode(z[f"logits_{k}"]) == text


@pytest.mark.parametrize("case", CASES)
def test_loss_and_closed_form_gradient_on_reference_logits(case):
    z, meta = load(case)
    h = meta["hyper"]
    lg = z["logits_0"]
    dc = meta.get("div_coef", 0.0)
    val, grad = O.suta_loss_grad_closed(lg, h["em_coef"], h["reweight"], h["temp"], h["not_blank"], dc)
    assert abs(val - z["losses"][0]) / abs(z["losses"][0]) < 1e-5          # loss of step 1's training forward
    t = torch.tensor(lg[None], dtype=torch.float64, requires_grad=True)
    O.suta_loss(t, h["em_coef"], h["reweight"], h["temp"], h["not_blank"], dc).backward()
    np.testing.assert_allclose(grad, t.grad[0].numpy(), rtol=1e-7, atol=1e-14)


@pytest.mark.parametrize("case", TINY)
def test_oracle_adaptation_reproduces_reference(case):
    """Every optimizer (AdamW, Adam(beta), SGD), StepLR, bias_only, div_loss, em_coef 0 / 1, temperature 1, all-frames
    entropy, 20 steps + extra noise: the oracle's loop against what the unmodified reference produced."""
    z, meta = load(case)
    cfg = getattr(O.W2V2Config, meta["cfg"])()
    sd = O.init_weights(cfg, meta["weight_seed"], blank_bias=meta["blank_bias"], ln_jitter=meta["ln_jitter"])
    x = O.normalize_audio(O.synth_audio(meta["n_samples"], meta["audio_seed"], meta.get("extra_noise", 0.0)))
    res = O.adapt_utterance(cfg, sd, x, keep_all_logits=True, **oracle_kwargs(meta))
    np.testing.assert_allclose(res.logits0, z["logits_0"], atol=5e-5)
    np.testing.assert_allclose(res.losses, z["losses"], rtol=2e-5)
    for k, text in meta["texts"].items():
        if int(k):
            np.testing.assert_allclose(res.logits[int(k)], z[f"logits_{k}"], atol=3e-4)
            assert O.ctc_greedy_decode(res.logits[int(k)]) == text
    names = O.collect_param_names(cfg, bias_only=meta.get("bias_only", False), train_feature=meta["train_feature"],
                                  train_all=meta.get("train_all", False))
    assert sorted(names) == sorted(n.lstrip(".") for n in meta["names"])    # same duplicates as REF/main.py:62-103
    for n in set(names):
        if n not in sd or n.endswith("k_proj.bias"):     # train_all: masked_spec_embed (no gradient); key biases (zero gradient: Adam on rounding noise)
            continue
        ref = z["param:" + n]
        if ref.dtype == np.float32:
            d_ref, d_got = ref - sd[n].numpy(), res.params[n] - sd[n].numpy()
            assert np.abs(d_ref - d_got).max() <= 0.05 * np.abs(d_ref).max() + 1e-9


@pytest.mark.parametrize("case", ["tiny_ln", "tiny_feat", "tiny_steplr", "tiny_adam_beta", "tiny_feat_sgd", "tiny_bias_only",
                                  "tiny_lv60_ln", "tiny_all"])
def test_hf_reference_loop_reproduces_reference(case):
    """oracle/hf_reference.py (the loop bench.py times as the reference's CPU / eager-GPU path: real HF modules,
    autograd and torch.optim under a restatement of main.py's driver) against what the unmodified reference produced."""
    from oracle.hf_reference import ReferenceLoop
    z, meta = load(case)
    cfg = getattr(O.W2V2Config, meta["cfg"])()
    sd = O.init_weights(cfg, meta["weight_seed"], blank_bias=meta["blank_bias"], ln_jitter=meta["ln_jitter"])
    x = O.normalize_audio(O.synth_audio(meta["n_samples"], meta["audio_seed"], meta.get("extra_noise", 0.0)))
    h = meta["hyper"]
    loop = ReferenceLoop(cfg, sd, "cpu", train_feature=meta["train_feature"], bias_only=meta.get("bias_only", False),
                         opt=meta.get("opt", "AdamW"), lr=h["lr"], beta=meta.get("beta", 0.9), sched_gamma=meta.get("sched_gamma"),
                         train_all=meta.get("train_all", False))
    assert loop.names == meta["names"]
    for _ in range(2):                                   # twice: the episodic restore brings everything back
        res = loop.adapt(x, steps=meta["steps"], em_coef=h["em_coef"], reweight=h["reweight"], temp=h["temp"],
                         not_blank=h["not_blank"], div_coef=meta.get("div_coef", 0.0))
        np.testing.assert_allclose(res.logits0, z["logits_0"], atol=1e-5)
        np.testing.assert_allclose(res.losses, z["losses"], rtol=1e-5)
        last = str(meta["steps"])
        np.testing.assert_allclose(res.logits[meta["steps"]], z["logits_" + last], atol=2e-5)
        for k, text in meta["texts"].items():
            assert res.texts.get(int(k), text) == text


@pytest.mark.parametrize("case", ["tiny_sdpl", "tiny_sdpl_mix", "tiny_lv60_sdpl"])
def test_oracle_sdpl_loop_reproduces_reference(case):
    """REF/main_SDPL.py through the unmodified script's functions (make_golden.make_sdpl): the oracle's pseudo-label CTC
    loss and the adapted logits / parameters after the fixture's steps."""
    z, meta = load(case)
    cfg = getattr(O.W2V2Config, meta["cfg"])()
    sd = O.init_weights(cfg, meta["weight_seed"], blank_bias=meta["blank_bias"], ln_jitter=meta["ln_jitter"], special_bias=meta["special_bias"])
    x = O.normalize_audio(O.synth_audio(meta["n_samples"], meta["audio_seed"]))
    res = O.adapt_utterance(cfg, sd, x, steps=meta["steps"], lr=meta["lr"], em_coef=meta["em_coef"], reweight=meta["reweight"],
                            temp=meta["temp"], not_blank=meta["not_blank"], train_feature=meta["train_feature"], opt=meta["opt"],
                            pl_coef=meta["pl_coef"], keep_all_logits=True)
    np.testing.assert_allclose(res.logits0, z["logits_0"], atol=5e-5)
    np.testing.assert_allclose(res.logits[meta["steps"]], z[f"logits_{meta['steps']}"], atol=3e-4)
    np.testing.assert_allclose(float(O.pseudo_labeling_loss(torch.tensor(z["logits_0"][None]))), z["pl_losses"][0], rtol=1e-5)
    skip = set(meta.get("zero_gradient_params", []))
    for n in set(meta["names"]) - skip:
        ref = z["param:" + n]
        if ref.dtype == np.float32:
            d_ref, d_got = ref - sd[n].numpy(), res.params[n] - sd[n].numpy()
            assert np.abs(d_ref - d_got).max() <= 0.05 * np.abs(d_ref).max() + 1e-9, n


def test_lv60_layernorm_only_set_matches_survey():
    """the large-lv60 shape lists 114 tensors / 108 544 elements, incl. 7 conv LayerNorm(512)."""
    cfg = O.W2V2Config.large_lv60()
    names = O.collect_param_names(cfg)
    assert len(names) == 114 and len(set(names)) == 114
    sd_shapes = {n: (cfg.conv_dim[0] if "feature_" in n else cfg.hidden_size) for n in names}
    assert sum(sd_shapes.values()) == 108544
    assert sum("feature_extractor.conv_layers" in n for n in names) == 14


def test_param_multiplicities_match_survey():
    names = O.collect_param_names(O.W2V2Config.base(), train_feature=True)
    mult = {}
    for n in names:
        mult[n] = mult.get(n, 0) + 1
    assert len(names) == 96 and len(mult) == 63
    assert mult["wav2vec2.feature_extractor.conv_layers.3.conv.weight"] == 4
    assert mult["wav2vec2.feature_extractor.conv_layers.0.layer_norm.weight"] == 4
    assert mult["wav2vec2.feature_projection.layer_norm.bias"] == 3
    assert mult["wav2vec2.feature_projection.projection.weight"] == 2
    assert mult["wav2vec2.encoder.layers.5.final_layer_norm.weight"] == 1
    assert len(O.collect_param_names(O.W2V2Config.base())) == 52


def test_adam_multiplicity_equals_sequential_substeps():
    rng = np.random.default_rng(0)
    g = rng.standard_normal(50).astype(np.float32)
    p1 = rng.standard_normal(50).astype(np.float32)
    p4 = p1.copy()
    m1, v1, m4, v4 = (np.zeros(50, np.float32) for _ in range(4))
    s = 0
    for _ in range(4):
        s = O.adam_update(p1, g, m1, v1, s, 1e-3, k=1)
    O.adam_update(p4, g, m4, v4, 0, 1e-3, k=4)
    np.testing.assert_allclose(p1, p4, rtol=0, atol=0)


def test_decode_known_answer():
    ids = [0, 11, 11, 0, 5, 15, 15, 0, 15, 8, 4, 4, 18, 8, 13, 15, 14, 0, 1, 2, 3]
    assert O.ctc_ids_to_text(O.ctc_collapse(ids)) == "HELLO WORLD<s></s><unk>"
    assert O.ctc_ids_to_text(O.ctc_collapse([0, 0, 0])) == ""
    assert O.ctc_ids_to_text(O.ctc_collapse([])) == ""


def test_wer_known_answers():
    assert O.wer(["a b c d"], ["a x c"]) == 0.5
    assert O.wer_counts(["a b", "c"], ["a b", ""]) == (1, 3)
    assert O.wer(["hello world"], ["hello world"]) == 0.0
