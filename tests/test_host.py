"""CPU: host-side logic of the package (no CUDA calls)."""
import types

import numpy as np
import pytest

from oracle import suta_oracle as O
from suta_b200 import ModelConfig
from suta_b200.data import librispeech_shaped
from suta_b200.shard import bucket_batches, shard_lpt, utterance_cost
from suta_b200.text import CTCVocab
from suta_b200.wer import wer, wer_counts


def test_frames_formula_matches_oracle_and_survey():
    for cfg, o in ((ModelConfig.base(), O.W2V2Config.base()), (ModelConfig.tiny(), O.W2V2Config.tiny())):
        for n in (400, 2000, 80000, 104000, 480000, 560000, 123457):
            assert cfg.frames(n) == o.frames(n)
    b = ModelConfig.base()
    assert (b.frames(80000), b.frames(104000), b.frames(480000), b.frames(560000)) == (249, 324, 1499, 1749)
    assert b.frames(b.min_samples) == 1 and b.frames(b.min_samples - 1) < 1


def test_text_and_wer_agree_with_oracle():
    v = CTCVocab()
    rng = np.random.default_rng(0)
    for _ in range(50):
        ids = rng.integers(0, 32, rng.integers(0, 60)).tolist()
        col = O.ctc_collapse(ids)
        assert v.ids_to_text(col) == O.ctc_ids_to_text(col)
    words = ["A", "B", "C", "D", "E"]
    for _ in range(50):
        r = " ".join(rng.choice(words, rng.integers(1, 12)))
        h = " ".join(rng.choice(words, rng.integers(0, 12)))
        assert wer_counts([r], [h]) == O.wer_counts([r], [h])
    assert wer("a b c d", "a x c") == 0.5
    with pytest.raises(ValueError):
        wer_counts(["a"], [])


def test_vocab_matches_reference_table():
    assert CTCVocab().id_to_tok == O.VOCAB


def test_shard_lpt_is_a_balanced_partition():
    utts = librispeech_shaped(500)
    cfg = ModelConfig.base()
    costs = [utterance_cost(cfg.frames(u.n_samples)) for u in utts]
    for w in (1, 2, 4, 8):
        shards = shard_lpt(costs, w)
        assert sorted(i for s in shards for i in s) == list(range(500))
        loads = [sum(costs[i] for i in s) for s in shards]
        assert max(loads) / (sum(loads) / w) < 1.05


def test_bucket_batches_respect_limits_and_cover():
    utts = librispeech_shaped(300)
    cfg = ModelConfig.base()
    frames = [cfg.frames(u.n_samples) for u in utts]
    batches = bucket_batches(frames, list(range(300)), 64, 20000)
    assert sorted(i for b in batches for i in b) == list(range(300))
    for b in batches:
        assert len(b) <= 64 and (len(b) == 1 or sum(frames[i] for i in b) <= 20000)
        assert [frames[i] for i in b] == sorted((frames[i] for i in b), reverse=True)


def test_synthetic_set_shape():
    utts = librispeech_shaped(2939)
    d = np.array([u.duration for u in utts])
    assert len(utts) == 2939 and d.min() >= 2.0 and d.max() <= 35.0
    assert 5.0 < d.sum() / 3600 < 5.6          # LibriSpeech test-other is ~5.3 h
    a, b = utts[3].audio(), utts[3].audio()
    assert a.dtype == np.float32 and len(a) == utts[3].n_samples and np.array_equal(a, b)


def test_collect_params_names_and_duplicates_match_reference_walk(capsys):
    from suta_b200 import api
    for cfg, ocfg in ((ModelConfig.base(), O.W2V2Config.base()), (ModelConfig.tiny(), O.W2V2Config.tiny()),
                      (ModelConfig.large_lv60(), O.W2V2Config.large_lv60())):
        names_all = {f"{nm}.{leaf}" for nm, _k, leaves in api._module_order(cfg) for leaf in leaves}
        model = types.SimpleNamespace(cfg=cfg, engine=types.SimpleNamespace(train_feature=True),
                                      _params={n: types.SimpleNamespace(name=n, requires_grad=False) for n in names_all})
        for tf in (False, True):
            for bias_only in (False, True):
                params, names = api.collect_params(model, bias_only, tf, False, True)
                assert sorted(names) == sorted(O.collect_param_names(ocfg, bias_only=bias_only, train_feature=tf))
                assert [p.name for p in params] == names
    capsys.readouterr()


def test_train_all_walk_matches_the_hf_module_tree_and_the_reference_walk(capsys):
    """REF/main.py:96-100 lists every parameter once per enclosing module.  The host-side module tree is checked against
    the real HF module tree (names, order, own parameters), the listed names and their order against the reference's
    walk over the real model (oracle/hf_reference.collect_params restates REF/main.py:62-103 over nn.Modules)."""
    import contextlib, io, warnings
    from transformers import Wav2Vec2ForCTC
    from oracle import hf_reference as HR
    from suta_b200 import api
    for cfg, ocfg in ((ModelConfig.tiny(), O.W2V2Config.tiny()), (ModelConfig.tiny_lv60(), O.W2V2Config.tiny_lv60())):
        hf = Wav2Vec2ForCTC(ocfg.to_hf()).eval()
        assert [(nm, leaves) for nm, _k, leaves in api._module_tree_all(cfg)] == \
            [(nm, [n for n, _ in m.named_parameters(recurse=False)]) for nm, m in hf.named_modules()]
        names_all = {(f"{nm}.{leaf}" if nm else leaf) for nm, _k, leaves in api._module_tree_all(cfg) for leaf in leaves}
        model = types.SimpleNamespace(cfg=cfg, engine=types.SimpleNamespace(train_feature=True, train_all=True),
                                      _params={n: types.SimpleNamespace(name=n, requires_grad=False) for n in names_all})
        for bias_only, tf, train_LN in ((False, False, True), (False, True, True), (True, False, True), (False, False, False)):
            with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                _, ref_names = HR.collect_params(hf, bias_only, tf, True, train_LN)
            params, names = api.collect_params(model, bias_only, tf, True, train_LN)
            assert names == ref_names
            mult = api.reference_multiplicities(cfg, bias_only, tf, train_LN, train_all=True)
            want = {}
            for p in params:
                want[p.name] = want.get(p.name, 0) + 1
            assert mult == want
            assert sorted(n for n in (x.lstrip(".") for x in ref_names)) == \
                sorted(O.collect_param_names(ocfg, bias_only=bias_only, train_feature=tf, train_LN=train_LN, train_all=True))
        assert mult["wav2vec2.encoder.layers.0.attention.q_proj.weight"] == 7 and mult["lm_head.weight"] == 2
        assert mult["wav2vec2.encoder.pos_conv_embed.conv.parametrizations.weight.original1"] == 7
    assert mult["wav2vec2.feature_extractor.conv_layers.3.layer_norm.weight"] == 6       # lv60: 6 enclosing modules, train_LN off in the last pass
    capsys.readouterr()


def test_load_checkpoint_reads_a_local_hf_directory(tmp_path):
    """`--asr <local dir>` -- a Wav2Vec2ForCTC checkpoint written by save_pretrained (safetensors, new
    weight-norm parametrisation) and one in the legacy layout (pytorch_model.bin, weight_g / weight_v) both load into
    the names the engine packs; so does a checkpoint of the lv60 family (stable LayerNorm / conv LayerNorm / conv bias)."""
    import pytest
    import torch
    from transformers import Wav2Vec2Config, Wav2Vec2ForCTC
    from oracle import suta_oracle as O
    from suta_b200.config import ModelConfig
    from suta_b200.weights import load_checkpoint
    ocfg = O.W2V2Config.tiny()
    model = Wav2Vec2ForCTC(ocfg.to_hf()).eval()
    model.load_state_dict(O.init_weights(ocfg, 3, blank_bias=0.5, ln_jitter=0.1), strict=False)
    d1 = tmp_path / "new"
    model.save_pretrained(str(d1))
    cfg, sd = load_checkpoint(str(d1))
    assert cfg == ModelConfig.tiny()
    ref = model.state_dict()
    for k in ("wav2vec2.feature_extractor.conv_layers.3.conv.weight", "wav2vec2.encoder.layers.1.final_layer_norm.bias",
              "wav2vec2.encoder.pos_conv_embed.conv.parametrizations.weight.original1", "lm_head.bias"):
        assert torch.equal(sd[k], ref[k]), k
    # legacy layout
    d2 = tmp_path / "old"
    d2.mkdir()
    model.config.save_pretrained(str(d2))
    old = {k: v for k, v in ref.items()}
    old["wav2vec2.encoder.pos_conv_embed.conv.weight_g"] = old.pop("wav2vec2.encoder.pos_conv_embed.conv.parametrizations.weight.original0")
    old["wav2vec2.encoder.pos_conv_embed.conv.weight_v"] = old.pop("wav2vec2.encoder.pos_conv_embed.conv.parametrizations.weight.original1")
    torch.save(old, str(d2 / "pytorch_model.bin"))
    cfg2, sd2 = load_checkpoint(str(d2))
    assert cfg2 == cfg and torch.equal(sd2["wav2vec2.encoder.pos_conv_embed.conv.parametrizations.weight.original0"],
                                       ref["wav2vec2.encoder.pos_conv_embed.conv.parametrizations.weight.original0"])
    # the lv60 family (REF/main_SDPL.py:238-241): conv LayerNorms and conv biases come along
    lcfg = O.W2V2Config.tiny_lv60()
    m3 = Wav2Vec2ForCTC(lcfg.to_hf()).eval()
    m3.load_state_dict(O.init_weights(lcfg, 3, blank_bias=0.5, ln_jitter=0.1), strict=False)
    d3 = tmp_path / "lv60"
    m3.save_pretrained(str(d3))
    cfg3, sd3 = load_checkpoint(str(d3))
    assert cfg3 == ModelConfig.tiny_lv60()
    for k in ("wav2vec2.feature_extractor.conv_layers.5.layer_norm.weight", "wav2vec2.feature_extractor.conv_layers.0.conv.bias"):
        assert torch.equal(sd3[k], m3.state_dict()[k]), k
    # a combination no checkpoint uses is refused loudly, not mis-computed
    d4 = tmp_path / "odd"
    Wav2Vec2Config(feat_extract_norm="group", conv_bias=True).save_pretrained(str(d4))
    with pytest.raises(NotImplementedError):
        load_checkpoint(str(d4))


def test_preset_launcher_builds_the_reference_presets():
    """scripts/suta.sh = REF/scripts/{LS,CH,CV,TD}.sh as one launcher: same flags, LibriSpeech at the three noise levels."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["bash", os.path.join(root, "scripts", "suta.sh"), "LS", "/data/LibriSpeech", "--", "--num_utts", "5"],
                       env=dict(os.environ, DRY_RUN="1", BATCH_UTTS="0"), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cmds = [ln for ln in r.stderr.splitlines() if ln.startswith("+ ")]
    assert [c.split("--extra_noise ")[1].split()[0] for c in cmds] == ["0", "0.005", "0.01"]
    for c in cmds:
        for flag in ("--asr facebook/wav2vec2-base-960h", "--dataset_name librispeech", "--dataset_dir /data/LibriSpeech", "--steps 10",
                     "--episodic", "--lr 2e-5", "--temp 2.5", "--em_coef 0.3", "--reweight", "--non_blank", "--train_feature",
                     "--log_dir exps", "--num_utts 5"):
            assert flag in c, (flag, c)
        assert "--batch_utts" not in c
    r = subprocess.run(["bash", os.path.join(root, "scripts", "suta.sh"), "TD", "/data/ted", "0.01"], env=dict(os.environ, DRY_RUN="1"),
                       capture_output=True, text=True)
    cmds = [ln for ln in r.stderr.splitlines() if ln.startswith("+ ")]
    assert len(cmds) == 1 and "--dataset_name ted" in cmds[0] and "--extra_noise 0.01" in cmds[0] and "--batch_utts 64" in cmds[0]
    assert subprocess.run(["bash", os.path.join(root, "scripts", "suta.sh"), "XX", "/d"], capture_output=True).returncode == 2
