"""GPU: the whole adaptation path against the oracle (live, tiny model) and the reference's golden vectors (base).

Stated tolerances (bf16 tensor-core operands, fp32 accumulation / norms / softmax / loss / optimizer):
  logits          max |diff| < 0.05   (logit std ~0.6; observed ~0.02 on wav2vec2-base, 5 s)
  per-step loss   relative   < 1e-3   (observed ~1e-5)
  adapted params  relative   < 1e-3   in value (observed ~2e-5); the 10-step parameter DELTA is reproduced to ~10-20 %
                                      in norm (Adam's first steps are sign-like, so bf16 gradient noise shows there)
  CTC decode      bit-exact given the same logits
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import e2e_checks
    return e2e_checks


def _assert_parity(m):
    assert m["logits0_maxabs"] < 0.05 and m["logitsN_maxabs"] < 0.05
    assert m["loss_rel_max"] < 1e-3
    assert m["param_value_rel"] < 1e-3
    assert m["param_delta_rel"] < 0.5
    assert m["decode_equal_given_logits"]


def test_forward_stages_vs_oracle(E):
    for name, rel in E.check_tiny_stages().items():
        assert rel < 0.03, name


def test_batched_adaptation_matches_per_utterance_oracle(E):
    for name, m in E.check_tiny_batch().items():
        _assert_parity(m)


@pytest.mark.parametrize("case", ["tiny_ln", "tiny_short", "base_ln_5s", "base_ln_5s_noblank"])
def test_against_reference_golden_vectors(E, case):
    m = E.check_golden(case)
    _assert_parity(m)
    assert m["argmax_agree0"] > 0.95


def test_train_feature_batched_vs_oracle(E):
    """--train_feature: per-utterance CNN/projection weights and the reference's duplicate-parameter Adam semantics."""
    for name, m in E.check_tiny_feat_batch().items():
        _assert_parity(m)
        assert m["param_delta_rel"] < 0.15           # the parameter DELTA itself is reproduced here (observed 3-5 %)
        for k, v in m.items():
            if k.startswith("delta:"):
                assert v < 0.25, (name, k, v)


@pytest.mark.parametrize("case", ["tiny_feat", "base_feat_2s"])
def test_train_feature_against_reference_golden_vectors(E, case):
    m = E.check_golden(case)
    _assert_parity(m)
    assert m["param_delta_rel"] < 0.15 and m["dlogits_rel"] < 0.15 and m["argmax_agree0"] > 0.95


def test_batch_composition_does_not_change_results(E):
    """An utterance adapted alone and inside a batch of other lengths gives the same result (utterance independence)."""
    from oracle import suta_oracle as O
    sd = O.init_weights(O.W2V2Config.tiny(), 3, blank_bias=0.5, ln_jitter=0.1)
    w = [O.synth_audio(n, s) for n, s in ((9000, 1), (4000, 2), (12345, 3))]
    alone = E.run_engine("tiny", sd, [w[0]], 5)[0]
    mixed = E.run_engine("tiny", sd, [w[1], w[0], w[2]], 5)[1]
    assert np.abs(alone["logits"][5] - mixed["logits"][5]).max() < 2e-3
    assert alone["ids"][5] == mixed["ids"][5] or np.abs(alone["logits"][5] - mixed["logits"][5]).max() < 2e-3


def test_drop_in_api_single_utterance(E, capsys):
    """forward_and_adapt & co. (the reference's function surface) drive the same engine path."""
    from oracle import suta_oracle as O
    from suta_b200 import ModelConfig, api
    ocfg = O.W2V2Config.tiny()
    sd = O.init_weights(ocfg, 3, blank_bias=0.5, ln_jitter=0.1)
    wav = O.synth_audio(12000, 11)
    model = api.configure_model(api.SutaModel(ModelConfig.tiny(), sd))
    params, names = api.collect_params(model, False, False, False, True)
    opt, sched = api.setup_optimizer(params, "AdamW", 2e-5)
    snap = api.copy_model_and_optimizer(model, opt, sched)
    x = torch.from_numpy(O.normalize_audio(wav))[None].cuda()
    ref = O.adapt_utterance(ocfg, sd, O.normalize_audio(wav), steps=3)
    for rep in range(2):                      # second pass checks the episodic restore
        model, opt, sched = api.load_model_and_optimizer(model, opt, *snap)
        out0 = model(x).logits
        assert out0.shape == (1, ref.logits0.shape[0], 32)
        assert np.abs(out0[0].cpu().numpy() - ref.logits0).max() < 0.05
        for i in range(3):
            out = api.forward_and_adapt(x, model, opt, 0.3, True, 2.5, True, sched, 0)
        assert np.abs(out[0].cpu().numpy() - ref.logits[3]).max() < 0.05
        got = {p.name: p.data[0].cpu().numpy() for p in params}
        num = sum(float(((got[n] - ref.params[n]) ** 2).sum()) for n in got)
        den = sum(float((ref.params[n] ** 2).sum()) for n in got)
        assert (num / den) ** 0.5 < 1e-3
    ent = api.softmax_entropy(out / 2.5)
    assert np.allclose(ent.cpu().numpy(), O.softmax_entropy(torch.tensor(out.cpu().numpy()) / 2.5).numpy(), atol=1e-5)
    mc = api.mcc_loss(out / 2.5, True)
    assert abs(float(mc) - float(O.mcc_loss(torch.tensor(out.cpu().numpy()) / 2.5, True))) < 1e-5
    capsys.readouterr()


def test_smoke_entry_point():
    import __graft_entry__
    __graft_entry__.smoke()


@pytest.mark.parametrize("switches", ["SUTA_NO_GEMM2 SUTA_NO_TAIL_SPLIT", "SUTA_NO_POSCONV_TC SUTA_NO_FUSED_DGRAD SUTA_ATTN_FWD_V1"])
def test_alternative_cuda_paths_keep_parity(switches):
    """The debug switches of INTEGRATION.md select other CUDA kernels for the same operators (one-SM GEMM instead of CTA
    pairs, generic-GEMM positional conv, col2im conv dgrad): the golden-vector parity must hold on those paths too.
    The switches are read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    env = dict(os.environ)
    for s in switches.split():
        env[s] = "1"
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", os.path.join(here, "test_gpu_e2e.py"), "-k",
                        "golden_vectors and (base_ln_5s_noblank or base_feat_2s or tiny_feat)"], env=env, capture_output=True, text=True,
                       cwd=os.path.dirname(here), timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
