"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference and transformers).  For each case it
  1. builds HF `Wav2Vec2ForCTC(cfg)` and loads the oracle's deterministic weights into it,
  2. drives the reference's own functions imported from /root/reference/main.py
     (configure_model, collect_params, setup_optimizer, copy/load_model_and_optimizer,
     forward_and_adapt) through the per-utterance loop of REF/main.py:319-402,
  3. decodes with the HF processor built offline from /root/reference/vocab.json,
  4. asserts the oracle restatement (oracle/suta_oracle.py) agrees, then
  5. writes small .npz fixtures (logits, losses, adapted params, transcripts).

Usage:  python tests/golden/make_golden.py [case ...]
"""
import contextlib
import io
import json
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import suta_oracle as O  # noqa: E402

REF = "/root/reference"

CASES = {
    # name: (cfg factory, n_samples, audio seed, weight seed, blank_bias, ln_jitter, steps, train_feature)
    "tiny_ln": ("tiny", 12000, 11, 3, 0.5, 0.1, 10, False),
    "tiny_feat": ("tiny", 9000, 12, 4, 0.35, 0.1, 10, True),
    "tiny_short": ("tiny", 2000, 13, 5, 0.35, 0.1, 5, False),
    "base_ln_5s": ("base", 80000, 1234, 0, 1.75, 0.0, 10, False),
    "base_ln_5s_noblank": ("base", 80000, 1234, 0, 0.0, 0.0, 3, False),
    "base_feat_2s": ("base", 32000, 77, 0, 1.75, 0.0, 3, True),
}
HYPER = dict(lr=2e-5, em_coef=0.3, reweight=True, temp=2.5, not_blank=True)   # REF/scripts/LS.sh:2-14
BIG = 1 << 16      # tensors above this many elements are stored as (checksum, head) only


def import_reference():
    stub = types.ModuleType("jiwer")
    stub.wer = lambda a, b: O.wer(a, b)
    sys.modules.setdefault("jiwer", stub)
    sys.path.insert(0, REF)
    import main as ref_main
    ref_main.scheduler = None           # module-global read by load_model_and_optimizer (REF/main.py:151)
    return ref_main


def build_processor():
    from transformers import Wav2Vec2CTCTokenizer, Wav2Vec2FeatureExtractor, Wav2Vec2Processor
    fe = Wav2Vec2FeatureExtractor(feature_size=1, sampling_rate=16000, padding_value=0.0, do_normalize=True,
                                  return_attention_mask=True)
    return Wav2Vec2Processor(fe, Wav2Vec2CTCTokenizer(os.path.join(REF, "vocab.json")))


def run_reference(ref_main, processor, cfg, sd, wav, steps, train_feature):
    from transformers import Wav2Vec2ForCTC
    torch.manual_seed(0)
    model = Wav2Vec2ForCTC(cfg.to_hf()).eval()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [m for m in missing if "masked_spec_embed" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = ref_main.configure_model(model)
        params, names = ref_main.collect_params(model, False, train_feature, False, True)
        optimizer, scheduler = ref_main.setup_optimizer(params, "AdamW", HYPER["lr"], scheduler=None)
        snap = ref_main.copy_model_and_optimizer(model, optimizer, scheduler)
        model, optimizer, scheduler = ref_main.load_model_and_optimizer(model, optimizer, *snap)
    x = processor([wav], return_tensors="pt", padding="longest").input_values
    out = dict(names=names, x=x[0].numpy().copy(), logits={}, texts={}, losses=[])
    with torch.no_grad():
        lg = model(x).logits
    out["logits"][0] = lg[0].numpy().copy()
    out["texts"][0] = processor.batch_decode(torch.argmax(lg, dim=-1))[0]
    for i in range(steps):
        with torch.no_grad():      # loss of the training forward, recomputed with the reference's functions
            lg = model(x).logits
            nb = torch.argmax(lg, -1) != 0
            e = ref_main.softmax_entropy(lg / HYPER["temp"])[nb].mean(0).mean()
            c = ref_main.mcc_loss(lg / HYPER["temp"], HYPER["reweight"], class_num=lg.shape[-1])
            out["losses"].append(float(e * HYPER["em_coef"] + c * (1 - HYPER["em_coef"])))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            lg = ref_main.forward_and_adapt(x, model, optimizer, HYPER["em_coef"], HYPER["reweight"], HYPER["temp"],
                                            HYPER["not_blank"], scheduler, 0)
        if (i + 1) in O.CHECKPOINT_STEPS:
            out["logits"][i + 1] = lg[0].detach().numpy().copy()
            out["texts"][i + 1] = processor.batch_decode(torch.argmax(lg, dim=-1))[0]
    msd = model.state_dict()
    out["params"] = {n: msd[n].detach().numpy().copy() for n in dict.fromkeys(names)}
    return out


def pack_param(a):
    a = np.asarray(a, dtype=np.float32)
    if a.size <= BIG:
        return a
    f = a.reshape(-1).astype(np.float64)
    return np.concatenate([[f.sum(), np.abs(f).sum(), (f * f).sum()], f[:4096]]).astype(np.float64)


def make(case, ref_main, processor):
    cfg_name, n, aseed, wseed, bb, jit, steps, tf = CASES[case]
    cfg = getattr(O.W2V2Config, cfg_name)()
    sd = O.init_weights(cfg, wseed, blank_bias=bb, ln_jitter=jit)
    wav = O.synth_audio(n, aseed)
    ref = run_reference(ref_main, processor, cfg, sd, wav, steps, tf)

    # --- pin the oracle against the reference --------------------------------------------
    x = O.normalize_audio(wav)
    np.testing.assert_allclose(x, ref["x"], rtol=0, atol=2e-6)
    names_o = O.collect_param_names(cfg, train_feature=tf)
    assert sorted(names_o) == sorted(ref["names"]), "collect_params multiplicities differ"
    ora = O.adapt_utterance(cfg, sd, x, steps=steps, train_feature=tf, **HYPER)
    assert ora.texts[0] == ref["texts"][0]
    np.testing.assert_allclose(ora.logits0, ref["logits"][0], rtol=0, atol=5e-5)
    np.testing.assert_allclose(ora.losses, ref["losses"], rtol=2e-5)
    for s, lg in ref["logits"].items():
        if s:
            np.testing.assert_allclose(ora.logits[s], lg, rtol=0, atol=2e-4)
            assert ora.texts[s] == ref["texts"][s], (s, ora.texts[s], ref["texts"][s])
    worst = 0.0
    for nme, p in ref["params"].items():
        d_ref = p - sd[nme].numpy()
        d_or = ora.params[nme] - sd[nme].numpy()
        worst = max(worst, float(np.abs(d_ref - d_or).max() / (np.abs(d_ref).max() + 1e-12)))
    # closed-form gradient vs autograd on the reference's step-0 logits
    lg0 = torch.tensor(ref["logits"][0][None], requires_grad=True)
    O.suta_loss(lg0, HYPER["em_coef"], HYPER["reweight"], HYPER["temp"], HYPER["not_blank"]).backward()
    lv, gcf = O.suta_loss_grad_closed(ref["logits"][0], HYPER["em_coef"], HYPER["reweight"], HYPER["temp"], HYPER["not_blank"])
    np.testing.assert_allclose(gcf, lg0.grad[0].numpy(), rtol=1e-3, atol=1e-9)
    np.testing.assert_allclose(lv, ref["losses"][0], rtol=1e-5)
    blank_frac = float((ref["logits"][0].argmax(-1) == 0).mean())
    print(f"[{case}] T={ref['logits'][0].shape[0]} blank_frac={blank_frac:.2f} losses={ref['losses'][0]:.6f}->"
          f"{ref['losses'][-1]:.6f} oracle-vs-ref worst delta mismatch={worst:.2e} texts={ref['texts']}")
    assert worst < 0.05, worst

    meta = dict(case=case, cfg=cfg_name, n_samples=n, audio_seed=aseed, weight_seed=wseed, blank_bias=bb,
                ln_jitter=jit, steps=steps, train_feature=tf, hyper=HYPER, blank_frac=blank_frac,
                names=ref["names"], texts={str(k): v for k, v in ref["texts"].items()},
                generator="tests/golden/make_golden.py", torch=torch.__version__,
                transformers=__import__("transformers").__version__)
    arrs = {"losses": np.asarray(ref["losses"], np.float64)}
    for s, lg in ref["logits"].items():
        arrs[f"logits_{s}"] = lg.astype(np.float32)
    for nme, p in ref["params"].items():
        arrs["param:" + nme] = pack_param(p)
    arrs["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, case + ".npz"), **arrs)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    ref_main = import_reference()
    processor = build_processor()
    for c in (sys.argv[1:] or list(CASES)):
        make(c, ref_main, processor)
