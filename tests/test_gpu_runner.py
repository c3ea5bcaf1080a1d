"""GPU: the dataset-level driver (REF/main.py:319-454) -- SutaRunner's sharded plan, the main.py CLI in both the
sequential (reference loop through the api.* surface) and the batched form, and the result files it writes."""
import io
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")


def _tiny_engine(train_feature=False):
    from suta_b200 import ModelConfig, SutaEngine
    from suta_b200.api import reference_multiplicities
    from suta_b200.weights import random_state_dict
    cfg = ModelConfig.tiny()
    return SutaEngine(cfg, random_state_dict(cfg, seed=0, blank_bias=0.5), train_feature=train_feature,
                      trainable_mult=reference_multiplicities(cfg, train_feature=train_feature))


@pytest.mark.parametrize("train_feature", [False, True])
def test_shards_merged_equal_the_unsharded_run(train_feature):
    """utterances are independent, so running shard k of n one after the other and merging transcripts and
    WER counters must reproduce the unsharded run (the N-GPU job without the GPUs)."""
    _need_gpu()
    from suta_b200.data import librispeech_shaped
    from suta_b200.runner import SutaRunner
    utts = librispeech_shaped(23, seed=3)
    eng = _tiny_engine(train_feature)
    whole = SutaRunner(eng, steps=5, max_utts=6, max_frames=4096).run(utts)
    merged_texts, merged_counts = {}, {}
    seen = []
    for r in range(3):
        part = SutaRunner(eng, steps=5, max_utts=4, max_frames=4096, rank=r, world_size=3).run(utts)
        for step, d in part["texts"].items():
            merged_texts.setdefault(step, {}).update(d)
        for step, (e, n) in part["wer_counts"].items():
            pe, pn = merged_counts.get(step, (0, 0))
            merged_counts[step] = (pe + e, pn + n)
        seen += sorted(part["texts"][0])
    assert sorted(seen) == list(range(len(utts)))                       # a partition: every utterance exactly once
    assert set(whole["texts"]) == {0, 1, 3, 5}                          # REF/main.py:349-398 checkpoints <= steps
    n_diff = 0
    for step in whole["texts"]:
        n_diff += sum(whole["texts"][step][i] != merged_texts[step][i] for i in whole["texts"][step])
    if not train_feature:
        # LayerNorm-only: every kernel's result for an utterance is independent of its position in the batch -> same bits
        assert n_diff == 0 and merged_counts == whole["wer_counts"]
    else:
        # train_feature: the conv0 / GroupNorm backward sums in chunks sized by the LONGEST utterance of the batch, so
        # the summation order (not the mathematics) depends on the batch; random-init logits are nearly tied, so a few
        # frames may decode differently.  Same utterances, same reference texts, transcripts equal but for such frames.
        assert n_diff <= 0.1 * len(utts) * len(whole["texts"]), n_diff
        for step, (e, n) in whole["wer_counts"].items():
            assert merged_counts[step][1] == n and abs(merged_counts[step][0] - e) <= 0.02 * n + 2
    eng.close()


def _run_main(tmp_path, *extra):
    os.makedirs(str(tmp_path), exist_ok=True)
    log_dir = str(tmp_path / "exps")
    cmd = [sys.executable, os.path.join(ROOT, "main.py"), "--asr", "random-tiny", "--num_utts", "7", "--steps", "10", "--episodic",
           "--em_coef", "0.3", "--reweight", "--lr", "2e-5", "--non_blank", "--temp", "2.5", "--log_dir", log_dir, *extra]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    files = sorted(os.listdir(log_dir))
    assert len(files) == 2 and files[1] == files[0] + ".csv"
    return r.stdout, files[0], open(os.path.join(log_dir, files[0])).read(), open(os.path.join(log_dir, files[1])).read()


def test_main_cli_writes_the_reference_result_files(tmp_path):
    """REF/main.py:267 (exp_name), :405-418 (stdout summary), :421-450 (log file), :452-454 (pandas CSV)."""
    _need_gpu()
    import pandas as pd
    out, name, log, csv = _run_main(tmp_path)
    assert name == ("synthetic_0.3_10_2.5_random-tiny_non_blankTrue_noise_0.0_rew_True_div_0.0_bias_False_feat_False_all_False_LN_True")
    lines = log.splitlines()
    assert [ln.split(":")[0] for ln in lines[:5]] == ["original WER", "TTA-1 WER", "TTA-3 WER", "TTA-5 WER", "TTA-10 WER"]
    for ln in lines[:5]:
        float(ln.split(": ")[1])
    assert lines[5:] == ["eposidic? True", "lr = 2e-05", "optim = AdamW", "step = 10", "em_coef = 0.3", "reweight = True",
                         "batch size = 1", "temperature = 2.5", "non_blank = True", "extra_noise = 0.0", "scheduler = None",
                         "div_coef = 0.0", "bias_only = False", "train_feature = False", "train_all = False", "train_LN = True"]
    for ln in lines[:5]:                                                 # the same summary goes to stdout (REF/main.py:408-415)
        assert ln in out
    assert out.count("original WER:  ") == 7 and out.count("adapt-10 WER: ") == 7 and "dataset num = 7" in out
    df = pd.read_csv(io.StringIO(csv), index_col=0)
    assert list(df.columns) == ["duration", "WERR"] and len(df) == 7
    assert df.to_csv() == csv                                            # byte-identical to what pandas (the reference) writes


def test_main_cli_batched_extension_matches_the_sequential_loop(tmp_path):
    """--batch_utts adapts many utterances per step; the corpus WERs must equal the one-utterance-at-a-time loop."""
    _need_gpu()
    _, _, log_seq, csv_seq = _run_main(tmp_path / "a")
    _, _, log_bat, csv_bat = _run_main(tmp_path / "b", "--batch_utts", "4")
    assert log_seq == log_bat          # LayerNorm-only results do not depend on the batch composition (same bits)
    assert csv_seq == csv_bat


def test_main_cli_rejects_batching_without_episodic(tmp_path):
    _need_gpu()
    cmd = [sys.executable, os.path.join(ROOT, "main.py"), "--asr", "random-tiny", "--num_utts", "3", "--steps", "3",
           "--batch_utts", "2", "--log_dir", str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode != 0 and "--episodic" in (r.stdout + r.stderr)


def test_main_cli_train_all(tmp_path):
    """`main.py --train_all` (REF/main.py:96-100, :267 exp_name, :449 log line): the whole model adapted per utterance, one
    utterance at a time (what the adaptation does to the model is checked by test_gpu_e2e.py::test_train_all_* and
    test_drop_in_api_train_all).  Through the runner (`--batch_utts`: sharding, gather) the batches hold ONE utterance each
    and the results are the sequential loop's."""
    _need_gpu()
    out, name, log, csv = _run_main(tmp_path, "--train_all")
    assert name.endswith("_feat_False_all_True_LN_True")
    assert "train_all = True" in log.splitlines()
    assert "wav2vec2.encoder.layers.1.attention.q_proj.weight" in out          # print(param_names), REF/main.py:311
    _, _, log_ln, _ = _run_main(tmp_path / "ln")
    assert log.splitlines()[0] == log_ln.splitlines()[0]                        # original WER: the same un-adapted model
    _, _, log_run, csv_run = _run_main(tmp_path / "runner", "--train_all", "--batch_utts", "4")
    assert log_run == log and csv_run == csv


def test_device_side_noise_and_truncation():
    """SURVEY.md 8f rank 1 (REF/data.py:19-23): waveforms are clamped to 600000 samples when the batch is laid out, and
    `extra_noise * randn` is added on the device before the normalisation: N(0, sigma^2), reproducible for a seed,
    different per utterance, and independent of the batch an utterance is adapted in."""
    _need_gpu()
    from oracle import suta_oracle as O
    eng = _tiny_engine()
    rng = np.random.default_rng(0)
    wavs = [(0.1 * rng.standard_normal(n)).astype(np.float32) for n in (610000, 40000, 40000)]
    eng.begin_batch(wavs)
    assert list(eng.lengths) == [600000, 40000, 40000]                     # truncated like REF/data.py:19-21
    raw = eng.debug_buffer("wav")[0].clone()
    eng.add_noise(0.01, seed=7, utt_ids=[5, 6, 9])
    noisy = eng.debug_buffer("wav")[0].clone()
    d = (noisy - raw).cpu().numpy()
    offs, lens = eng.sample_off, eng.lengths
    z = [d[o:o + n] / 0.01 for o, n in zip(offs, lens)]
    assert abs(z[0].mean()) < 0.01 and abs(z[0].std() - 1.0) < 0.01       # N(0,1) scaled by sigma
    assert abs(np.mean(z[0] ** 3)) < 0.02 and abs(np.mean(z[0] ** 4) - 3.0) < 0.05
    assert abs(np.corrcoef(z[1], z[2])[0, 1]) < 0.02                       # different utterance ids -> independent noise
    assert abs(np.corrcoef(z[0][:-1], z[0][1:])[0, 1]) < 0.01              # white
    eng.forward()
    x = eng.debug_buffer("wav_norm")[0].cpu().numpy()
    for o, n, w in zip(offs, lens, wavs):
        np.testing.assert_allclose(x[o:o + n], O.normalize_audio(noisy[o:o + n].cpu().numpy()), atol=2e-5)   # noise first, then HF:95
    # the same utterance id in another batch position / composition gets the same noise
    eng.begin_batch([wavs[2], wavs[1]])
    raw2 = eng.debug_buffer("wav")[0].clone()
    eng.add_noise(0.01, seed=7, utt_ids=[9, 6])
    d2 = (eng.debug_buffer("wav")[0] - raw2).cpu().numpy()
    o2 = eng.sample_off
    np.testing.assert_array_equal(d2[o2[0]:o2[0] + 40000], d[offs[2]:offs[2] + 40000])
    np.testing.assert_array_equal(d2[o2[1]:o2[1] + 40000], d[offs[1]:offs[1] + 40000])
    # another seed -> other noise
    eng.begin_batch([wavs[1]])
    raw3 = eng.debug_buffer("wav")[0].clone()
    eng.add_noise(0.01, seed=8, utt_ids=[6])
    d3 = (eng.debug_buffer("wav")[0] - raw3).cpu().numpy()[:40000]
    assert abs(np.corrcoef(d3, d[offs[1]:offs[1] + 40000])[0, 1]) < 0.02
    eng.close()


def test_main_sdpl_cli_writes_the_reference_log(tmp_path):
    """REF/main_SDPL.py: exp_name (:272), log file ending in `pl_coef = ...` (:406-431), no CSV; sequential and batched agree."""
    _need_gpu()
    logs = []
    for sub, extra in (("a", []), ("b", ["--batch_utts", "4"])):
        log_dir = str(tmp_path / sub)
        cmd = [sys.executable, os.path.join(ROOT, "main_SDPL.py"), "--asr", "random-tiny", "--num_utts", "6", "--steps", "10",
               "--episodic", "--lr", "1e-4", "--log_dir", log_dir, *extra]
        r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        files = os.listdir(log_dir)
        assert files == ["synthetic_1.0_10_2.5_random-tiny_non_blankFalse_noise_0.0_rew_False_div_0.0_bias_False_feat_False_se__pl_1"]
        logs.append(open(os.path.join(log_dir, files[0])).read())
        assert "pl_coef = 1" in r.stdout and "optim = Adam" in logs[-1]
    lines = logs[0].splitlines()
    assert [ln.split(":")[0] for ln in lines[:5]] == ["original WER", "TTA-1 WER", "TTA-3 WER", "TTA-5 WER", "TTA-10 WER"]
    assert lines[-1] == "pl_coef = 1" and lines[-2] == "train_feature = False"
    assert logs[0] == logs[1]


def test_local_hf_checkpoint_runs_through_the_engine(tmp_path):
    """`--asr <local HF directory>` (the offline stand-in for from_pretrained, REF/main.py:302-303): weights written by
    save_pretrained go through load_checkpoint into the engine; logits match the HF module's own CPU forward."""
    _need_gpu()
    from transformers import Wav2Vec2ForCTC
    from oracle import suta_oracle as O
    from suta_b200 import api
    from suta_b200.weights import load_checkpoint
    ocfg = O.W2V2Config.tiny()
    hf = Wav2Vec2ForCTC(ocfg.to_hf()).eval()
    hf.load_state_dict(O.init_weights(ocfg, 7, blank_bias=0.5, ln_jitter=0.1), strict=False)
    hf.save_pretrained(str(tmp_path / "ckpt"))
    cfg, sd = load_checkpoint(str(tmp_path / "ckpt"))
    model = api.configure_model(api.SutaModel(cfg, sd))
    x = torch.from_numpy(O.normalize_audio(O.synth_audio(9000, 5)))[None]
    with torch.no_grad():
        ref = hf(x).logits[0].numpy()
    got = model(x.cuda()).logits[0].cpu().numpy()
    assert got.shape == ref.shape and np.abs(got - ref).max() < 0.05
    cmd = [sys.executable, os.path.join(ROOT, "main.py"), "--asr", str(tmp_path / "ckpt"), "--num_utts", "2", "--steps", "3", "--episodic",
           "--log_dir", str(tmp_path / "exps")]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert os.path.exists(str(tmp_path / "exps" / "synthetic_1.0_3_2.5_ckpt_non_blankFalse_noise_0.0_rew_False_div_0.0_bias_False_feat_False_all_False_LN_True"))


def test_main_cli_reads_a_corpus_directory_with_an_lv60_model(tmp_path):
    """`--dataset_name ted --dataset_dir <tree>` walks a TED-LIUM-shaped directory of WAV segments
    like REF/corpus/ted.py (suta_b200/corpus.py), with a model of the lv60 family (LayerNorm feature extractor, pre-LN
    encoder); the batched extension reproduces the one-utterance-at-a-time loop bit for bit."""
    _need_gpu()
    import wave
    root = tmp_path / "ted"
    os.makedirs(root / "wav_segment"); os.makedirs(root / "transcription")
    rng = np.random.default_rng(0)
    for j, (n, text) in enumerate(((9000, "THE OF AND"), (14000, "HE WAS THAT IT HIS"), (6000, "A"), (11000, "WITH AS HAD FOR"))):
        with wave.open(str(root / "wav_segment" / f"seg_{j:03d}.wav"), "wb") as f:
            f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000)
            f.writeframes((rng.standard_normal(n) * 3000).astype("<i2").tobytes())
        (root / "transcription" / f"seg_{j:03d}.txt").write_text(text + "\n")
    logs = []
    for k, extra in enumerate(([], ["--batch_utts", "3"])):
        log_dir = str(tmp_path / f"exps{k}")
        cmd = [sys.executable, os.path.join(ROOT, "main.py"), "--asr", "random-tiny_lv60", "--dataset_name", "ted", "--dataset_dir", str(root),
               "--steps", "10", "--episodic", "--em_coef", "0.3", "--reweight", "--lr", "2e-5", "--non_blank", "--log_dir", log_dir, *extra]
        r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        assert "There are 4 samples" in r.stdout and "dataset num = 4" in r.stdout
        assert "conv_layers.6.layer_norm.weight" in r.stdout               # print(param_names): the conv LayerNorms are collected
        name = "ted_0.3_10_2.5_random-tiny_lv60_non_blankTrue_noise_0.0_rew_True_div_0.0_bias_False_feat_False_all_False_LN_True"
        logs.append(open(os.path.join(log_dir, name)).read())
        assert logs[-1].splitlines()[0].startswith("original WER: ") and logs[-1].splitlines()[4].startswith("TTA-10 WER: ")
    assert logs[0] == logs[1]
