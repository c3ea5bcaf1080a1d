"""The reference's driver (REF/main.py:219-455, REF/main_SDPL.py:213-442) over the B200 engine: same CLI flags, same
per-utterance loop, same stdout lines and result files.  main.py and main_SDPL.py at the repository root are the two
entry points; the function surface they re-export lives in suta_b200.api.

Offline differences: `--asr` takes `random-base` / `random-large` / `random-large_lv60` / `random-tiny` or a
LOCAL HuggingFace checkpoint directory (no network); `--dataset_name synthetic` (the default) generates the
LibriSpeech-test-other-shaped set, `librispeech` / `chime` / `ted` / `commonvoice` read a corpus from `--dataset_dir` the
way REF/corpus/*.py do (suta_b200/corpus.py).  Extensions: `--num_utts`, `--batch_utts N` (N utterances per adaptation step,
each with its own parameters; needs --episodic).
"""
from __future__ import annotations

import argparse
import os

import torch

from . import api
from .wer import wer

SAMPLE_RATE = 16000
CK = (1, 3, 5, 10, 20, 40)                  # REF/main.py:349-398


def build_parser(sdpl: bool) -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="TTA ASR")
    p.add_argument('--asr', type=str, default="random-base")
    p.add_argument('--steps', type=int, default=10 if sdpl else 40)            # REF/main_SDPL.py:217 / REF/main.py:223
    p.add_argument('--episodic', action='store_true')
    p.add_argument('--div_coef', type=float, default=0.)
    p.add_argument('--opt', type=str, default='Adam' if sdpl else 'AdamW')     # REF/main_SDPL.py:220 / REF/main.py:226
    p.add_argument('--dataset_name', type=str, default='synthetic')
    p.add_argument('--dataset_dir', type=str, default='')
    p.add_argument('--split', default=['test-other'])
    p.add_argument('--lr', type=float, default=1e-4)
    p.add_argument('--em_coef', type=float, default=1.)
    p.add_argument('--reweight', action='store_true')
    p.add_argument('--bias_only', action='store_true')
    p.add_argument('--train_feature', action='store_true')
    if not sdpl:
        p.add_argument('--train_all', action='store_true')
    p.add_argument('--batch_size', type=int, default=1)
    p.add_argument('--temp', type=float, default=2.5)
    p.add_argument('--non_blank', action='store_true')
    p.add_argument('--log_dir', type=str, default='./exps')
    p.add_argument('--extra_noise', type=float, default=0.)
    p.add_argument('--scheduler', default=None)
    if sdpl:
        p.add_argument('--pl_coef', type=float, default=1)
    p.add_argument('--num_utts', type=int, default=2939, help='synthetic set size (extension)')
    p.add_argument('--batch_utts', type=int, default=0, help='adapt this many utterances per step (extension)')
    return p


def main(argv=None, sdpl: bool = False):
    args = build_parser(sdpl).parse_args(argv)
    asr, steps, episodic, opt = args.asr, args.steps, args.episodic, args.opt
    dataset_name, lr, em_coef, reweight = args.dataset_name, args.lr, args.em_coef, args.reweight
    batch_size, temp, non_blank, log_dir = args.batch_size, args.temp, args.non_blank, args.log_dir
    extra_noise, scheduler, div_coef = args.extra_noise, args.scheduler, args.div_coef
    bias_only, train_feature = args.bias_only, args.train_feature
    train_all = False if sdpl else args.train_all
    train_LN = True
    # REF/main_SDPL.py:345-346 calls forward_and_adapt with pl_coef=1. and div_coef=0 whatever the flags say; --pl_coef
    # only reaches exp_name and the log.  Here the flag is honoured (default 1 = the reference's behaviour).
    pl_coef = float(args.pl_coef) if sdpl else 0.0

    stem = (dataset_name + '_' + str(em_coef) + '_' + str(steps) + '_' + str(temp) + '_' + asr.split('/')[-1] + '_' + 'non_blank'
            + str(non_blank) + '_noise_' + str(extra_noise) + '_rew_' + str(reweight) + '_div_' + str(div_coef) + '_bias_'
            + str(bias_only) + '_feat_' + str(train_feature))
    if sdpl:                                                             # REF/main_SDPL.py:272
        exp_name = stem + '_se_' + '_pl_' + str(args.pl_coef)
    else:                                                                # REF/main.py:267
        exp_name = stem + '_all_' + str(train_all) + '_LN_' + str(train_LN)

    from .config import ModelConfig
    from .data import librispeech_shaped
    from .text import CTCVocab
    from .weights import load_checkpoint, random_state_dict
    if batch_size != 1:
        raise SystemExit("--batch_size: the reference only works with 1 (REF/main.py:32); use --batch_utts for batching")
    if dataset_name == 'synthetic':
        dataset = librispeech_shaped(args.num_utts, extra_noise=extra_noise)
    else:                                                                # REF/main.py:299: load_dataset(split, name, dir, ...)
        from .corpus import create_dataset
        dataset = create_dataset(args.split, dataset_name, args.dataset_dir, batch_size, extra_noise)
    vocab = CTCVocab()

    print('------------------------------------')
    print(f'exp: {exp_name}')
    print(f'eposidic? {episodic}')
    for k in ('lr', 'opt', 'steps', 'em_coef', 'reweight', 'batch_size', 'temp', 'non_blank', 'extra_noise', 'scheduler',
              'div_coef', 'bias_only', 'train_feature'):
        print(f'{k} = {getattr(args, k)}')
    if sdpl:
        print(f'pl_coef = {args.pl_coef}')
    else:
        print(f'train_all = {train_all}')
        print(f'train_LN = {train_LN}')

    if asr.startswith("random-"):
        cfg = getattr(ModelConfig, asr.split("-", 1)[1])()
        # special tokens out of the greedy transcript, as in a trained model (the SDPL loss re-encodes the transcript
        # character by character: a literal "<s>" is a KeyError in REF/main_SDPL.py:199-203)
        sd = random_state_dict(cfg, seed=0, blank_bias=1.75, special_bias=-10.0 if sdpl else 0.0)
    else:
        cfg, sd = load_checkpoint(asr)
    model = api.SutaModel(cfg, sd, train_feature=train_feature, pseudo_label=sdpl, train_all=train_all)

    # set up for tent
    model = api.configure_model(model)
    params, param_names = api.collect_params(model, bias_only, train_feature, train_all, train_LN)
    optimizer, scheduler = api.setup_optimizer(params, opt, lr, scheduler=scheduler, gamma=0.85 if sdpl else 0.7)
    if episodic:
        model_state, optimizer_state, scheduler_state = api.copy_model_and_optimizer(model, optimizer, scheduler)
    print(param_names)

    transcriptions = {k: [] for k in (0,) + CK}
    gt_texts, durations, werrs = [], [], []
    count = 0

    if args.batch_utts > 1:
        # batched extension: many utterances per adaptation step, each with its own parameters
        # (independent utterances = the reference's --episodic semantics; carrying state between utterances serialises them)
        from .runner import SutaRunner
        if not episodic:
            raise SystemExit("--batch_utts adapts independent utterances: pass --episodic (without it the reference carries "
                             "model and optimizer state from one utterance to the next, which cannot be batched)")
        hp = optimizer.hp                      # the optimizer built by setup_optimizer above (opt, lr, betas, weight decay)
        hp.em_coef, hp.temp, hp.reweight, hp.not_blank, hp.div_coef, hp.pl_coef = em_coef, temp, reweight, non_blank, div_coef, pl_coef
        # --train_all makes every weight the utterance's own: the runner (sharding over ranks, device-side noise, gather) then
        # adapts batches of ONE utterance, as the reference does
        out = SutaRunner(model.engine, steps, hp, max_utts=1 if train_all else args.batch_utts, vocab=vocab,
                         sched_gamma=scheduler.gamma if scheduler is not None else None,
                         sched_step=scheduler.step_size if scheduler is not None else 1, extra_noise=extra_noise).run(dataset)
        for k, d in out["texts"].items():
            transcriptions[k] = [d[i] for i in sorted(d)]
        gt_texts = [u.text for u in dataset]
        durations = [u.duration for u in dataset]
        if 10 in out["texts"]:
            werrs = [wer(u.text, out["texts"][0][u.index]) - wer(u.text, out["texts"][10][u.index]) for u in dataset]
    else:
        for utt in dataset:
            wav = torch.from_numpy(utt.audio())
            input_values = ((wav - wav.mean()) / torch.sqrt(wav.var(unbiased=False) + 1e-7))[None].cuda()   # processor(...)
            duration = input_values.shape[1] / SAMPLE_RATE
            durations.append(duration)
            texts = [utt.text]
            if episodic:
                model, optimizer, scheduler = api.load_model_and_optimizer(model, optimizer, model_state, optimizer_state,
                                                                           scheduler_state, scheduler)
            # vanilla forward
            with torch.no_grad():
                outputs = model(input_values).logits
            ori_transcription = vocab.batch_to_text(model.engine.decode_ids())
            transcriptions[0] += ori_transcription
            ori_wer = wer(list(texts), list(ori_transcription))
            print("original WER: ", ori_wer)
            # SUTA
            for i in range(steps):
                outputs = api.forward_and_adapt(input_values, model, optimizer, em_coef, reweight, temp, non_blank, scheduler,
                                                div_coef, pl_coef=pl_coef)
                if episodic and (i + 1) in CK:
                    transcription = vocab.batch_to_text(model.engine.decode_ids())
                    ada_wer = wer(list(texts), list(transcription))
                    print(f"adapt-{i + 1} WER:  " if i + 1 < 10 else f"adapt-{i + 1} WER: ", ada_wer)   # REF/main.py:355-396
                    if i + 1 == 10:
                        werrs.append(ori_wer - ada_wer)
                    transcriptions[i + 1] += transcription
            del input_values
            gt_texts += texts

    print("asr:", asr)
    print(f'non-adapted count = {count}')
    print(f'dataset num = {len(dataset)}')
    lines = [f"original WER: {wer(gt_texts, transcriptions[0])}"]
    for k, need in ((1, 10), (3, 10), (5, 10), (10, 10), (20, 20), (40, 40)):
        if steps >= need and len(transcriptions[k]) == len(gt_texts):
            lines.append(f"TTA-{k} WER: {wer(gt_texts, transcriptions[k])}")
    print("\n".join(lines))
    print('------------------------------------')

    if not os.path.exists(log_dir):
        os.makedirs(log_dir)
    with open(os.path.join(log_dir, exp_name), 'w') as f:                # REF/main.py:424-450 / REF/main_SDPL.py:406-431
        f.write("\n".join(lines) + "\n")
        f.write(f'eposidic? {episodic}\n')
        f.write(f'lr = {lr}\noptim = {opt}\nstep = {steps}\nem_coef = {em_coef}\nreweight = {reweight}\n')
        f.write(f'batch size = {batch_size}\ntemperature = {temp}\nnon_blank = {str(non_blank)}\n')
        f.write(f'extra_noise = {extra_noise}\nscheduler = {str(args.scheduler)}\ndiv_coef = {str(div_coef)}\n')
        f.write(f'bias_only = {str(bias_only)}\ntrain_feature = {str(train_feature)}\n')
        if sdpl:
            f.write(f'pl_coef = {args.pl_coef}\n')
        else:
            f.write(f'train_all = {str(train_all)}\ntrain_LN = {str(train_LN)}\n')
    if not sdpl:                                                         # REF/main.py:452-454 (main_SDPL.py writes no CSV)
        with open(os.path.join(log_dir, exp_name + '.csv'), 'w') as f:   # the same bytes as pandas' DataFrame.to_csv
            f.write(",duration,WERR\n")
            for i, (d, w_) in enumerate(zip(durations, werrs)):
                f.write(f"{i},{d},{w_}\n")
