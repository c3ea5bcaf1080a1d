"""Drop-in replacement for the reference's main.py (REF/main.py): same function names, same CLI flags, same
per-utterance loop and result files -- the arithmetic runs on the B200 engine (suta_b200).

Differences that the offline setting forces: `--asr` takes `random-base` / `random-large` or a
LOCAL HuggingFace checkpoint directory (no network), and `--dataset_name synthetic` generates the
LibriSpeech-test-other-shaped set instead of reading a corpus from `--dataset_dir`.
`--batch_utts N` (extension) adapts N utterances per step with per-utterance parameters instead of one.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "test-time-adaptation-asr-suta_b200"))

import torch  # noqa: E402

from suta_b200.api import (collect_params, configure_model, copy_model_and_optimizer, div_loss,  # noqa: E402,F401
                           forward_and_adapt, load_model_and_optimizer, mcc_loss, setup_optimizer, softmax_entropy,
                           SutaModel)
from suta_b200.wer import wer  # noqa: E402

import argparse  # noqa: E402

if __name__ == '__main__':
    SAMPLE_RATE = 16000
    parser = argparse.ArgumentParser(description="TTA ASR")
    parser.add_argument('--asr', type=str, default="random-base")
    parser.add_argument('--steps', type=int, default=40)
    parser.add_argument('--episodic', action='store_true')
    parser.add_argument('--div_coef', type=float, default=0.)
    parser.add_argument('--opt', type=str, default='AdamW')
    parser.add_argument('--dataset_name', type=str, default='synthetic')
    parser.add_argument('--dataset_dir', type=str, default='')
    parser.add_argument('--split', default=['test-other'])
    parser.add_argument('--lr', type=float, default=1e-4)
    parser.add_argument('--em_coef', type=float, default=1.)
    parser.add_argument('--reweight', action='store_true')
    parser.add_argument('--bias_only', action='store_true')
    parser.add_argument('--train_feature', action='store_true')
    parser.add_argument('--train_all', action='store_true')
    parser.add_argument('--batch_size', type=int, default=1)
    parser.add_argument('--temp', type=float, default=2.5)
    parser.add_argument('--non_blank', action='store_true')
    parser.add_argument('--log_dir', type=str, default='./exps')
    parser.add_argument('--extra_noise', type=float, default=0.)
    parser.add_argument('--scheduler', default=None)
    parser.add_argument('--num_utts', type=int, default=2939, help='synthetic set size (extension)')
    parser.add_argument('--batch_utts', type=int, default=0, help='adapt this many utterances per step (extension)')
    args = parser.parse_args()

    asr, steps, episodic, opt = args.asr, args.steps, args.episodic, args.opt
    dataset_name, lr, em_coef, reweight = args.dataset_name, args.lr, args.em_coef, args.reweight
    batch_size, temp, non_blank, log_dir = args.batch_size, args.temp, args.non_blank, args.log_dir
    extra_noise, scheduler, div_coef = args.extra_noise, args.scheduler, args.div_coef
    bias_only, train_feature, train_all = args.bias_only, args.train_feature, args.train_all
    skip_short_thd = None
    train_LN = True

    exp_name = dataset_name+'_'+str(em_coef)+'_'+str(steps)+'_'+str(temp)+'_'+asr.split('/')[-1]+'_'+'non_blank'+str(non_blank)+'_noise_'+str(extra_noise)+'_rew_'+str(reweight)+'_div_'+str(div_coef)+'_bias_'+str(bias_only)+'_feat_'+str(train_feature)+'_all_'+str(train_all)+'_LN_'+str(train_LN)

    from suta_b200.config import ModelConfig
    from suta_b200.data import librispeech_shaped
    from suta_b200.text import CTCVocab
    from suta_b200.weights import load_checkpoint, random_state_dict
    if dataset_name != 'synthetic':
        raise SystemExit("only --dataset_name synthetic is available offline (corpus loaders: SURVEY.md 8f rank 4)")
    if batch_size != 1:
        raise SystemExit("--batch_size: the reference only works with 1 (REF/main.py:32); use --batch_utts for batching")
    dataset = librispeech_shaped(args.num_utts, extra_noise=extra_noise)
    vocab = CTCVocab()

    print('------------------------------------')
    print(f'exp: {exp_name}')
    print(f'eposidic? {episodic}')
    for k in ('lr', 'opt', 'steps', 'em_coef', 'reweight', 'batch_size', 'temp', 'non_blank', 'extra_noise', 'scheduler',
              'div_coef', 'bias_only', 'train_feature', 'train_all'):
        print(f'{k} = {getattr(args, k)}')
    print(f'train_LN = {train_LN}')

    if asr.startswith("random-"):
        cfg = getattr(ModelConfig, asr.split("-", 1)[1])()
        sd = random_state_dict(cfg, seed=0, blank_bias=1.75)
    else:
        cfg, sd = load_checkpoint(asr)
    model = SutaModel(cfg, sd, train_feature=train_feature)

    # set up for tent
    model = configure_model(model)
    params, param_names = collect_params(model, bias_only, train_feature, train_all, train_LN)
    optimizer, scheduler = setup_optimizer(params, opt, lr, scheduler=scheduler)
    if episodic:
        model_state, optimizer_state, scheduler_state = copy_model_and_optimizer(model, optimizer, scheduler)
    print(param_names)

    CK = (1, 3, 5, 10, 20, 40)
    transcriptions = {k: [] for k in (0,) + CK}
    gt_texts, durations, werrs = [], [], []
    count = 0

    if args.batch_utts > 1:
        # batched extension: many utterances per adaptation step, each with its own parameters
        # (independent utterances = the reference's --episodic semantics; carrying state between utterances serialises them)
        from suta_b200.runner import SutaRunner
        if not episodic:
            raise SystemExit("--batch_utts adapts independent utterances: pass --episodic (without it the reference carries "
                             "model and optimizer state from one utterance to the next, which cannot be batched)")
        hp = optimizer.hp                      # the optimizer built by setup_optimizer above (opt, lr, betas, weight decay)
        hp.em_coef, hp.temp, hp.reweight, hp.not_blank, hp.div_coef = em_coef, temp, reweight, non_blank, div_coef
        out = SutaRunner(model.engine, steps, hp, max_utts=args.batch_utts, vocab=vocab,
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
                model, optimizer, scheduler = load_model_and_optimizer(model, optimizer, model_state, optimizer_state,
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
                outputs = forward_and_adapt(input_values, model, optimizer, em_coef, reweight, temp, non_blank, scheduler, div_coef)
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
    with open(os.path.join(log_dir, exp_name), 'w') as f:
        f.write("\n".join(lines) + "\n")
        f.write(f'eposidic? {episodic}\n')
        f.write(f'lr = {lr}\noptim = {opt}\nstep = {steps}\nem_coef = {em_coef}\nreweight = {reweight}\n')
        f.write(f'batch size = {batch_size}\ntemperature = {temp}\nnon_blank = {str(non_blank)}\n')
        f.write(f'extra_noise = {extra_noise}\nscheduler = {str(args.scheduler)}\ndiv_coef = {str(div_coef)}\n')
        f.write(f'bias_only = {str(bias_only)}\ntrain_feature = {str(train_feature)}\ntrain_all = {str(train_all)}\n')
        f.write(f'train_LN = {str(train_LN)}\n')
    csv_path = os.path.join(log_dir, exp_name + '.csv')
    with open(csv_path, 'w') as f:      # same columns as the reference's pandas dump (REF/main.py:452-454)
        f.write(",duration,WERR\n")
        for i, (d, w_) in enumerate(zip(durations, werrs)):
            f.write(f"{i},{d},{w_}\n")
