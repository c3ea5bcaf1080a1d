"""Benchmark of the SUTA hot path: audio-seconds adapted per second for 10-step SUTA on wav2vec2-base.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA engine (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU cores

A "step" is one adaptation batch: up to 64 utterances drawn from the LibriSpeech-test-other-shaped synthetic set
(SURVEY.md 8d), each episodically reset, forwarded, adapted for 10 steps and decoded at the reference's checkpoints
(REF/main.py:319-402).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "test-time-adaptation-asr-suta_b200")]

METRIC = "audio_seconds_adapted_per_second"
UNIT = "audio-s/s"
SUTA_STEPS = 10
MAX_UTTS = 64
MAX_FRAMES = 36864


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="base", choices=["base", "large", "tiny"])
    ap.add_argument("--mode", default="feature", choices=["feature", "ln"],
                    help="feature = BASELINE.json configs[1] (--train_feature preset of REF/scripts/LS.sh); ln = LayerNorm-only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--max-utts", type=int, default=MAX_UTTS, help="utterances per adaptation batch")
    ap.add_argument("--max-frames", type=int, default=MAX_FRAMES, help="frames per adaptation batch")
    ap.add_argument("--cpu-seconds", type=float, default=5.0, help="duration of the CPU-baseline utterance")
    ap.add_argument("--suta-steps", type=int, default=10, help="adaptation steps per utterance (configs[4]: 20)")
    ap.add_argument("--extra-noise", type=float, default=0.0, help="REF/data.py:23 noise added before normalisation (configs[4]: 0.01)")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))), hbm=float(d["hbm_gbs"]),
                    source="MEASURED_PEAKS.json (bf16_tflops_sustained)")
    return dict(tflops=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.th = index, [], threading.Event(), None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def start(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag.set()
        if self.th:
            self.th.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(self.samples))


def utt_flops(cfg, n_samples, steps, train_feature=False):
    H, I, NL, V = cfg.hidden_size, cfg.intermediate_size, cfg.num_hidden_layers, cfg.vocab_size
    Ls, L = [], n_samples
    for k, s in zip(cfg.conv_kernel, cfg.conv_stride):
        L = (L - k) // s + 1
        Ls.append(L)
    T = Ls[-1]
    f_conv, cin = 0.0, 1
    for c, k, Lo in zip(cfg.conv_dim, cfg.conv_kernel, Ls):
        f_conv += 2.0 * c * cin * k * Lo
        cin = c
    f_lin = 2.0 * T * (4 * H * H + 2 * H * I)
    f_proj = 2.0 * T * cfg.conv_dim[-1] * H
    f_pos = 2.0 * T * H * (H // cfg.num_conv_pos_embedding_groups) * cfg.num_conv_pos_embeddings
    f_head = 2.0 * T * H * V
    enc_f = f_proj + f_pos + NL * (f_lin + 4.0 * T * T * H) + f_head
    enc_b = f_proj + f_pos + NL * (f_lin + 8.0 * T * T * H) + f_head
    if train_feature:   # CNN forward every step, dgrad (no input grad for conv0) + wgrad in every backward
        f_conv0 = 2.0 * cfg.conv_dim[0] * cfg.conv_kernel[0] * Ls[0]
        return (steps + 1) * (f_conv + enc_f) + steps * (enc_b + f_proj + 2 * f_conv - f_conv0)
    return f_conv + (steps + 1) * enc_f + steps * enc_b


def select_batches(utts, cfg, n_batches, offset=0, max_utts=MAX_UTTS, max_frames=MAX_FRAMES):
    from suta_b200.shard import bucket_batches
    frames = [cfg.frames(u.n_samples) for u in utts]
    batches = bucket_batches(frames, list(range(len(utts))), max_utts, max_frames)
    nb = len(batches)
    pick = [int(round(offset + (i + 0.5) * nb / n_batches)) % nb for i in range(n_batches)]
    return [[utts[i] for i in batches[j]] for j in pick]


# ======================================================================================================
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from suta_b200 import AdaptHyper, ModelConfig, SutaEngine
    from suta_b200.data import librispeech_shaped
    from suta_b200.runner import adapt_batch
    from suta_b200.text import CTCVocab
    from suta_b200.weights import random_state_dict

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = getattr(ModelConfig, args.model)()
    tf = args.mode == "feature"
    mult = None
    if tf:              # multiplicities of REF/main.py:62-103 with train_feature (conv x4, proj LN x3, projection x2)
        from suta_b200.api import reference_multiplicities
        mult = reference_multiplicities(cfg, train_feature=True)
    eng = SutaEngine(cfg, random_state_dict(cfg, seed=0, blank_bias=1.75), train_feature=tf, trainable_mult=mult)
    hp, vocab = AdaptHyper(), CTCVocab()
    utts = librispeech_shaped(2939, seed=rank, extra_noise=args.extra_noise)           # weak scaling: every rank adapts its own draw of the set
    K, W = args.steps, max(args.warmup, 0)
    timed = select_batches(utts, cfg, K, max_utts=args.max_utts, max_frames=args.max_frames)
    warm = select_batches(utts, cfg, max(W, 1), offset=0.25, max_utts=args.max_utts, max_frames=args.max_frames)[:W]

    def stage(batch):                                     # synthetic-data generation is outside the timed region
        lens = np.asarray([u.n_samples for u in batch], dtype=np.int32)
        eng.begin_batch_lengths(lens)
        host = torch.zeros(eng.total_samples, dtype=torch.float32).pin_memory()
        hv = host.numpy()
        for u, o in zip(batch, eng.sample_off):
            hv[o:o + u.n_samples] = u.audio()
        return lens, host

    staged_w = [stage(b) for b in warm]
    staged = [stage(b) for b in timed]
    dev_audio = [h.cuda() for _, h in staged]

    def one_step(lens, audio):
        eng.begin_batch_lengths(lens)
        return adapt_batch(eng, audio, lens, SUTA_STEPS, hp, vocab)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_region(items):
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(items) + 1)]
        l0 = eng.launch_count
        ev[0].record()
        for i, (lens, audio) in enumerate(items):
            one_step(lens, audio)
            ev[i + 1].record()
        barrier()
        ms = torch.tensor([ev[0].elapsed_time(ev[-1])], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(len(items))]
        return float(ms.item()), eng.launch_count - l0, per_step

    for lens, host in staged_w:                           # warm-up (>= 3 by default)
        one_step(lens, host)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, launches, step_ms = timed_region([(l, a) for (l, _), a in zip(staged, dev_audio)])   # inputs resident in HBM
    ms_e2e, _, step_ms_e2e = timed_region([(l, h) for l, h in staged])                          # pinned host -> device
    clocks = sampler.stop() if rank == 0 else None

    # roofline leg: same steps with CUDA-event pairs around every tcgen05 GEMM launch
    eng.profile(True)
    for (l, _), a in zip(staged, dev_audio):
        one_step(l, a)
    gemm_ms, gemm_n, gemm_fl = eng.profile(False)
    rep = eng.profile_report()
    tot_ms = sum(v[0] for v in rep.values()) or 1.0
    breakdown = [{"kernel": k, "share": round(v[0] / tot_ms, 4), "ms_per_step": round(v[0] / K, 3), "launches_per_step": v[2] / K,
                  "tflops": round(v[1] / (v[0] * 1e-3) / 1e12, 1) if v[1] else None}
                 for k, v in sorted(rep.items(), key=lambda kv: -kv[1][0])]

    audio_s = sum(u.duration for b in timed for u in b)
    tot = torch.tensor([audio_s, float(sum(len(b) for b in timed)),
                        sum(utt_flops(cfg, u.n_samples, SUTA_STEPS, tf) for b in timed for u in b)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    audio_all, utts_all, flops_all = (float(x) for x in tot.tolist())
    peaks = measured_peaks()
    h2d = sum(h.numel() * 4 for _, h in staged) / K
    n_dec = 1 + sum(1 for c in (1, 3, 5, 10, 20, 40) if c <= SUTA_STEPS)      # REF/main.py:331-398 decode points
    d2h = sum(sum(cfg.frames(u.n_samples) for u in b) * 4 * n_dec + len(b) * 4 * n_dec for b in timed) / K    # collapsed ids + lengths
    out = {
        "metric": METRIC, "value": audio_all / (ms_dev / 1e3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic (0.1*randn audio, random-init wav2vec2-base weights; no checkpoint/dataset offline)",
        "config": {"workload": f"wav2vec2-{args.model} CTC, LibriSpeech-test-other-shaped synthetic set (2939 utts, 2-35 s), "
                               f"{SUTA_STEPS}-step EM+MCC SUTA, " + ("train_feature (LayerNorm + CNN front end + projection adapted per utterance)"
                                                                if tf else "LayerNorm-only") + ", "
                               f"<= {args.max_utts} utts / {args.max_frames} frames per adaptation batch, length-bucketed",
                   "utts_per_step": utts_all / K / world, "audio_s_per_step": audio_all / K / world,
                   "l2": "every step works on a different batch; workspace per step (GBs) >> 126 MB L2",
                   "suta": {"steps": SUTA_STEPS, "em_coef": hp.em_coef, "temp": hp.temp, "reweight": hp.reweight,
                            "non_blank": hp.not_blank, "lr": hp.lr, "opt": hp.opt, "episodic": True}},
        "e2e": {"value": audio_all / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "gemm2_kernel (cta_group::2 pairs) + gemm_bf16_tc_kernel (tcgen05, all dense contractions)",
                     "achieved": gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None, "peak": peaks["tflops"],
                     "unit": "TFLOP/s", "frac": (gemm_fl / (gemm_ms * 1e-3) / 1e12 / peaks["tflops"]) if gemm_ms else None,
                     "traffic": NCU_GEMM_DRAM_BYTES_PER_LAUNCH, "traffic_note": NCU_GEMM_TRAFFIC_NOTE,
                     "peak_source": peaks["source"], "launches": gemm_n,
                     "avg_launch_us": gemm_ms * 1e3 / gemm_n if gemm_n else None,
                     "gemm_share_of_step": gemm_ms / ms_dev if ms_dev else None},
        "path": {"algorithmic_tflop_per_step": flops_all / K / 1e12,
                 "achieved_tflops": flops_all / (ms_dev * 1e-3) / 1e12 / world,
                 "frac_of_peak": flops_all / (ms_dev * 1e-3) / 1e12 / world / peaks["tflops"]},
        "rtf_p50": None,
        "breakdown": breakdown,
    }
    # p50 RTF: per-utterance latency (wall time of the batch that carried it, end to end) / its duration, this rank
    rtfs = sorted((t / 1e3) / u.duration for b, t in zip(timed, step_ms_e2e) for u in b)
    out["rtf_p50"] = rtfs[len(rtfs) // 2]
    out["rtf_throughput"] = (ms_e2e / 1e3) / (audio_all / world)      # GPU-seconds per audio-second
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args.cpu_seconds, args.model, train_feature=tf)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per tcgen05 GEMM launch, from the committed `ncu --set full` capture
# (profiles/r01g_ncu_full_pair_gemm_posconv.md: 10 launches of the CTA-pair kernel in tools/profile_step.py --utts 32
# --seconds 6 --mode feature, 9.8 k frames per batch, i.e. about 1/2 of a bench batch; 72.6 MB on average).  Reads equal the
# operand bytes -- e.g. the FFN dgrad launch reads 80.3 MB for 15 MB dY + 60 MB GELU' + 4.7 MB W, the accumulate launch
# 95.3 MB for 60 MB dH + 4.7 MB W + 30 MB fp32 reduce target -- so no re-reads; results are mostly still in L2 at kernel end.
NCU_GEMM_DRAM_BYTES_PER_LAUNCH = 72.6e6
NCU_GEMM_TRAFFIC_NOTE = ("ncu capture of the 9.8k-frame profile batch (profiles/r01g_ncu_full_pair_gemm_posconv.md), not of this run; "
                         "tensor-bound kernel, traffic ~= algorithmic bytes")


def cpu_baseline(seconds, model="base", threads=None, train_feature=False):
    """The oracle (CPU restatement of the reference loop, pinned to the reference by tests/golden) on the host cores."""
    import torch
    from oracle import suta_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg = getattr(O.W2V2Config, model)()
    sd = O.init_weights(cfg, 0, blank_bias=1.75)
    n = int(seconds * 16000)
    x = O.normalize_audio(O.synth_audio(n, 1234))
    t0 = time.time()
    O.adapt_utterance(cfg, sd, x, steps=SUTA_STEPS, train_feature=train_feature)
    dt = time.time() - t0
    mode = "train_feature" if train_feature else "LayerNorm-only"
    return {"value": seconds / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"one {seconds:g} s utterance, {SUTA_STEPS}-step {mode} SUTA, fp32 torch CPU oracle "
                      f"(reset + vanilla forward + 10 x (forward, backward, AdamW, forward)), {dt:.1f} s wall"}


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm (oracle port; /root/reference does not travel to the box)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import torch
    from oracle import suta_oracle as O
    threads = os.cpu_count()
    torch.set_num_threads(threads)
    cfg = getattr(O.W2V2Config, args.model)()
    sd = O.init_weights(cfg, 0, blank_bias=1.75)
    seconds = args.cpu_seconds
    n = int(seconds * 16000)
    xs = [O.normalize_audio(O.synth_audio(n, 100 + i)) for i in range(args.warmup + args.steps)]
    tf = args.mode == "feature"
    mode = "train_feature (LayerNorm + CNN front end + projection)" if tf else "LayerNorm-only"
    for i in range(args.warmup):
        O.adapt_utterance(cfg, sd, xs[i], steps=SUTA_STEPS, train_feature=tf)
    t0 = time.time()
    for i in range(args.steps):
        O.adapt_utterance(cfg, sd, xs[args.warmup + i], steps=SUTA_STEPS, train_feature=tf)
    dt = time.time() - t0
    val = seconds * args.steps / dt
    sample = (f"each step = ONE {seconds:g} s utterance of the same synthetic generator (bounded sample of the batch the "
              f"CUDA arm adapts per step), {SUTA_STEPS}-step {mode} SUTA, fp32 torch on {threads} host threads")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (0.1*randn audio, random-init wav2vec2 weights)",
        "config": {"workload": f"wav2vec2-{args.model} CTC, LibriSpeech-test-other-shaped synthetic set, {SUTA_STEPS}-step EM+MCC SUTA, "
                               f"{mode}, reference algorithm on the host CPU", "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


if __name__ == "__main__":
    a = parse()
    SUTA_STEPS = a.suta_steps
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
