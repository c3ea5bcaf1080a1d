"""Adapt the same full-size batch twice and report which outputs are bit-equal (run under different SUTA_* switches to
bisect a source of run-to-run differences).  Usage: python tools/determinism_probe.py [feature|ln] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "test-time-adaptation-asr-suta_b200")]
import torch  # noqa: E402

from suta_b200 import AdaptHyper, ModelConfig, SutaEngine  # noqa: E402
from suta_b200.api import reference_multiplicities  # noqa: E402
from suta_b200.data import librispeech_shaped  # noqa: E402
from suta_b200.shard import bucket_batches  # noqa: E402
from suta_b200.weights import random_state_dict  # noqa: E402

tf = (sys.argv[1] if len(sys.argv) > 1 else "feature") == "feature"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = ModelConfig.base()
utts = librispeech_shaped(2939, seed=0)
frames = [cfg.frames(u.n_samples) for u in utts]
batch = bucket_batches(frames, list(range(len(utts))), 64, 36864)[22]
wavs = [utts[i].audio() for i in batch]
eng = SutaEngine(cfg, random_state_dict(cfg, 0, 1.75), train_feature=tf,
                 trainable_mult=reference_multiplicities(cfg, train_feature=tf))
hp = AdaptHyper()


def run():
    eng.begin_batch(wavs)
    eng.reset()
    eng.forward()
    out = {"logits0": eng.logits().clone(), "params_before": eng.params().clone()}
    for name in ("wav_norm", "conv0", "conv1", "conv6", "h0", "hE", "h1_0", "h2_0", "h2_11"):
        out[name] = eng.debug_buffer(name).clone()
    for s in range(steps):
        eng.loss_backward(hp)
        out[f"dlogits{s}"] = eng.dlogits().clone()
        out[f"grads{s}"] = eng.grads().clone()
        if tf and s == 0:
            for name in ["d_yfp", "d_feat", "dh0_pad"] + [f"dpre{l}" for l in range(6, -1, -1)] + [f"cpre{l}" for l in range(7)]:
                out[name] = eng.debug_buffer(name).clone()
        eng.optimizer_step(hp)
        eng.forward()
        out[f"logits{s + 1}"] = eng.logits().clone()
    out["params"] = eng.params().clone()
    return out


a, b = run(), run()
sw = {k: v for k, v in os.environ.items() if k.startswith("SUTA_")}
print("switches", sw, "mode", "feature" if tf else "ln", "frames", eng.total_frames)
for k in a:
    same = torch.equal(a[k], b[k])
    extra = ""
    if not same:
        d = (a[k].float() - b[k].float()).abs()
        extra = f" max|diff| {float(d.max()):.3e}, {int((d > 0).sum())} of {d.numel()} elements differ"
        if k.startswith("grads"):
            n = eng.n_params
            segs = [(name, off, size) for name, off, size in eng.segments]
            bad = [name for name, off, size in segs if not torch.equal(a[k][:, off:off + size], b[k][:, off:off + size])]
            extra += f"; segments: {bad[:6]}{'...' if len(bad) > 6 else ''} ({len(bad)} of {len(segs)})"
    print(f"  {k:10s} {'bit-equal' if same else 'DIFFERENT'}{extra}")
    if not same and k.startswith("dpre") and a[k].dim() == 2:
        idx = (a[k].float() != b[k].float()).nonzero()
        rows = idx[:, 0]
        print("     rows with differences:", len(torch.unique(rows)), "first", torch.unique(rows)[:12].tolist())
        print("     row % 256:", sorted(set((torch.unique(rows) % 256).tolist()))[:40])
        print("     cols: min", int(idx[:, 1].min()), "max", int(idx[:, 1].max()), "distinct", len(torch.unique(idx[:, 1])))
        l = int(k[4:])
        Ls = []
        for wv in wavs:                                  # valid rows of layer l per utterance; regions are 256-row aligned (+1 spare row)
            L = len(wv)
            for kk, ss in list(zip(cfg.conv_kernel, cfg.conv_stride))[:l + 1]:
                L = (L - kk) // ss + 1
            Ls.append(L)
        offs, o = [], 0
        for L in Ls:
            offs.append(o)
            o += (L + 1 + 255) & ~255
        for r in torch.unique(rows)[:24].tolist():
            u = max(i for i in range(len(offs)) if offs[i] <= r)
            print(f"     row {r}: utterance {u}, row-in-utterance {r - offs[u]} of L = {Ls[u]} (region {((Ls[u] + 1 + 255) & ~255)}), "
                  f"{'VALID' if r - offs[u] < Ls[u] else 'padding'}; differing cols {int((idx[:, 0] == r).sum())}")
        if l == 1:       # which run is right?  recompute d pre_1 rows from d pre_2, the utterance's conv2 weight and GELU'(pre_1)
            L2s, offs2, o2 = [], [], 0
            for wv in wavs:
                L = len(wv)
                for kk, ss in list(zip(cfg.conv_kernel, cfg.conv_stride))[:3]:
                    L = (L - kk) // ss + 1
                L2s.append(L)
                offs2.append(o2)
                o2 += (L + 1 + 255) & ~255
            wname = "wav2vec2.feature_extractor.conv_layers.2.conv.weight"
            woff, wsize = [(o_, s_) for n_, o_, s_ in eng.segments if n_ == wname][0]
            for r in torch.unique(rows)[:6].tolist():
                u = max(i for i in range(len(offs)) if offs[i] <= r)
                t = r - offs[u]
                j = t // 2
                W = a["params_before"][u, woff:woff + wsize].view(512, 3, 512).bfloat16().float()       # [Cout][tap][Cin]
                dY = a["dpre2"][offs2[u] + j].float()
                dYm = a["dpre2"][offs2[u] + j - 1].float() if j > 0 else torch.zeros_like(dY)
                acc = dY @ W[:, 1, :] if t % 2 else dY @ W[:, 0, :] + dYm @ W[:, 2, :]
                ref = (acc * a["cpre1"][r].float()).bfloat16().float()
                e1 = float((a[k][r].float() - ref).abs().max() / (ref.abs().max() + 1e-30))
                e2 = float((b[k][r].float() - ref).abs().max() / (ref.abs().max() + 1e-30))
                z = float((acc * 0).abs().max())
                print(f"     row {r} (t={t}, {'odd' if t % 2 else 'even'}): rel err vs recomputation: run1 {e1:.3e}  run2 {e2:.3e}; |ref|max {float(ref.abs().max()):.3e}")
        for r, c in idx[:2].tolist():
            print(f"     [{r},{c}] run1 {float(a[k][r, c]):.6e} run2 {float(b[k][r, c]):.6e}")
