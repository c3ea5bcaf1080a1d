"""CPU study: how much of the parameter-delta / logit-change mismatch against the fp32 reference is intrinsic to
rounding the GEMM operands (bf16 / tf32 / split bf16)?  Emulates operand rounding inside the oracle's forward with
straight-through quantisers (RF: round in the forward; RB: round the gradient in the backward), then runs the
oracle's adaptation loop.  TEST/ANALYSIS TOOL -- imports the oracle, never used by the product path.

    python tools/precision_study.py tiny|base [fwd_mode] [bwd_mode] ...
"""
import os, sys, json, time
import numpy as np
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import suta_oracle as O


def rnd(x, mode):
    if mode == "fp32":
        return x
    if mode == "bf16":
        return x.bfloat16().float()
    if mode == "bf16x2":
        hi = x.bfloat16().float()
        return hi + (x - hi).bfloat16().float()
    if mode == "tf32":
        i = x.contiguous().view(torch.int32)
        i = (i + 0x1000) & ~0x1FFF
        return i.view(torch.float32)
    raise ValueError(mode)


class _RF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mode):
        return rnd(x, mode)
    @staticmethod
    def backward(ctx, g):
        return g, None


class _RB(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, mode):
        ctx.mode = mode
        return x.view_as(x)
    @staticmethod
    def backward(ctx, g):
        return rnd(g, ctx.mode), None


def forward_q(cfg, sd, x, fm, bm):
    """oracle.model_forward with operand rounding: fm for forward operands, bm for gradient operands."""
    rf = lambda t: _RF.apply(t, fm)
    rb = lambda t: _RB.apply(t, bm)
    lin = lambda h, w, b: rb(F.linear(rf(h), rf(w), b))
    h = x[:, None]
    for i, s in enumerate(cfg.conv_stride):
        w = sd[f"wav2vec2.feature_extractor.conv_layers.{i}.conv.weight"]
        if i == 0:
            h = F.conv1d(h, w, stride=s)
            h = F.group_norm(h, cfg.conv_dim[0], sd["wav2vec2.feature_extractor.conv_layers.0.layer_norm.weight"],
                             sd["wav2vec2.feature_extractor.conv_layers.0.layer_norm.bias"], eps=1e-5)
        else:
            h = rb(F.conv1d(rf(h), rf(w), stride=s))
        h = F.gelu(h)
    h = h.transpose(1, 2)
    C = cfg.conv_dim[-1]
    h = F.layer_norm(h, (C,), sd["wav2vec2.feature_projection.layer_norm.weight"],
                     sd["wav2vec2.feature_projection.layer_norm.bias"], cfg.layer_norm_eps)
    h = lin(h, sd["wav2vec2.feature_projection.projection.weight"], sd["wav2vec2.feature_projection.projection.bias"])
    K = cfg.num_conv_pos_embeddings
    pc = rb(F.conv1d(rf(h.transpose(1, 2)), rf(O.pos_conv_weight(sd)), sd["wav2vec2.encoder.pos_conv_embed.conv.bias"],
                     padding=K // 2, groups=cfg.num_conv_pos_embedding_groups))
    if K % 2 == 0:
        pc = pc[:, :, :-1]
    h = h + F.gelu(pc).transpose(1, 2)
    H = cfg.hidden_size
    h = F.layer_norm(h, (H,), sd["wav2vec2.encoder.layer_norm.weight"], sd["wav2vec2.encoder.layer_norm.bias"], cfg.layer_norm_eps)
    nh, hd = cfg.num_attention_heads, H // cfg.num_attention_heads
    for l in range(cfg.num_hidden_layers):
        p = f"wav2vec2.encoder.layers.{l}."
        T = h.shape[1]
        q = lin(h, sd[p + "attention.q_proj.weight"], sd[p + "attention.q_proj.bias"]).view(1, T, nh, hd).transpose(1, 2)
        k = lin(h, sd[p + "attention.k_proj.weight"], sd[p + "attention.k_proj.bias"]).view(1, T, nh, hd).transpose(1, 2)
        v = lin(h, sd[p + "attention.v_proj.weight"], sd[p + "attention.v_proj.bias"]).view(1, T, nh, hd).transpose(1, 2)
        s = rb((rf(q) @ rf(k).transpose(-1, -2))) * hd ** -0.5
        a = rb(rf(torch.softmax(s, dim=-1)) @ rf(v))
        a = a.transpose(1, 2).reshape(1, T, H)
        h = h + lin(a, sd[p + "attention.out_proj.weight"], sd[p + "attention.out_proj.bias"])
        h = F.layer_norm(h, (H,), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], cfg.layer_norm_eps)
        f = F.gelu(lin(h, sd[p + "feed_forward.intermediate_dense.weight"], sd[p + "feed_forward.intermediate_dense.bias"]))
        h = h + lin(f, sd[p + "feed_forward.output_dense.weight"], sd[p + "feed_forward.output_dense.bias"])
        h = F.layer_norm(h, (H,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], cfg.layer_norm_eps)
    return lin(h, sd["lm_head.weight"], sd["lm_head.bias"])


def adapt(cfg, sd, x, steps, fm, bm, lr=2e-5):
    allnames = O.collect_param_names(cfg, train_feature=TF)
    mult = {}
    for n in allnames:
        mult[n] = mult.get(n, 0) + 1
    names = list(mult)
    w = {k: v.clone() for k, v in sd.items()}
    for n in names:
        w[n].requires_grad_(True)
    st = {n: dict(m=torch.zeros_like(w[n]), v=torch.zeros_like(w[n]), step=0) for n in names}
    xt = torch.from_numpy(x)[None]
    with torch.no_grad():
        l0 = forward_q(cfg, w, xt, fm, bm)[0].numpy().copy()
    g0 = None
    losses = []
    for i in range(steps):
        lg = forward_q(cfg, w, xt, fm, bm)
        loss = O.suta_loss(lg, 0.3, True, 2.5, NB)
        grads = torch.autograd.grad(loss, [w[n] for n in names])
        losses.append(float(loss))
        if i == 0:
            g0 = {n: g.numpy().copy() for n, g in zip(names, grads)}
        with torch.no_grad():
            for n, g in zip(names, grads):
                st[n]["step"] = O.adam_update(w[n], g, st[n]["m"], st[n]["v"], st[n]["step"], lr, k=mult[n])
    with torch.no_grad():
        lN = forward_q(cfg, w, xt, fm, bm)[0].numpy().copy()
        lN_fp32 = O.model_forward(cfg, w, xt)[0].numpy().copy()
    return dict(losses=losses, l0=l0, lN=lN, lN_fp32=lN_fp32, params={n: w[n].detach().numpy().copy() for n in names}, g0=g0)


NB = bool(int(os.environ.get("NB", 1)))
TF = bool(int(os.environ.get("TF", 0)))


def rel(a, b):
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1]
    steps = int(os.environ.get("STEPS", 10))
    if which == "tiny":
        cfg = O.W2V2Config.tiny(); sd = O.init_weights(cfg, 3, blank_bias=0.5, ln_jitter=0.1); x = O.normalize_audio(O.synth_audio(12000, 11))
    else:
        cfg = O.W2V2Config.base(); sd = O.init_weights(cfg, 0, blank_bias=float(os.environ.get("BB", 1.75))); x = O.normalize_audio(O.synth_audio(int(os.environ.get("N", 80000)), 1234))
    ref = adapt(cfg, sd, x, steps, "fp32", "fp32")
    print("ref logit change rms", float(np.sqrt(np.mean((ref["lN"] - ref["l0"]) ** 2))), "logit rms", float(np.sqrt(np.mean(ref["l0"] ** 2))))
    modes = sys.argv[2:] or ["bf16/bf16", "fp32/bf16", "bf16/fp32", "tf32/tf32", "bf16x2/bf16x2"]
    for m in modes:
        fm, bm = m.split("/")
        t0 = time.time()
        r = adapt(cfg, sd, x, steps, fm, bm)
        num = sum(float(((r["params"][n] - ref["params"][n]) ** 2).sum()) for n in ref["params"])
        den = sum(float(((ref["params"][n] - sd[n].numpy()) ** 2).sum()) for n in ref["params"])
        gnum = sum(float(((r["g0"][n] - ref["g0"][n]) ** 2).sum()) for n in ref["g0"])
        gden = sum(float((ref["g0"][n] ** 2).sum()) for n in ref["g0"])
        lr_ = [abs(a - b) / abs(b) for a, b in zip(r["losses"], ref["losses"])]
        print(json.dumps(dict(mode=m, loss_rel_per_step=[float("%.2e" % v) for v in lr_], logitsN_maxabs=float(np.abs(r["lN"] - ref["lN"]).max()), logits0_rel=rel(r["l0"], ref["l0"]), grad0_rel=(gnum / gden) ** 0.5, param_delta_rel=(num / den) ** 0.5,
                              dlogits_rel=rel(r["lN"] - r["l0"], ref["lN"] - ref["l0"]),
                              dlogits_via_fp32_fwd_rel=rel(r["lN_fp32"] - ref["l0"], ref["lN"] - ref["l0"]), sec=round(time.time() - t0, 1))), flush=True)
