"""Numerical checks of every CUDA operator against fp32 torch / the oracle.  Each check returns a dict of error
metrics; tests/test_gpu_ops.py asserts on them and tools/gpu_diag.py prints them (for blind debugging on the box)."""
import ctypes as C
import math

import numpy as np
import torch

from oracle import suta_oracle as O
from suta_b200 import _lib
from suta_b200._lib import Hyper, check

DEV = "cuda"


def P(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def stream():
    return torch.cuda.current_stream().cuda_stream


def relerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def maxabs(a, b):
    return float((a.double() - b.double()).abs().max())


def gemm(a, b, M, N, K, a_rows=None, a_stride=None, out_f32=None, out_bf16=None, out_ld=None, bias=None, residual=None,
         act=0, aux_in=None, aux_out=None):
    lib = _lib.load()
    out_ld = out_ld or N
    check(lib.suta_op_gemm(P(a), a_rows if a_rows is not None else M, a_stride if a_stride is not None else K, P(b),
                           b.shape[0], K, M, N, K, P(out_f32), P(out_bf16), out_ld, P(bias), P(residual), out_ld, act,
                           P(aux_in), P(aux_out), out_ld, stream()))
    torch.cuda.synchronize()


def check_gemm_plain(M=300, N=256, K=192, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    b = torch.randn(N, K, device=DEV, generator=g).bfloat16()
    out = torch.full((M, N), float("nan"), device=DEV)
    gemm(a, b, M, N, K, out_f32=out)
    ref = a.float() @ b.float().t()
    return dict(rel=relerr(out, ref), maxabs=maxabs(out, ref), nan=int(torch.isnan(out).sum()))


def check_gemm_epilogue(M=517, N=384, K=256, seed=1):
    g = torch.Generator(device=DEV).manual_seed(seed)
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    b = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g)
    res = torch.randn(M, N, device=DEV, generator=g)
    pre_ref = a.float() @ b.float().t() + bias
    out = {}
    # bias + gelu (+ saved pre-activation), bf16 out
    o16 = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    aux = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    gemm(a, b, M, N, K, out_bf16=o16, bias=bias, act=1, aux_out=aux)
    out["gelu_rel"] = relerr(o16.float(), torch.nn.functional.gelu(pre_ref))
    xr = pre_ref.clone().requires_grad_(True)
    torch.nn.functional.gelu(xr).sum().backward()
    out["aux_rel"] = relerr(aux.float(), xr.grad)            # the forward epilogue saves GELU'(pre-activation)
    # bias + residual, fp32 + bf16 out
    o32 = torch.zeros(M, N, device=DEV)
    o16b = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    gemm(a, b, M, N, K, out_f32=o32, out_bf16=o16b, bias=bias, residual=res)
    out["res_rel"] = relerr(o32, pre_ref + res)
    out["res16_rel"] = relerr(o16b.float(), pre_ref + res)
    # gelu backward: acc * aux_in (the saved derivative)
    o32c = torch.zeros(M, N, device=DEV)
    gemm(a, b, M, N, K, out_f32=o32c, act=2, aux_in=aux)
    out["gelu_bwd_rel"] = relerr(o32c, (a.float() @ b.float().t()) * aux.float())
    out["gelu_bwd_vs_exact_rel"] = relerr(o32c, (a.float() @ b.float().t()) * xr.grad)
    # in-place accumulation (how the residual stream is updated): out = res; out += a b^T + bias
    o32d = res.clone()
    gemm(a, b, M, N, K, out_f32=o32d, bias=bias, act=4)
    out["acc_rel"] = relerr(o32d, pre_ref + res)
    # bf16 output alone through the dense path, ragged M
    o16c = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    gemm(a, b, M, N, K, out_bf16=o16c, bias=bias)
    out["bf16_rel"] = relerr(o16c.float(), pre_ref)
    o16d = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    gemm(a, b, M, N, K, out_bf16=o16d, act=2, aux_in=aux)
    out["gelu_bwd16_rel"] = relerr(o16d.float(), (a.float() @ b.float().t()) * aux.float())
    return out


def check_gemm_pair_tail(M=7700, N=768, K=1024, seed=21):
    """> 74 pair tiles with a small last wave: the CTA-pair kernel splits those tiles along K and reduce-adds the parts."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    b = (torch.randn(N, K, device=DEV, generator=g) / math.sqrt(K)).bfloat16()
    res = torch.randn(M, N, device=DEV, generator=g)
    out = res.clone()
    gemm(a, b, M, N, K, out_f32=out, act=4)
    ref = a.float() @ b.float().t() + res
    o16 = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    gemm(a, b, M, N, K, out_bf16=o16)
    return dict(acc_rel=relerr(out, ref), bf16_rel=relerr(o16.float(), a.float() @ b.float().t()))


def check_gemm_shapes():
    out = {}
    for (M, N, K) in [(249, 32, 768), (249, 768, 32), (1000, 768, 3072), (64, 2304, 768), (130, 64, 64), (5, 128, 128),
                      (4096, 512, 1536), (333, 48, 6144)]:
        r = check_gemm_plain(M, N, K, seed=M + N + K)
        out[f"{M}x{N}x{K}"] = r["rel"]
    return out


def check_gemm_window(R=700, CG=48, Kp=128, N=48, seed=3):
    """A = overlapping K-tap windows of a channels-last slab (the implicit-GEMM conv view)."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn(R + 8, CG, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, Kp * CG, device=DEV, generator=g) / math.sqrt(Kp * CG)).bfloat16()
    M = R - Kp + 1
    out = torch.zeros(M, N, device=DEV)
    gemm(x, w, M, N, Kp * CG, a_rows=M, a_stride=CG, out_f32=out)
    win = x[:R].float().unfold(0, Kp, 1).permute(0, 2, 1).reshape(M, Kp * CG)      # [M, (tap, c)]
    ref = win @ w.float().t()
    return dict(rel=relerr(out, ref))


def check_gemm_strided_conv(L=1001, C=64, k=3, s=2, N=64, seed=4):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn(L + 128, C, device=DEV, generator=g).bfloat16()
    w = (torch.randn(N, C, k, device=DEV, generator=g) / math.sqrt(k * C)).bfloat16()
    Lo = (L - k) // s + 1
    out = torch.zeros(Lo, N, device=DEV)
    wp = w.permute(0, 2, 1).reshape(N, k * C).contiguous()
    gemm(x, wp, Lo, N, k * C, a_rows=(L + 128 - k) // s + 1, a_stride=s * C, out_f32=out)
    ref = torch.nn.functional.conv1d(x[:L].float().t()[None], w.float(), stride=s)[0].t()
    return dict(rel=relerr(out, ref))


def check_gemm_mn(M=256, N=512, K=1000, seed=13):
    """MN-major operands: (A K-major, B [K,N]) and (A [K,M], B [K,N]) -- the dgrad-through-W and wgrad layouts."""
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(seed)
    a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
    at = a.t().contiguous()                                   # [K, M]
    bt = torch.randn(K, N, device=DEV, generator=g).bfloat16()    # [K, N]
    ref = a.float() @ bt.float()
    out = {}
    o = torch.zeros(M, N, device=DEV)
    check(lib.suta_op_gemm_mn(P(a), M, K, 0, P(bt), K, N, 1, M, N, K, P(o), N, stream()))
    torch.cuda.synchronize()
    out["kmajorA_mnB"] = relerr(o, ref)
    o2 = torch.zeros(M, N, device=DEV)
    check(lib.suta_op_gemm_mn(P(at), K, M, 1, P(bt), K, N, 1, M, N, K, P(o2), N, stream()))
    torch.cuda.synchronize()
    out["mnA_mnB"] = relerr(o2, ref)
    return out


def _row_utt(Ts):
    return torch.tensor(np.repeat(np.arange(len(Ts)), Ts), dtype=torch.int32, device=DEV)


def check_layernorm(N=768, Ts=(70, 3, 129), seed=5, bf16_in=False):
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(seed)
    M, U = sum(Ts), len(Ts)
    x = torch.randn(M, N, device=DEV, generator=g) * 2 + 0.3
    if bf16_in:
        x = x.bfloat16()
    n_par = 2 * N + 64
    Pm = torch.randn(U, n_par, device=DEV, generator=g)
    g_off, b_off = 64, 64 + N
    ru = _row_utt(Ts)
    y32 = torch.zeros(M, N, device=DEV); y16 = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    mean = torch.zeros(M, device=DEV); rstd = torch.zeros(M, device=DEV)
    check(lib.suta_op_layernorm_fwd(None if bf16_in else P(x), P(x) if bf16_in else None, P(ru), P(Pm), n_par, g_off, b_off,
                                    P(y32), P(y16), P(mean), P(rstd), M, N, 1e-5, stream()))
    dy = torch.randn(M, N, device=DEV, generator=g)
    G = torch.zeros(U, n_par, device=DEV)
    dx32 = torch.zeros(M, N, device=DEV); dx16 = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    tok_off = torch.tensor(np.concatenate([[0], np.cumsum(Ts)[:-1]]), dtype=torch.int64, device=DEV)
    Tt = torch.tensor(Ts, dtype=torch.int32, device=DEV)
    scratch = torch.empty(int(lib.suta_op_layernorm_bwd_scratch_floats(N, U)), device=DEV)
    G = torch.full((U, n_par), float("nan"), device=DEV)        # dgamma/dbeta are WRITTEN, the rest of G is left alone
    G2 = G.clone()
    for Gx in (G, G2):                                          # twice: the two-stage reduction is bit-reproducible
        check(lib.suta_op_layernorm_bwd(P(dy), None if bf16_in else P(x), P(x) if bf16_in else None, P(mean), P(rstd), P(ru),
                                        P(Pm), n_par, g_off, b_off, P(Gx), P(dx32), P(dx16), M, N, P(tok_off), P(Tt), U,
                                        P(scratch), stream()))
    torch.cuda.synchronize()
    seg = slice(g_off, b_off + N)
    bit_equal = bool(torch.equal(G[:, seg], G2[:, seg]))
    untouched = bool(torch.isnan(G[:, :g_off]).all())
    G = torch.where(torch.isnan(G), torch.zeros_like(G), G)
    xr = x.float().clone().requires_grad_(True)
    Pr = Pm.clone().requires_grad_(True)
    gam = Pr[ru.long(), g_off:g_off + N]; bet = Pr[ru.long(), b_off:b_off + N]
    yr = torch.nn.functional.layer_norm(xr, (N,), eps=1e-5) * gam + bet
    (yr * dy).sum().backward()
    return dict(y_rel=relerr(y32, yr), y16_rel=relerr(y16.float(), yr), dx_rel=relerr(dx32, xr.grad),
                dx16_rel=relerr(dx16.float(), xr.grad), dparam_rel=relerr(G, Pr.grad), dparam_bit_equal=bit_equal,
                dparam_writes_only_its_segment=untouched)


def check_layernorm_gelu(N=512, Ts=(70, 3, 129), gaps=(5, 0, 250), seed=9, dy_bf16=True):
    """GELU(LayerNorm(x)) forward / backward of the LayerNorm feature extractor (HF/modeling_wav2vec2.py:291-299) in a conv
    layer's row space: bf16 x, utterance regions separated by GAP rows (row_utt = -1) that must be left untouched, the
    gradient arriving in bf16 and replaced IN PLACE by d x (valid rows only), per-utterance gamma / beta."""
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(seed)
    U = len(Ts)
    offs, ru, r = [], [], 0
    for u, (T, gp) in enumerate(zip(Ts, gaps)):
        offs.append(r); ru += [u] * T + [-1] * gp; r += T + gp
    M = r
    ru = torch.tensor(ru, dtype=torch.int32, device=DEV)
    valid = ru >= 0
    x = (torch.randn(M, N, device=DEV, generator=g) * 2 + 0.3).bfloat16()
    n_par = 2 * N + 64
    Pm = torch.randn(U, n_par, device=DEV, generator=g) * 0.5 + 0.5
    g_off, b_off = 64, 64 + N
    y16 = torch.full((M, N), 7.0, device=DEV, dtype=torch.bfloat16)
    mean = torch.full((M,), float("nan"), device=DEV); rstd = torch.full((M,), float("nan"), device=DEV)   # gap rows stay NaN: must not leak
    check(lib.suta_op_layernorm_fwd_mode(None, P(x), P(ru), P(Pm), n_par, g_off, b_off, None, P(y16), P(mean), P(rstd), M, N, 1e-5,
                                         None, 1, stream()))
    dy = torch.randn(M, N, device=DEV, generator=g)
    dy[~valid] = 0
    dyb = dy.bfloat16() if dy_bf16 else dy.clone()
    dy_ref = dyb.float().clone()
    tok_off = torch.tensor(offs, dtype=torch.int64, device=DEV)
    Tt = torch.tensor(Ts, dtype=torch.int32, device=DEV)
    scratch = torch.empty(int(lib.suta_op_layernorm_bwd_scratch_floats(N, U)), device=DEV)
    G = torch.full((U, n_par), float("nan"), device=DEV)
    dx32 = torch.zeros(M, N, device=DEV) if not dy_bf16 else None
    check(lib.suta_op_layernorm_bwd_mode(None if dy_bf16 else P(dyb), P(dyb) if dy_bf16 else None, None, P(x), P(mean), P(rstd), P(ru),
                                         P(Pm), n_par, g_off, b_off, P(G), None if dy_bf16 else P(dx32), P(dyb) if dy_bf16 else None,
                                         None, 1, M, N, P(tok_off), P(Tt), U, P(scratch), stream()))
    torch.cuda.synchronize()
    dx = dyb.float() if dy_bf16 else dx32                     # bf16: written in place over dy
    G = torch.where(torch.isnan(G), torch.zeros_like(G), G)
    xr = x.float().clone().requires_grad_(True)
    Pr = Pm.clone().requires_grad_(True)
    idx = ru.long().clamp(min=0)
    gam = Pr[idx, g_off:g_off + N]; bet = Pr[idx, b_off:b_off + N]
    yr = torch.nn.functional.gelu(torch.nn.functional.layer_norm(xr, (N,), eps=1e-5) * gam + bet)
    (yr * dy_ref)[valid].sum().backward()
    return dict(y16_rel=relerr(y16.float()[valid], yr[valid]), gaps_untouched=bool((y16[~valid] == 7.0).all()) if (~valid).any() else True,
                dx_rel=relerr(dx[valid], xr.grad[valid]), dx_gaps_zero=bool((dx[~valid] == 0).all()) if (~valid).any() else True,
                dparam_rel=relerr(G, Pr.grad), nan=int(torch.isnan(dx).sum()) + int(torch.isnan(G).sum()))


def check_layernorm_keep_input(N=768, Ts=(70, 3, 129), seed=10):
    """The pre-LN encoder's LayerNorm (HF:638-645): forward y_bf16 = LN(x), y_f32 = x + bias (the residual stream seeded
    for the GEMM that accumulates the branch onto it); backward d x = d_residual + LayerNorm-backward(d branch), in place."""
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(seed)
    M, U = sum(Ts), len(Ts)
    x = torch.randn(M, N, device=DEV, generator=g) * 2 + 0.3
    n_par = 2 * N + 64
    Pm = torch.randn(U, n_par, device=DEV, generator=g)
    g_off, b_off = 64, 64 + N
    ru = _row_utt(Ts)
    bias = torch.randn(N, device=DEV, generator=g)
    y32 = torch.zeros(M, N, device=DEV); y16 = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    mean = torch.zeros(M, device=DEV); rstd = torch.zeros(M, device=DEV)
    check(lib.suta_op_layernorm_fwd_mode(P(x), None, P(ru), P(Pm), n_par, g_off, b_off, P(y32), P(y16), P(mean), P(rstd), M, N, 1e-5,
                                         P(bias), 2, stream()))
    dy = torch.randn(M, N, device=DEV, generator=g)
    dres = torch.randn(M, N, device=DEV, generator=g)
    dres0 = dres.clone()
    tok_off = torch.tensor(np.concatenate([[0], np.cumsum(Ts)[:-1]]), dtype=torch.int64, device=DEV)
    Tt = torch.tensor(Ts, dtype=torch.int32, device=DEV)
    scratch = torch.empty(int(lib.suta_op_layernorm_bwd_scratch_floats(N, U)), device=DEV)
    G = torch.full((U, n_par), float("nan"), device=DEV)
    dx16 = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    check(lib.suta_op_layernorm_bwd_mode(P(dy), None, P(x), None, P(mean), P(rstd), P(ru), P(Pm), n_par, g_off, b_off, P(G), P(dres),
                                         P(dx16), P(dres), 2, M, N, P(tok_off), P(Tt), U, P(scratch), stream()))
    torch.cuda.synchronize()
    G = torch.where(torch.isnan(G), torch.zeros_like(G), G)
    xr = x.clone().requires_grad_(True)
    Pr = Pm.clone().requires_grad_(True)
    gam = Pr[ru.long(), g_off:g_off + N]; bet = Pr[ru.long(), b_off:b_off + N]
    yr = torch.nn.functional.layer_norm(xr, (N,), eps=1e-5) * gam + bet
    (yr * dy).sum().backward()
    want = dres0 + xr.grad
    return dict(y32_rel=relerr(y32, x + bias), y16_rel=relerr(y16.float(), yr), dx_rel=relerr(dres, want),
                dx16_rel=relerr(dx16.float(), want), dparam_rel=relerr(G, Pr.grad))


def attn_table(Ts):
    tab, off = [], 0
    for T in Ts:
        for m0 in range(0, T, 128):
            tab.append((off, T, m0, 0))
        off += T
    return torch.tensor(tab, dtype=torch.int32, device=DEV)


def check_attention(Ts=(249, 64, 1, 130), heads=2, seed=6):
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(seed)
    H, M = heads * 64, sum(Ts)
    qkv = torch.randn(M, 3 * H, device=DEV, generator=g).bfloat16()
    dO = torch.randn(M, H, device=DEV, generator=g).bfloat16()
    tab = attn_table(Ts)
    O_ = torch.zeros(M, H, device=DEV, dtype=torch.bfloat16)
    lse = torch.zeros(heads, M, device=DEV)
    check(lib.suta_op_attention_fwd(P(qkv), P(O_), P(lse), P(tab), tab.shape[0], H, heads, M, stream()))
    D = torch.zeros(heads, M, device=DEV)
    dqkv = torch.zeros(M, 3 * H, device=DEV, dtype=torch.bfloat16)
    check(lib.suta_op_attention_bwd(P(qkv), P(O_), P(dO), P(lse), P(D), P(dqkv), P(tab), tab.shape[0], H, heads, M, stream()))
    torch.cuda.synchronize()
    x = qkv.float().clone().requires_grad_(True)
    outs, off = [], 0
    for T in Ts:
        q, k, v = [x[off:off + T, i * H:(i + 1) * H].view(T, heads, 64).transpose(0, 1) for i in range(3)]
        a = torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v
        outs.append(a.transpose(0, 1).reshape(T, H))
        off += T
    ref = torch.cat(outs, 0)
    (ref * dO.float()).sum().backward()
    return dict(o_rel=relerr(O_.float(), ref), dq_rel=relerr(dqkv[:, :H].float(), x.grad[:, :H]),
                dk_rel=relerr(dqkv[:, H:2 * H].float(), x.grad[:, H:2 * H]),
                dv_rel=relerr(dqkv[:, 2 * H:].float(), x.grad[:, 2 * H:]), nan=int(torch.isnan(dqkv.float()).sum()))


def check_loss(Ts=(249, 37, 6), seed=7, em_coef=0.3, reweight=True, not_blank=True, temp=2.5, blank_bias=1.0, div_coef=0.0):
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(seed)
    M, U = sum(Ts), len(Ts)
    logits = torch.randn(M, 32, device=DEV, generator=g) * 0.6
    logits[:, 0] += blank_bias
    off = torch.tensor(np.concatenate([[0], np.cumsum(Ts)[:-1]]), dtype=torch.int64, device=DEV)
    Tt = torch.tensor(Ts, dtype=torch.int32, device=DEV)
    loss = torch.zeros(3 * U, device=DEV)
    d32 = torch.zeros(M, 32, device=DEV); d16 = torch.zeros(M, 32, device=DEV, dtype=torch.bfloat16)
    check(lib.suta_op_loss(P(logits), P(off), P(Tt), U, em_coef, temp, int(reweight), int(not_blank), div_coef, P(loss), P(d32),
                           P(d16), stream()))
    torch.cuda.synchronize()
    lg = logits.cpu().numpy()
    worst_l = worst_g = 0.0
    nan_loss, finite_grad, nan_agree = 0, True, True
    o = 0
    for u, T in enumerate(Ts):
        lt = torch.tensor(lg[o:o + T][None], requires_grad=True)
        lref = O.suta_loss(lt, em_coef, reweight, temp, not_blank, div_coef)
        lref.backward()
        gk = d32[o:o + T].cpu().double()
        if not np.isfinite(float(lref)):
            # every frame blank with non_blank masking: REF/main.py:190 takes the mean of an empty selection -> the loss is
            # NaN; torch's backward then poisons the gradient with NaN too.  The kernel reports the NaN loss and keeps the
            # gradient finite (the entropy term contributes nothing): state which of the two happened
            nan_loss += 1
            nan_agree = nan_agree and not np.isfinite(float(loss[u]))
            finite_grad = finite_grad and bool(torch.isfinite(gk).all())
            o += T
            continue
        nan_agree = nan_agree and bool(np.isfinite(float(loss[u])))
        worst_l = max(worst_l, abs(float(loss[u]) - float(lref)) / abs(float(lref)))
        worst_g = max(worst_g, float((gk - lt.grad[0].double()).norm() / lt.grad[0].double().norm()))
        o += T
    return dict(loss_rel=worst_l, grad_rel=worst_g, bf16_rel=relerr(d16.float(), d32), nan_loss_utts=nan_loss,
                nan_where_the_reference_is_nan=nan_agree, grad_finite_where_loss_nan=finite_grad)


def check_ctc(Ts=(249, 37, 1, 6, 700), seed=10, blank_bias=1.0, all_blank_utt=1):
    """Pseudo-label CTC loss + gradient (csrc/ctc.cu) vs torch.nn.CTCLoss autograd on the CPU through the oracle's
    restatement of REF/main_SDPL.py:194-209 (log_softmax over TIME, target = stripped greedy transcript)."""
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(seed)
    M, U = sum(Ts), len(Ts)
    logits = torch.randn(M, 32, device=DEV, generator=g) * 1.5
    logits[:, 0] += blank_bias
    logits[:, 1:4] -= 20.0                                       # no literal <s> </s> <unk> in the transcript (KeyError in the reference)
    logits[:, 4] += 1.0                                          # plenty of word delimiters, also at the ends
    off_np = np.concatenate([[0], np.cumsum(Ts)[:-1]])
    if all_blank_utt is not None:
        o = off_np[all_blank_utt]
        logits[o:o + Ts[all_blank_utt], 0] += 50.0               # empty target
    rep = torch.rand(M, device=DEV, generator=g) < 0.3           # repeated frames -> repeated characters need the blank path
    lg = logits.cpu().numpy()
    r = rep.cpu().numpy()
    for i in range(1, M):
        if r[i]:
            lg[i] = lg[i - 1]
    logits = torch.tensor(lg, device=DEV)
    off = torch.tensor(off_np, dtype=torch.int64, device=DEV)
    Tt = torch.tensor(Ts, dtype=torch.int32, device=DEV)
    ids = torch.zeros(M, dtype=torch.int32, device=DEV); col = torch.zeros(M, dtype=torch.int32, device=DEV)
    ln = torch.zeros(U, dtype=torch.int32, device=DEV)
    check(lib.suta_op_decode(P(logits), P(off), P(Tt), U, 32, P(ids), P(col), P(ln), stream()))
    n_alpha = sum(T * (2 * T + 1) for T in Ts) + 2 * U + 8
    alpha = torch.empty(n_alpha, device=DEV); gbuf = torch.empty(M, 32, device=DEV)
    loss = torch.zeros(4 * U, device=DEV); d32 = torch.zeros(M, 32, device=DEV); tl = torch.zeros(U, dtype=torch.int32, device=DEV)
    outs = []
    for _ in range(2):                                           # twice: bit-reproducible
        check(lib.suta_op_ctc_pseudo_label(P(logits), P(off), P(Tt), U, P(col), P(ln), P(alpha), P(gbuf), P(loss), P(d32), P(tl),
                                           stream()))
        torch.cuda.synchronize()
        outs.append((loss.clone(), d32.clone()))
    worst_l = worst_g = 0.0
    tlen_ok = True
    o = 0
    for u, T in enumerate(Ts):
        lt = torch.tensor(lg[o:o + T][None], requires_grad=True)
        tgt = O.pseudo_label_targets(lt)
        lref = O.pseudo_labeling_loss(lt)
        lref.backward()
        tlen_ok = tlen_ok and int(tl[u]) == len(tgt)
        worst_l = max(worst_l, abs(float(loss[3 * U + u]) - float(lref)) / max(abs(float(lref)), 1e-6))
        gk = d32[o:o + T].cpu().double()
        worst_g = max(worst_g, float((gk - lt.grad[0].double()).norm() / (lt.grad[0].double().norm() + 1e-30)))
        o += T
    return dict(loss_rel=worst_l, grad_rel=worst_g, target_len_equal=tlen_ok, finite=bool(torch.isfinite(d32).all()),
                bit_equal=bool(torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])))


def check_adam(n=5000, U=3, steps=4, seed=8, opt="AdamW", lr=2e-5, beta1=0.9, wd=0.0, max_mult=4):
    """The fused update vs oracle.adam_update (= torch's single-tensor CPU loop, k sub-steps for multiplicity k).
    max_mult: 4 = --train_feature's conv weights, 7 = an encoder Linear under --train_all, 10 = a conv weight under both."""
    lib = _lib.load()
    rng = np.random.default_rng(seed)
    p0 = rng.standard_normal((U, n)).astype(np.float32)
    mult = rng.integers(0, max_mult + 1, n).astype(np.uint8)
    Pd = torch.tensor(p0, device=DEV); Md = torch.zeros(U, n, device=DEV); Vd = torch.zeros(U, n, device=DEV)
    multd = torch.tensor(mult, device=DEV)
    h = Hyper(0.3, 2.5, 1, 1, {"AdamW": 0, "SGD": 1, "Adam": 2}[opt], lr, beta1, 0.999, 1e-8, wd, 0.0)
    p_ref = torch.tensor(p0.copy()); m_ref = torch.zeros(U, n); v_ref = torch.zeros(U, n)
    step_ref = np.zeros(n, dtype=np.int64)
    for s in range(steps):
        gnp = (rng.standard_normal((U, n)) * 10.0 ** rng.uniform(-6, 0, (U, n))).astype(np.float32)
        Gd = torch.tensor(gnp, device=DEV)
        check(lib.suta_op_adam(P(Pd), P(Gd), P(Md), P(Vd), P(multd), n, U, s, C.byref(h), None, stream()))
        gt = torch.tensor(gnp)
        for k in range(1, max_mult + 1):
            idx = torch.tensor(np.nonzero(mult == k)[0])
            if len(idx) == 0:
                continue
            pp, mm, vv = p_ref[:, idx].clone(), m_ref[:, idx].clone(), v_ref[:, idx].clone()
            if opt == "SGD":
                for _ in range(k):
                    pp -= lr * (gt[:, idx] + wd * pp)
            else:
                O.adam_update(pp, gt[:, idx], mm, vv, k * s, lr, k=k, beta1=beta1, weight_decay=wd, decoupled=(opt == "AdamW"))
            p_ref[:, idx], m_ref[:, idx], v_ref[:, idx] = pp, mm, vv
    torch.cuda.synchronize()
    d_ref = p_ref - torch.tensor(p0)
    d_k = Pd.cpu() - torch.tensor(p0)
    return dict(delta_rel=relerr(d_k, d_ref), frozen_moved=float((d_k[:, torch.tensor(mult == 0)]).abs().max()),
                m_rel=relerr(Md.cpu(), m_ref), v_rel=relerr(Vd.cpu(), v_ref))


def check_decode(Ts=(249, 37, 1, 700), seed=9):
    lib = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(seed)
    M, U = sum(Ts), len(Ts)
    logits = torch.randn(M, 32, device=DEV, generator=g)
    logits[:, 0] += 1.5
    rep = torch.rand(M, device=DEV, generator=g) < 0.4          # force repeats
    for i in range(1, M):
        pass
    lg = logits.cpu().numpy()
    r = rep.cpu().numpy()
    for i in range(1, M):
        if r[i]:
            lg[i] = lg[i - 1]
    logits = torch.tensor(lg, device=DEV)
    off = torch.tensor(np.concatenate([[0], np.cumsum(Ts)[:-1]]), dtype=torch.int64, device=DEV)
    Tt = torch.tensor(Ts, dtype=torch.int32, device=DEV)
    ids = torch.zeros(M, dtype=torch.int32, device=DEV); col = torch.zeros(M, dtype=torch.int32, device=DEV)
    ln = torch.zeros(U, dtype=torch.int32, device=DEV)
    check(lib.suta_op_decode(P(logits), P(off), P(Tt), U, 32, P(ids), P(col), P(ln), stream()))
    torch.cuda.synchronize()
    bad = int((ids.cpu().numpy() != lg.argmax(-1)).sum())
    o = 0
    mism = 0
    for u, T in enumerate(Ts):
        ref = O.ctc_collapse(lg[o:o + T].argmax(-1).tolist())
        got = col[o:o + int(ln[u])].cpu().tolist()
        mism += int(ref != got)
        o += T
    return dict(argmax_mismatch=bad, collapse_mismatch=mism)


ALL = [("gemm_plain", check_gemm_plain), ("gemm_epilogue", check_gemm_epilogue), ("gemm_shapes", check_gemm_shapes),
       ("gemm_mn", check_gemm_mn), ("gemm_mn2", lambda: check_gemm_mn(128, 64, 64, 14)),
       ("gemm_window", check_gemm_window), ("gemm_strided_conv", check_gemm_strided_conv),
       ("layernorm", check_layernorm), ("layernorm_bf16in_512", lambda: check_layernorm(512, (33, 70), 11, True)),
       ("layernorm_64", lambda: check_layernorm(64, (33, 70), 12, True)),
       ("attention", check_attention), ("loss", check_loss),
       ("loss_variants", lambda: {f"{e}{r}{n}": check_loss(em_coef=e, reweight=r, not_blank=n)["grad_rel"]
                                  for e in (0.3, 1.0, 0.0) for r in (False, True) for n in (False, True)}),
       ("loss_div", lambda: check_loss(div_coef=0.25)),
       ("loss_all_blank", lambda: check_loss(Ts=(40, 5, 1), blank_bias=50.0)),
       ("ctc", check_ctc), ("ctc_long", lambda: check_ctc(Ts=(1874, 2), seed=11, all_blank_utt=None)),
       ("adam", check_adam), ("adam_beta_l2", lambda: check_adam(opt="Adam", beta1=0.8, wd=0.01, lr=1e-3)),
       ("adamw_decay", lambda: check_adam(opt="AdamW", wd=0.01, lr=1e-3)), ("sgd_wd", lambda: check_adam(opt="SGD", lr=0.05, wd=0.01)),
       ("decode", check_decode)]
