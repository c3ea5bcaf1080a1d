"""Python host side of the B200 engine: weight packing, workspace ownership, and the batched adaptation calls.

PyTorch is used here for device memory (the workspace and the packed weights are torch tensors), streams and the
one-off layout transforms of the frozen weights.  All arithmetic of the hot path happens in libsuta_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import Hyper, ModelCfg, ParamSeg, Weights, check
from .config import ModelConfig

CHECKPOINT_STEPS = (1, 3, 5, 10, 20, 40)       # REF/main.py:350-398

_MODULE_NAMES = {0: "wav2vec2.feature_projection.layer_norm", 1: "wav2vec2.encoder.layer_norm",
                 2: "wav2vec2.encoder.layers.{i}.layer_norm", 3: "wav2vec2.encoder.layers.{i}.final_layer_norm",
                 4: "wav2vec2.feature_extractor.conv_layers.{i}", 5: "wav2vec2.feature_projection.projection",
                 6: "wav2vec2.feature_extractor.conv_layers.{i}.layer_norm",
                 # SUTA_FLAG_TRAIN_ALL (REF/main.py:96-100)
                 7: "wav2vec2.encoder.layers.{i}.attention.q_proj", 8: "wav2vec2.encoder.layers.{i}.attention.k_proj",
                 9: "wav2vec2.encoder.layers.{i}.attention.v_proj", 10: "wav2vec2.encoder.layers.{i}.attention.out_proj",
                 11: "wav2vec2.encoder.layers.{i}.feed_forward.intermediate_dense",
                 12: "wav2vec2.encoder.layers.{i}.feed_forward.output_dense", 13: "lm_head",
                 14: "wav2vec2.encoder.pos_conv_embed.conv"}
_KIND_LEAF = {0: "weight", 1: "bias", 2: "layer_norm.weight", 3: "layer_norm.bias", 4: "conv.weight", 5: "weight", 6: "bias",
              7: "conv.bias", 8: "parametrizations.weight.original0", 9: "parametrizations.weight.original1"}
MAX_MULTIPLICITY = 12                           # csrc/optim.cu MAXK


@dataclass
class AdaptHyper:
    """REF/main.py:172 (forward_and_adapt) + :8 (setup_optimizer); defaults = REF/scripts/LS.sh preset."""
    em_coef: float = 0.3
    temp: float = 2.5
    reweight: bool = True
    not_blank: bool = True
    opt: str = "AdamW"
    lr: float = 2e-5
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    weight_decay: float = 0.0
    div_coef: float = 0.0          # REF/main.py:201-203 (no script sets it)
    pl_coef: float = 0.0           # REF/main_SDPL.py:176 pseudo-label CTC weight (the SDPL baseline runs with 1.0)

    _OPT_KIND = {"AdamW": 0, "SGD": 1, "Adam": 2}     # Adam = L2 weight decay (torch.optim.Adam), AdamW = decoupled

    def to_c(self) -> Hyper:
        if self.opt not in self._OPT_KIND:
            raise ValueError(f"unsupported optimizer {self.opt!r} (AdamW, Adam, SGD)")
        return Hyper(self.em_coef, self.temp, int(self.reweight), int(self.not_blank), self._OPT_KIND[self.opt],
                     self.lr, self.beta1, self.beta2, self.eps, self.weight_decay, self.div_coef, self.pl_coef)


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


class SutaEngine:
    """One frozen wav2vec2-CTC model on one GPU + the batched SUTA loop over independent utterances."""

    def __init__(self, cfg, state_dict: Dict[str, torch.Tensor], train_feature: bool = False,
                 trainable_mult: Optional[Dict[str, int]] = None, device: Optional[torch.device] = None,
                 pseudo_label: bool = False, train_all: bool = False):
        if not torch.cuda.is_available():
            raise _lib.SutaError("suta_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg = ModelConfig.from_any(cfg)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.train_all = bool(train_all)         # REF/main.py:96-100: every parameter is the utterance's own; one utterance per batch
        self.train_feature = bool(train_feature) or self.train_all      # (its layout contains train_feature's)
        c = self.cfg
        if self.train_all and (c.feat_extract_norm == "layer") != bool(c.do_stable_layer_norm):
            raise NotImplementedError("--train_all: GroupNorm + post-LN or LayerNorm feature extractor + pre-LN (the two shipped families)")
        cc = ModelCfg()
        cc.hidden, cc.layers, cc.heads, cc.intermediate, cc.vocab = (c.hidden_size, c.num_hidden_layers,
                                                                      c.num_attention_heads, c.intermediate_size, c.vocab_size)
        cc.n_conv = len(c.conv_dim)
        for i in range(cc.n_conv):
            cc.conv_dim[i], cc.conv_kernel[i], cc.conv_stride[i] = c.conv_dim[i], c.conv_kernel[i], c.conv_stride[i]
        cc.pos_k, cc.pos_groups, cc.ln_eps = c.num_conv_pos_embeddings, c.num_conv_pos_embedding_groups, c.layer_norm_eps
        cc.feat_norm_layer, cc.stable_layer_norm = int(c.feat_extract_norm == "layer"), int(c.do_stable_layer_norm)
        if self.train_feature and c.feat_extract_norm == "layer" and not c.conv_bias:
            raise NotImplementedError("--train_feature on a LayerNorm feature extractor without conv bias (no checkpoint has one)")
        h = C.c_void_p()
        self.pseudo_label = bool(pseudo_label)         # SDPL: reserves the CTC lattice scratch in every batch workspace
        flags = int(self.train_feature) | (2 if self.pseudo_label else 0) | (4 if self.train_all else 0)
        check(self.lib.suta_engine_create(C.byref(cc), flags, C.byref(h)))
        self._h = h
        self.n_params = int(self.lib.suta_engine_param_count(h))
        n = C.c_int()
        check(self.lib.suta_engine_param_layout(h, None, 0, C.byref(n)))
        segs = (ParamSeg * n.value)()
        check(self.lib.suta_engine_param_layout(h, segs, n.value, C.byref(n)))
        self.segments = []      # (hf name, offset, size)
        self.seg_kind = {}      # hf name -> kind (4 = conv weight stored as [Cout, tap, Cin])
        for s in segs:
            name = _MODULE_NAMES[s.module].format(i=s.index) + "." + _KIND_LEAF[s.kind]
            self.segments.append((name, int(s.offset), int(s.size)))
            self.seg_kind[name] = int(s.kind)
        self._keep: List[torch.Tensor] = []
        self._pack_weights(state_dict, trainable_mult)
        self._ws: Optional[torch.Tensor] = None
        self.n_utts = 0

    # ------------------------------------------------------------------ weights
    def _dev(self, t: torch.Tensor, dtype) -> torch.Tensor:
        t = t.detach().to(device=self.device, dtype=dtype).contiguous()
        self._keep.append(t)
        return t

    def _pack_weights(self, sd: Dict[str, torch.Tensor], trainable_mult: Optional[Dict[str, int]]):
        c = self.cfg
        bf, f32 = torch.bfloat16, torch.float32
        w = Weights()
        p = lambda t: C.c_void_p(t.data_ptr())
        g = lambda k: sd[k].detach().to(torch.float32)

        def mat(t):       # [out,in] fp32 -> (bf16 W, bf16 W^T)
            t = t.to(self.device)
            return self._dev(t, bf), self._dev(t.t(), bf)

        fe = "wav2vec2.feature_extractor.conv_layers."
        w.conv0_w = p(self._dev(g(fe + "0.conv.weight").reshape(c.conv_dim[0], c.conv_kernel[0]), f32))
        if c.feat_extract_norm == "layer":     # conv LayerNorms are trainable segments; the conv biases are frozen
            for l in range(len(c.conv_dim)):
                if c.conv_bias:
                    w.conv_b[l] = p(self._dev(g(fe + f"{l}.conv.bias"), f32))
        else:
            w.gn_g = p(self._dev(g(fe + "0.layer_norm.weight"), f32))
            w.gn_b = p(self._dev(g(fe + "0.layer_norm.bias"), f32))
        for l in range(1, len(c.conv_dim)):
            cw = g(fe + f"{l}.conv.weight")                       # [Cout, Cin, k] -> [Cout, (k, Cin)]
            w.conv_w[l] = p(self._dev(cw.permute(0, 2, 1).reshape(cw.shape[0], -1), bf))
        pw, pwt = mat(g("wav2vec2.feature_projection.projection.weight"))
        w.proj_w, w.proj_w_t = p(pw), p(pwt)
        w.proj_b = p(self._dev(g("wav2vec2.feature_projection.projection.bias"), f32))
        # positional conv: fold weight_norm once (frozen), HF/modeling_wav2vec2.py:344-352
        pre = "wav2vec2.encoder.pos_conv_embed.conv."
        if pre + "weight" in sd:
            pcw = g(pre + "weight")
        else:
            wg, wv = g(pre + "parametrizations.weight.original0"), g(pre + "parametrizations.weight.original1")
            pcw = wv * (wg / wv.norm(dim=(0, 1), keepdim=True))
        H, G, K = c.hidden_size, c.num_conv_pos_embedding_groups, c.num_conv_pos_embeddings
        CG = H // G
        w.pos_w = p(self._dev(pcw.permute(0, 2, 1).reshape(H, K * CG), bf))                 # [co, (tap, ci)]
        wf = pcw.view(G, CG, CG, K).flip(-1).permute(0, 2, 3, 1)                              # [g, ci, tap', co]
        w.pos_w_t = p(self._dev(wf.reshape(H, K * CG), bf))
        w.pos_b = p(self._dev(g(pre + "bias"), f32))
        for l in range(c.num_hidden_layers):
            b = f"wav2vec2.encoder.layers.{l}."
            lw = w.layer[l]
            qkv = torch.cat([g(b + f"attention.{n}_proj.weight") for n in ("q", "k", "v")], 0)
            a, at = mat(qkv); lw.wqkv, lw.wqkv_t = p(a), p(at)
            a, at = mat(g(b + "attention.out_proj.weight")); lw.wo, lw.wo_t = p(a), p(at)
            a, at = mat(g(b + "feed_forward.intermediate_dense.weight")); lw.w1, lw.w1_t = p(a), p(at)
            a, at = mat(g(b + "feed_forward.output_dense.weight")); lw.w2, lw.w2_t = p(a), p(at)
            lw.bqkv = p(self._dev(torch.cat([g(b + f"attention.{n}_proj.bias") for n in ("q", "k", "v")], 0), f32))
            lw.bo = p(self._dev(g(b + "attention.out_proj.bias"), f32))
            lw.b1 = p(self._dev(g(b + "feed_forward.intermediate_dense.bias"), f32))
            lw.b2 = p(self._dev(g(b + "feed_forward.output_dense.bias"), f32))
        a, at = mat(g("lm_head.weight")); w.lm_w, w.lm_w_t = p(a), p(at)
        w.lm_b = p(self._dev(g("lm_head.bias"), f32))
        # pristine trainable vector + per-element multiplicity (REF/main.py:62-103 lists some tensors several times)
        if self.train_all and pre + "parametrizations.weight.original0" not in sd:
            # a folded positional-conv weight: weight_norm's own initialisation (g = ||w|| per tap, v = w) reproduces it
            sd = dict(sd)
            sd[pre + "parametrizations.weight.original0"] = pcw.norm(dim=(0, 1), keepdim=True)
            sd[pre + "parametrizations.weight.original1"] = pcw.clone()
        p0 = torch.empty(self.n_params, dtype=f32)
        mult = torch.zeros(self.n_params, dtype=torch.uint8)
        self._hf_shapes = {name: tuple(sd[name].shape) for name, _o, _s in self.segments}
        for name, off, size in self.segments:
            p0[off:off + size] = self.to_engine_layout(name, g(name))
            mult[off:off + size] = 1 if trainable_mult is None else int(trainable_mult.get(name, 0))
        self.params0 = self._dev(p0, f32)
        self.mult = self._dev(mult, torch.uint8)
        w.params0, w.mult = p(self.params0), p(self.mult)
        self._weights = w
        check(self.lib.suta_engine_set_weights(self._h, C.byref(w)))

    def to_engine_layout(self, name: str, t: torch.Tensor) -> torch.Tensor:
        """HF tensor -> flat segment of the trainable vector (conv weights go [Cout, Cin, k] -> [Cout, k, Cin])."""
        if self.seg_kind[name] == 4:
            t = t.permute(0, 2, 1)
        return t.reshape(-1)

    def from_engine_layout(self, name: str, flat: torch.Tensor) -> torch.Tensor:
        """Flat segment -> tensor of the HF parameter's shape."""
        shp = self._hf_shapes[name]
        if self.seg_kind[name] == 4:
            return flat.view(shp[0], shp[2], shp[1]).permute(0, 2, 1)
        return flat.view(*shp)

    def set_trainable(self, mult_by_name: Dict[str, int]):
        """Re-select what the optimizer updates (collect_params' result): name -> multiplicity (0 = frozen)."""
        if max(list(mult_by_name.values()) + [0]) > MAX_MULTIPLICITY:
            raise ValueError(f"a parameter listed more than {MAX_MULTIPLICITY} times (csrc/optim.cu MAXK)")
        m = torch.zeros(self.n_params, dtype=torch.uint8)
        for name, off, size in self.segments:
            m[off:off + size] = int(mult_by_name.get(name, 0))
        self.mult.copy_(m.to(self.device))

    # ------------------------------------------------------------------ batches
    def begin_batch(self, wavs: Sequence[np.ndarray], max_samples: Optional[int] = 600000):
        """Start adapting a batch of raw (un-normalised) fp32 waveforms; copies them to the device.
        max_samples: REF/data.py:9,19-21 -- a waveform of >= 600000 samples keeps its first 600000 (the clamp is part of
        the batch layout: samples beyond it never cross PCIe)."""
        lens = np.asarray([len(w) for w in wavs], dtype=np.int32)
        if max_samples is not None:
            lens = np.minimum(lens, max_samples).astype(np.int32)
        self.begin_batch_lengths(lens)
        host = torch.zeros(self.total_samples, dtype=torch.float32).pin_memory()
        hv = host.numpy()
        for w, n, o in zip(wavs, lens, self.sample_off):
            hv[o:o + n] = np.asarray(w[:n], dtype=np.float32)
        self.set_audio(host)

    def begin_batch_lengths(self, lens: np.ndarray):
        lens = np.ascontiguousarray(lens, dtype=np.int32)
        U = len(lens)
        lp = lens.ctypes.data_as(C.POINTER(C.c_int32))
        need = int(self.lib.suta_batch_workspace_bytes(self._h, U, lp))
        if need < 0:
            check(2)
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        check(self.lib.suta_batch_begin(self._h, U, lp, C.c_void_p(self._ws.data_ptr()), self._ws.numel(), _stream_ptr()))
        M, S = C.c_int64(), C.c_int64()
        frames = (C.c_int32 * U)()
        foff = (C.c_int64 * U)()
        soff = (C.c_int64 * U)()
        check(self.lib.suta_batch_info(self._h, C.byref(M), frames, foff, soff, C.byref(S)))
        self.n_utts, self.total_frames, self.total_samples = U, M.value, S.value
        self.frames = np.asarray(frames[:], dtype=np.int64)
        self.frame_off = np.asarray(foff[:], dtype=np.int64)
        self.sample_off = np.asarray(soff[:], dtype=np.int64)
        self.lengths = lens.copy()

    def set_audio(self, packed: torch.Tensor, normalized: bool = False):
        """packed: fp32 [total_samples] (pinned host or device), utterance u at sample_off[u].
        normalized=True: the caller already applied the HF processor's zero-mean/unit-variance (REF/main.py:322)."""
        assert packed.dtype == torch.float32 and packed.numel() >= self.total_samples
        self._audio_keep = packed
        flags = int(not packed.is_cuda) | (2 if normalized else 0)
        check(self.lib.suta_batch_set_audio(self._h, C.c_void_p(packed.data_ptr()), flags, _stream_ptr()))

    def add_noise(self, sigma: float, seed: int = 0, utt_ids: Optional[Sequence[int]] = None):
        """REF/data.py:23 on the device: wav += sigma * N(0,1) on the raw waveform set by set_audio (before normalisation).
        utt_ids: stable utterance ids, so the noise of an utterance does not depend on the batch it is adapted in."""
        ids = None
        if utt_ids is not None:
            self._noise_ids = np.ascontiguousarray(utt_ids, dtype=np.int32)
            assert len(self._noise_ids) == self.n_utts
            ids = self._noise_ids.ctypes.data_as(C.POINTER(C.c_int32))
        check(self.lib.suta_batch_add_noise(self._h, float(sigma), int(seed) & (2 ** 64 - 1), ids, _stream_ptr()))

    def profile(self, enable: bool):
        """Read-and-clear the per-launch GEMM timers, then switch collection on/off.
        Returns (gemm_ms, gemm_launches, gemm_flops) accumulated since the previous call."""
        ms, n, fl = C.c_double(), C.c_int64(), C.c_double()
        check(self.lib.suta_profile(self._h, int(enable), C.byref(ms), C.byref(n), C.byref(fl)))
        return ms.value, n.value, fl.value

    def profile_report(self):
        """Per-kernel-class breakdown of the last profile() read-out: {tag: (ms, flops, launches, algorithmic bytes)}."""
        out = {}
        for line in self.lib.suta_profile_report(self._h).decode().splitlines():
            tag, ms, fl, n, by = line.split("\t")
            out[tag] = (float(ms), float(fl), int(n), float(by))
        return out

    def _view(self, ptr: int, shape, dtype) -> torch.Tensor:
        off = ptr - self._ws.data_ptr()
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        return self._ws[off:off + n].view(dtype).view(*shape)

    # ------------------------------------------------------------------ the path
    def reset(self):
        check(self.lib.suta_reset(self._h, _stream_ptr()))

    def frontend(self):
        check(self.lib.suta_frontend(self._h, _stream_ptr()))

    def forward(self) -> torch.Tensor:
        check(self.lib.suta_forward(self._h, _stream_ptr()))
        return self.logits()

    def loss_backward(self, hp: AdaptHyper):
        h = hp.to_c()
        check(self.lib.suta_loss_backward(self._h, C.byref(h), _stream_ptr()))

    def optimizer_step(self, hp: AdaptHyper):
        h = hp.to_c()
        check(self.lib.suta_optimizer_step(self._h, C.byref(h), _stream_ptr()))

    def adapt_step(self, hp: AdaptHyper) -> torch.Tensor:
        h = hp.to_c()
        check(self.lib.suta_adapt_step(self._h, C.byref(h), _stream_ptr()))
        return self.logits()

    def decode_ids(self) -> List[List[int]]:
        """Greedy CTC on the device; returns the collapsed id sequence per utterance (one D2H copy)."""
        check(self.lib.suta_decode(self._h, _stream_ptr()))
        ids = self._view(self.lib.suta_collapsed_ids(self._h), (self.total_frames,), torch.int32)
        lens = self._view(self.lib.suta_collapsed_len(self._h), (self.n_utts,), torch.int32)
        ids_h, lens_h = ids.cpu().numpy(), lens.cpu().numpy()
        return [ids_h[o:o + n].tolist() for o, n in zip(self.frame_off, lens_h)]

    # ------------------------------------------------------------------ views
    def logits(self) -> torch.Tensor:
        return self._view(self.lib.suta_logits(self._h), (self.total_frames, self.cfg.vocab_size), torch.float32)

    def dlogits(self) -> torch.Tensor:
        return self._view(self.lib.suta_dlogits(self._h), (self.total_frames, self.cfg.vocab_size), torch.float32)

    def losses(self) -> torch.Tensor:
        """[4, U]: total, entropy term, MCC term, pseudo-label CTC term (REF/main.py:188-199, REF/main_SDPL.py:176)."""
        return self._view(self.lib.suta_losses(self._h), (4, self.n_utts), torch.float32)

    def params(self) -> torch.Tensor:
        return self._view(self.lib.suta_params(self._h), (self.n_utts, self.n_params), torch.float32)

    def grads(self) -> torch.Tensor:
        return self._view(self.lib.suta_grads(self._h), (self.n_utts, self.n_params), torch.float32)

    def params_written(self):
        """Tell the engine that params() was written directly (bf16 operand copies / cached CNN output are stale)."""
        check(self.lib.suta_params_written(self._h, _stream_ptr()))

    def exp_avg(self) -> torch.Tensor:
        """Adam first moment of every utterance of the live batch (torch.optim state 'exp_avg')."""
        return self._view(self.lib.suta_adam_exp_avg(self._h), (self.n_utts, self.n_params), torch.float32)

    def exp_avg_sq(self) -> torch.Tensor:
        return self._view(self.lib.suta_adam_exp_avg_sq(self._h), (self.n_utts, self.n_params), torch.float32)

    @property
    def opt_steps(self) -> int:
        """optimizer.step() calls since the last reset (bias correction uses multiplicity * opt_steps + j)."""
        return int(self.lib.suta_opt_steps(self._h))

    @opt_steps.setter
    def opt_steps(self, n: int):
        check(self.lib.suta_set_opt_steps(self._h, int(n)))

    def argmax_ids(self) -> torch.Tensor:
        return self._view(self.lib.suta_argmax_ids(self._h), (self.total_frames,), torch.int32)

    def utt_logits(self, u: int) -> torch.Tensor:
        o, t = int(self.frame_off[u]), int(self.frames[u])
        return self.logits()[o:o + t]

    def utt_params(self, u: int) -> Dict[str, torch.Tensor]:
        """Adapted trainables of utterance u, keyed by HF parameter name, in the HF tensor shapes."""
        P = self.params()[u]
        return {name: self.from_engine_layout(name, P[off:off + size]) for name, off, size in self.segments}

    def debug_buffer(self, name: str) -> torch.Tensor:
        r, c, dt = C.c_int64(), C.c_int64(), C.c_int()
        ptr = self.lib.suta_debug_buffer(self._h, name.encode(), C.byref(r), C.byref(c), C.byref(dt))
        if not ptr:
            raise KeyError(name)
        return self._view(ptr, (r.value, c.value), torch.bfloat16 if dt.value == 1 else torch.float32)

    @property
    def launch_count(self) -> int:
        return int(self.lib.suta_launch_count(self._h))

    @property
    def graph_replays(self) -> int:
        """Forward / backward launch chains replayed as CUDA graphs so far (small, launch-bound batches only)."""
        return int(self.lib.suta_graph_replays(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self.lib.suta_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
