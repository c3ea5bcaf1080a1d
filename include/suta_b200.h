/* suta_b200 -- C ABI of the B200-native SUTA adaptation engine (libsuta_b200.so).
 *
 * The reference (ishine/Test-time-adaptation-ASR-SUTA) has no FFI of its own: its boundary is the Python
 * function surface of main.py plus the torch / transformers calls underneath.  Each entry point below names
 * the reference interface it replaces (REF = reference repo, HF = transformers/models/wav2vec2).
 *
 * Conventions: plain C, plain pointers and sizes; every pointer marked DEV is a CUDA device pointer owned by
 * the caller; every call is asynchronous on `stream` (a cudaStream_t passed as void*), never allocates device
 * memory, and returns 0 on success or a SUTA_ERR_* code (text via suta_last_error()).
 * There is no CPU implementation behind this ABI: without a CUDA device every compute call fails.
 */
#ifndef SUTA_B200_H
#define SUTA_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SUTA_MAX_LAYERS 48
#define SUTA_MAX_CONV 8
#define SUTA_ABI_VERSION 4

typedef struct suta_engine suta_engine;

/* HF/configuration_wav2vec2.py:165-211 (the fields the hot path reads) */
typedef struct suta_model_cfg {
  int32_t hidden, layers, heads, intermediate, vocab;
  int32_t n_conv;
  int32_t conv_dim[SUTA_MAX_CONV], conv_kernel[SUTA_MAX_CONV], conv_stride[SUTA_MAX_CONV];
  int32_t pos_k, pos_groups;
  float ln_eps;
  /* the "lv60" family (wav2vec2-large-960h-lv60[-self], large-robust; REF/main_SDPL.py:238-241) */
  int32_t feat_norm_layer;               /* feat_extract_norm == "layer": every conv layer is Conv1d(+bias) -> LayerNorm(C)
                                            -> GELU (HF:275-299); its 2 x n_conv LayerNorm vectors are trainable */
  int32_t stable_layer_norm;             /* do_stable_layer_norm: pre-LN encoder layers + final encoder LayerNorm (HF:612-655,730-803) */
} suta_model_cfg;

/* Frozen weights, already in the engine's operand formats (see INTEGRATION.md for the packing recipe).
 * bf16 matrices are row-major [out, in] ("W") and their transposes [in, out] ("W_t", used by the backward). */
typedef struct suta_layer_weights {
  const void *wqkv, *wqkv_t;     /* DEV bf16 [3H,H] (q|k|v rows), [H,3H]     HF:500-507 */
  const void *wo, *wo_t;         /* DEV bf16 [H,H]                            HF:546     */
  const void *w1, *w1_t;         /* DEV bf16 [I,H], [H,I]                     HF:565     */
  const void *w2, *w2_t;         /* DEV bf16 [H,I], [I,H]                     HF:569     */
  const float *bqkv, *bo, *b1, *b2; /* DEV fp32 */
} suta_layer_weights;

typedef struct suta_weights {
  const float* conv0_w;                  /* DEV fp32 [C0,k0]                                  HF:319 */
  const float *gn_g, *gn_b;              /* DEV fp32 [C0] GroupNorm affine (NULL with feat_norm_layer) HF:320 */
  const float* conv_b[SUTA_MAX_CONV];    /* DEV fp32 [C_l] conv bias (feat_norm_layer only; NULL = none)  HF:286 */
  const void* conv_w[SUTA_MAX_CONV];     /* DEV bf16 [C_l, k_l*C_{l-1}], K order (tap,cin)    HF:269 */
  const void* conv_w_t[SUTA_MAX_CONV];   /* DEV bf16 [k_l*C_{l-1}, C_l] (train_feature dgrad)        */
  const void *proj_w, *proj_w_t;         /* DEV bf16 [H,C], [C,H]                             HF:431 */
  const float* proj_b;                   /* DEV fp32 [H] */
  const void *pos_w, *pos_w_t;           /* DEV bf16 [H, K*(H/G)] weight-normed, (tap,cin);   HF:360-368
                                            transposed-flipped [H(g,cin), K*(H/G)] (tap',cout) for dgrad */
  const float* pos_b;                    /* DEV fp32 [H] */
  suta_layer_weights layer[SUTA_MAX_LAYERS];
  const void *lm_w, *lm_w_t;             /* DEV bf16 [V,H], [H,V]                             HF:1708 */
  const float* lm_b;                     /* DEV fp32 [V] */
  const float* params0;                  /* DEV fp32 [n_params] pristine trainable vector (episodic snapshot) */
  const uint8_t* mult;                   /* DEV u8   [n_params] how many times REF/main.py:62-103 lists each element */
} suta_weights;

/* One segment of the per-utterance trainable vector. kind: 0 LN gamma, 1 LN beta, 2 GroupNorm gamma, 3 GroupNorm beta,
 * 4 conv weight (layer index in `index`), 5 linear weight [out,in], 6 linear / conv bias, 7 conv bias (feat_norm_layer +
 * TRAIN_FEATURE), 8 weight_norm magnitude g [K] and 9 weight_norm direction v [H, H/G, K] of the positional conv (TRAIN_ALL).
 * module: 0 feature_projection.layer_norm, 1 encoder.layer_norm, 2 layers[index].layer_norm,
 *         3 layers[index].final_layer_norm, 4 feature_extractor.conv_layers[index], 5 feature_projection.projection,
 *         6 feature_extractor.conv_layers[index].layer_norm (feat_norm_layer: kinds 0 / 1),
 *         TRAIN_ALL only: 7 / 8 / 9 / 10 layers[index].attention.{q,k,v,out}_proj, 11 layers[index].feed_forward.intermediate_dense,
 *         12 layers[index].feed_forward.output_dense, 13 lm_head, 14 encoder.pos_conv_embed.conv */
typedef struct suta_param_seg {
  int32_t kind, module, index;
  int64_t offset, size;
} suta_param_seg;

/* REF/main.py:172 forward_and_adapt's hyper-parameters + REF/main.py:8 setup_optimizer's */
typedef struct suta_hyper {
  float em_coef, temp;
  int32_t reweight, not_blank;
  int32_t opt_kind;                      /* 0 AdamW (decoupled weight decay), 1 SGD, 2 Adam (L2 weight decay) */
  float lr, beta1, beta2, eps, weight_decay;
  float div_coef;                        /* REF/main.py:201-203 div_loss weight (0: skipped) */
  float pl_coef;                         /* REF/main_SDPL.py:176 pseudo-label CTC weight (0: skipped; needs SUTA_FLAG_PSEUDO_LABEL) */
} suta_hyper;

const char* suta_last_error(void);
int suta_abi_version(void);
int suta_device_sm_count(void);

/* ---- engine lifecycle (replaces Wav2Vec2ForCTC.from_pretrained(...).eval().cuda() + configure_model +
 *      collect_params, REF/main.py:302-307; train_feature as REF/main.py:88-94) ---------------------------- */
#define SUTA_FLAG_TRAIN_FEATURE 1          /* REF/main.py:88-94: CNN front end + projection adapted per utterance */
#define SUTA_FLAG_PSEUDO_LABEL 2           /* REF/main_SDPL.py: reserve the CTC scratch (alpha lattice) in every batch workspace */
#define SUTA_FLAG_TRAIN_ALL 4              /* REF/main.py:96-100: EVERY parameter of the model is the utterance's own (implies
                                            * TRAIN_FEATURE's layout; one utterance per batch, as the reference adapts:
                                            * nothing is shared between utterances any more) */
int suta_engine_create(const suta_model_cfg* cfg, int flags, suta_engine** out);
void suta_engine_destroy(suta_engine* e);
int64_t suta_engine_param_count(const suta_engine* e);
int suta_engine_param_layout(const suta_engine* e, suta_param_seg* segs, int max_segs, int* n_segs);
int suta_engine_set_weights(suta_engine* e, const suta_weights* w);

/* ---- batch of independent utterances (the reference's per-utterance loop body, REF/main.py:319-402,
 *      executed for n_utts utterances at once, each with its own adapted parameters) ---------------------- */
int64_t suta_batch_workspace_bytes(suta_engine* e, int n_utts, const int32_t* n_samples);
int suta_batch_begin(suta_engine* e, int n_utts, const int32_t* n_samples, void* workspace /*DEV*/, int64_t bytes,
                     void* stream);
int suta_batch_info(const suta_engine* e, int64_t* total_frames, int32_t* frames /*[U]*/, int64_t* frame_off /*[U]*/,
                    int64_t* sample_off /*[U]*/, int64_t* total_samples);
/* waveform packed at sample_off. flags: bit0 = `wav` is host memory (pinned recommended), bit1 = already normalised
 * by the caller (HF processor output, REF/main.py:322) so the device normalisation is skipped */
int suta_batch_set_audio(suta_engine* e, const float* wav, int flags, void* stream);

/* REF/data.py:23 (wav += extra_noise * randn_like(wav)) on the device, on the raw waveform of the live batch, before the
 * normalisation; counter-based RNG keyed by (seed, utt_ids[u] -- host array, NULL = position in the batch) */
int suta_batch_add_noise(suta_engine* e, float sigma, uint64_t seed, const int32_t* utt_ids /*HOST, may be NULL*/, void* stream);

int suta_reset(suta_engine* e, void* stream);                       /* load_model_and_optimizer, REF/main.py:147-155 */
int suta_frontend(suta_engine* e, void* stream);                    /* HF/feature_extraction_wav2vec2.py:95 + HF:409-419 */
int suta_forward(suta_engine* e, void* stream);                     /* model(x).logits, REF/main.py:181/:214/:332 */
int suta_loss_backward(suta_engine* e, const suta_hyper* h, void* stream);   /* REF/main.py:183-205 */
int suta_optimizer_step(suta_engine* e, const suta_hyper* h, void* stream);  /* optimizer.step(), REF/main.py:206 */
int suta_decode(suta_engine* e, void* stream);                      /* argmax + batch_decode collapse, REF/main.py:333-334 */
/* one whole forward_and_adapt (REF/main.py:172-215): loss+backward on the current logits, update, forward again */
int suta_adapt_step(suta_engine* e, const suta_hyper* h, void* stream);

/* device views into the workspace (valid until the next suta_batch_begin) */
float* suta_logits(const suta_engine* e);        /* DEV fp32 [total_frames, V] */
float* suta_dlogits(const suta_engine* e);       /* DEV fp32 [total_frames, V] */
float* suta_losses(const suta_engine* e);        /* DEV fp32 [4][U]: total, entropy, mcc, pseudo-label ctc */
float* suta_params(const suta_engine* e);        /* DEV fp32 [U][n_params] */
float* suta_grads(const suta_engine* e);         /* DEV fp32 [U][n_params] */
int32_t* suta_argmax_ids(const suta_engine* e);  /* DEV i32 [total_frames] */
int32_t* suta_collapsed_ids(const suta_engine* e); /* DEV i32 [total_frames], utterance u at frame_off[u] */
int32_t* suta_collapsed_len(const suta_engine* e); /* DEV i32 [U] */
/* optimizer state of the live batch (torch.optim state_dict()['state'], REF/main.py:140,150): both Adam moments and the
 * number of optimizer.step() calls since the last reset.  A caller that carries a NON-episodic ("continual", the
 * reference's default without --episodic, REF/main.py:319-348) model from one utterance to the next copies P, both
 * moments and the step count across suta_batch_begin with these. */
/* call after writing suta_params() directly (model.load_state_dict, REF/main.py:149): refreshes the bf16 GEMM-operand
 * copies of the per-utterance matrices and invalidates the cached CNN output */
int suta_params_written(suta_engine* e, void* stream);
float* suta_adam_exp_avg(const suta_engine* e);     /* DEV fp32 [U][n_params] */
float* suta_adam_exp_avg_sq(const suta_engine* e);  /* DEV fp32 [U][n_params] */
int suta_opt_steps(const suta_engine* e);
int suta_set_opt_steps(suta_engine* e, int steps);
const void* suta_debug_buffer(const suta_engine* e, const char* name, int64_t* rows, int64_t* cols, int* dtype);
int64_t suta_launch_count(const suta_engine* e); /* kernels launched by this engine since creation */
/* Small batches (<= 4096 frames: the reference's one-utterance-at-a-time operating point, REF/main.py:319-402) are
 * launch-bound; suta_forward / suta_loss_backward record their launch chain as a CUDA graph the second time it runs in a
 * batch and replay it afterwards (SUTA_NO_GRAPH=1: never; SUTA_GRAPH_MAX_TOKENS=n moves the bound).  Number of replays so far: */
int64_t suta_graph_replays(const suta_engine* e);
/* per-launch CUDA-event timing of the tcgen05 GEMM (bench.py roofline leg). Reads and clears the counters collected
 * so far (any pointer may be NULL), then switches collection on/off. */
int suta_profile(suta_engine* e, int enable, double* gemm_ms, int64_t* gemm_launches, double* gemm_flops);
/* per-kernel-class breakdown of the last read-out: lines "tag<TAB>ms<TAB>flops<TAB>launches" */
const char* suta_profile_report(const suta_engine* e);

/* debug hook for kernel tuning: CTA 0 of every following tcgen05 GEMM launch writes a per-tile clock64 timeline into
 * dev_buf[cap][8] (device memory; see csrc/gemm_tc.cu); NULL switches it off */
void suta_debug_set_gemm_trace(long long* dev_buf, int cap);

/* ---- single operators (the kernels behind the calls above, exposed for parity tests) ------------------- */
/* D[M,N] = A[M,K] B[N,K]^T (+bias)(act & 3 == 1: GELU, saving GELU' to aux_out; == 2: times aux_in)(+residual), tcgen05
 * GEMM; out_f32 and/or out_bf16; act & 4: out_f32 += D instead of = D (in-place accumulation, no residual pointer) */
int suta_op_gemm(const void* a, int64_t a_rows, int64_t a_row_stride, const void* b, int64_t b_rows, int64_t b_row_stride,
                 int M, int N, int K, float* out_f32, void* out_bf16, int out_ld, const float* bias,
                 const float* residual, int res_ld, int act, const void* aux_in, void* aux_out, int aux_ld, void* stream);
/* same contraction with per-operand memory order: a_mn / b_mn = 1 -> operand stored [K rows][M|N contiguous]
 * (weight-gradient GEMMs reduce over time, so both operands are naturally in that order); fp32 output */
int suta_op_gemm_mn(const void* a, int64_t a_rows, int64_t a_row_stride, int a_mn, const void* b, int64_t b_rows,
                    int64_t b_row_stride, int b_mn, int M, int N, int K, float* out_f32, int out_ld, void* stream);
int suta_op_layernorm_fwd(const float* x_f32, const void* x_bf16, const int32_t* row_utt, const float* P, int64_t pstride,
                          int g_off, int b_off, float* y_f32, void* y_bf16, float* mean, float* rstd, int64_t M, int N,
                          float eps, void* stream);
/* dgamma/dbeta are written (not accumulated) into G by a fixed-order two-stage reduction: bit-reproducible.
 * tok_off / T describe the utterances' packed rows; scratch: suta_op_layernorm_bwd_scratch_floats(N, n_utts) floats */
int suta_op_layernorm_bwd(const float* dy, const float* x_f32, const void* x_bf16, const float* mean, const float* rstd,
                          const int32_t* row_utt, const float* P, int64_t pstride, int g_off, int b_off, float* G,
                          float* dx_f32, void* dx_bf16, int64_t M, int N, const int64_t* tok_off, const int32_t* T,
                          int n_utts, float* scratch, void* stream);
/* The LayerNorm variants of the lv60 family.  mode 1: y = GELU(LN(x)) with a bf16 x -- a conv layer of the LayerNorm
 * feature extractor (HF:291-299); rows with row_utt < 0 (gaps of the conv layouts) are skipped.  mode 2: y_f32 = x +
 * y32_bias while y_bf16 = LN(x) -- the pre-LN encoder keeps the residual stream beside the normalised branch (HF:638-645).
 * mode 0: suta_op_layernorm_fwd (+ y32_bias on the fp32 output). */
int suta_op_layernorm_fwd_mode(const float* x_f32, const void* x_bf16, const int32_t* row_utt, const float* P, int64_t pstride,
                               int g_off, int b_off, float* y_f32, void* y_bf16, float* mean, float* rstd, int64_t M, int N,
                               float eps, const float* y32_bias, int mode, void* stream);
/* Their backward.  mode 1: dy (fp32 or bf16, exactly one non-null) is the gradient of the GELU output, x is bf16; dx_bf16
 * may alias dy_bf16.  mode 2: fp32 x and dy, dx += dx_add (may alias dx_f32). */
int suta_op_layernorm_bwd_mode(const float* dy_f32, const void* dy_bf16, const float* x_f32, const void* x_bf16, const float* mean,
                               const float* rstd, const int32_t* row_utt, const float* P, int64_t pstride, int g_off, int b_off,
                               float* G, float* dx_f32, void* dx_bf16, const float* dx_add, int mode, int64_t M, int N,
                               const int64_t* tok_off, const int32_t* T, int n_utts, float* scratch, void* stream);
int64_t suta_op_layernorm_bwd_scratch_floats(int N, int n_utts);
int suta_op_attention_fwd(const void* qkv, void* O, float* lse, const int32_t* blk_tab, int n_blk, int H, int heads,
                          int64_t M, void* stream);
int suta_op_attention_bwd(const void* qkv, const void* O, const void* dO, const float* lse, float* D, void* dqkv,
                          const int32_t* blk_tab, int n_blk, int H, int heads, int64_t M, void* stream);
int suta_op_loss(const float* logits, const int64_t* tok_off, const int32_t* T, int n_utts, float em_coef, float temp,
                 int reweight, int not_blank, float div_coef, float* loss /*[3U]*/, float* dlogits_f32, void* dlogits_bf16,
                 void* stream);
/* pseudo-label CTC loss (REF/main_SDPL.py:194-209) of every utterance against its own greedy transcript.
 * collapsed / collapsed_len: output of suta_op_decode; alpha: sum_u T_u (2 T_u + 1) floats of scratch, g: [M,32] scratch;
 * loss [U] and dlogits_f32 [M,32] are outputs (value of the CTC term alone, its gradient w.r.t. the logits) */
int suta_op_ctc_pseudo_label(const float* logits, const int64_t* tok_off, const int32_t* T, int n_utts, const int32_t* collapsed,
                             const int32_t* collapsed_len, float* alpha, float* g, float* loss, float* dlogits_f32,
                             int32_t* target_len, void* stream);
/* out[row] = entropy(softmax(logits[row]/temp)), V = 32: softmax_entropy, REF/main.py:26-28 */
int suta_op_softmax_entropy(const float* logits, int64_t rows, float temp, float* out, void* stream);
int suta_op_adam(float* P, const float* G, float* Mom, float* Var, const uint8_t* mult, int64_t n, int n_utts,
                 int step_index, const suta_hyper* h, void* shadow_bf16, void* stream);
int suta_op_decode(const float* logits, const int64_t* tok_off, const int32_t* T, int n_utts, int V, int32_t* ids,
                   int32_t* collapsed, int32_t* out_len, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SUTA_B200_H */
