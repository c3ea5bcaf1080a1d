// Host-callable launchers of the non-GEMM kernels (all asynchronous on `stream`, return SUTA_* status).
#pragma once
#include "common.cuh"

// Per-utterance trainable vector: parameter `off` of utterance u lives at P + u*stride + off.
struct UttParams {
  const float* P = nullptr;
  long long stride = 0;
};

// ---- attention_fwd2_tc.cu / attention_bwd_tc.cu (tcgen05 / TMEM) -----------------------------------
// qkv bf16 [M,3H], O bf16 [M,H], LSE fp32 [heads,M] (base 2), block table: one int4 {utt_row0, T_u, m0, 0} per 128 rows
int attention_forward(const bf16* qkv, bf16* O, float* LSE, const int4* blk_tab, int n_blk, int H, int heads, long long M,
                      cudaStream_t stream);
int attention_backward(const bf16* qkv, const bf16* O, const bf16* dO, const float* LSE, float* D, bf16* dqkv,
                       const int4* blk_tab, int n_blk, int H, int heads, long long M, cudaStream_t stream);

// ---- norm.cu --------------------------------------------------------------------------------
// y = LayerNorm(x) * gamma[u] + beta[u]; x is fp32 (x_f32) or bf16 (x_bf16), exactly one non-null.
// y32_bias (optional, fp32 [N]) is added to the fp32 output only: that copy seeds the residual sum the next GEMM accumulates
// into, so the GEMM's bias is folded in here.
// mode: LN_PLAIN; LN_GELU (bf16 input): the outputs are GELU(LN(x)) -- conv layers of the LayerNorm feature extractor,
// HF/modeling_wav2vec2.py:291-299; LN_KEEP_INPUT (fp32 input): y_f32 = x (+ y32_bias) instead of LN(x) -- the pre-LN encoder
// keeps the residual stream beside the normalised branch, HF:638-645.  LN_GELU skips rows with row_utt < 0 (conv-layout gaps).
enum { LN_PLAIN = 0, LN_GELU = 1, LN_KEEP_INPUT = 2 };
int layernorm_forward(const float* x_f32, const bf16* x_bf16, const int* row_utt, UttParams prm, int g_off, int b_off,
                      float* y_f32, bf16* y_bf16, float* mean, float* rstd, long long M, int N, float eps,
                      cudaStream_t stream, const float* y32_bias = nullptr, int mode = LN_PLAIN);
// dgamma/dbeta are WRITTEN into G (same layout as the parameter vector) by a fixed-order two-stage reduction through
// `scratch` (layernorm_backward_scratch_floats(N, n_utts) floats) -- bit-reproducible; G and the dx outputs are optional.
// tok_off / T: first packed row and row count of every utterance (the rows row_utt describes).
// defer != nullptr: the second stage is NOT launched; *defer describes it and the caller runs layernorm_backward_reduce
// once for all the LayerNorms of a backward pass (each needs its own scratch until then).
struct LnReduceItem {
  const float* part;
  int g_off, b_off, N, rows_per_cta;
  const long long* tok_off;    // optional: first row / row count of every utterance in THIS LayerNorm's row space
  const int* T;                //   (conv-layer layouts); null = the arguments of layernorm_backward_reduce
};
constexpr int LN_REDUCE_MAX = 2 * 48 + 4 + 8;
struct LnReduceBatch {
  LnReduceItem item[LN_REDUCE_MAX];
  int n = 0;
};
int layernorm_backward(const float* dy, const float* x_f32, const bf16* x_bf16, const float* mean, const float* rstd,
                       const int* row_utt, UttParams prm, int g_off, int b_off, float* G, float* dx_f32, bf16* dx_bf16,
                       long long M, int N, const long long* tok_off, const int* T, int n_utts, float* scratch,
                       cudaStream_t stream, LnReduceItem* defer = nullptr, const float* dx_add = nullptr);
// dx_add (fp32 [M,N], fp32 x only; may alias dx_f32): added to dx -- the residual path of the pre-LN encoder.
// Backward of GELU(LayerNorm(x)) (forward mode LN_GELU): dy is the gradient of the GELU output, fp32 or bf16 (exactly one
// non-null); dx_bf16 may alias dy_bf16.
int layernorm_gelu_backward(const float* dy_f32, const bf16* dy_bf16, const bf16* x_bf16, const float* mean, const float* rstd,
                            const int* row_utt, UttParams prm, int g_off, int b_off, float* G, float* dx_f32, bf16* dx_bf16,
                            long long M, int N, const long long* tok_off, const int* T, int n_utts, float* scratch,
                            cudaStream_t stream, LnReduceItem* defer = nullptr);
int layernorm_backward_reduce(const LnReduceBatch& b, const long long* tok_off, const int* T, int n_utts, float* G,
                              long long pstride, cudaStream_t stream);
long long layernorm_backward_scratch_floats(int N, int n_utts);

// ---- frontend.cu ----------------------------------------------------------------------------
// per-utterance (x - mean) / sqrt(var + 1e-7), optional additive noise is applied by the caller beforehand
int normalize_audio(const float* wav, float* out, const long long* samp_off, const int* n_samples, int n_utts,
                    int max_samples, double* stats_scratch /*[2*n_utts]*/, cudaStream_t stream);
// wav[u][i] += sigma * N(0,1), REF/data.py:23; Philox keyed by (seed, utt_id[u] or u) and counted by the sample index
int audio_add_noise(float* wav, const long long* samp_off, const int* n_samples, const int* utt_id, int n_utts,
                    int max_samples, float sigma, unsigned long long seed, cudaStream_t stream);
// conv0 (Cin=1) + GroupNorm(per channel over time) + GELU -> channels-last bf16
struct Conv0Args {
  const float* x;              // normalised audio, packed
  const long long* samp_off;   // [U]
  const int* L0;               // [U] output frames per utterance
  const long long* out_off;    // [U] first output row per utterance
  const float* w;              // [C, k] fp32 (shared) or per-utterance [U][C*k] when w_stride != 0
  long long w_stride;
  UttParams gn;                // GroupNorm affine via the trainable vector (gn.stride == 0 => shared)
  const float* gn_shared_g;    // used when gn.P == nullptr
  const float* gn_shared_b;
  int g_off, b_off;
  double* stats;               // [U][C][2] sum, sumsq of the conv output per (utterance, channel) (written by the launcher)
  const double* mom;           // [U][k + k(k+1)/2] audio moments from audio_conv0_moments
  bf16* out;                   // [rows, C]
  bf16* pre_out;               // optional: GELU'(normalised value) for the train_feature backward
  int n_utts, C, k, stride, max_L0;
};
int conv0_groupnorm_gelu(const Conv0Args& a, cudaStream_t stream);
// Sx[j] = sum_t x[s t + j] (k values) then R[j][j'] = sum_t x[s t + j] x[s t + j'] for j <= j', per utterance, in double
int audio_conv0_moments(const float* x, const long long* samp_off, const int* L0, double* mom, int k, int stride, int n_utts,
                        int max_L0, cudaStream_t stream);

// LayerNorm feature extractor, layer 0 (HF/modeling_wav2vec2.py:281-292): z0 = Conv1d(1 -> C, k, stride) + bias, channels-last
// bf16 at out_off[u] + t (bias may be null); w fp32 [C][k]
// w_stride / b_stride != 0: per-utterance taps / bias inside the trainable vector (w + u * w_stride), train_feature
int conv0_bias(const float* x, const long long* samp_off, const int* L0, const long long* out_off, const float* w,
               const float* bias, bf16* out, int n_utts, int C, int k, int stride, int max_L0, cudaStream_t stream,
               long long w_stride = 0, long long b_stride = 0);
// row_utt[off[u] + t] = u for t < L[u], every other row of [0, rows) = -1
int fill_row_utt(int* row_utt, long long rows, const long long* off, const int* L, int n_utts, int max_L, cudaStream_t stream);

// ---- convbwd.cu (train_feature backward of the CNN front end) ---------------------------------
int cast_params_bf16(const float* P, long long pstride, long long seg_off, long long size, int n_utts, bf16* out,
                     cudaStream_t stream);
// out[pad_off[u] + t] = d[row] * pre[row] (pre = GELU' saved by the forward; may be null: plain cast) into a 128-row-aligned, zero-gapped slab
int gelu_grad_to_padded(const float* d, const bf16* pre, bf16* out, const int* row_utt, const long long* tok_off,
                        const long long* pad_off, long long M, int C, cudaStream_t stream);
struct Col2imArgs {
  const bf16* Z;               // [rows_in, k*C] dgrad GEMM output of the layer above, (tap, cin) column order
  const bf16* pre;             // [rows_out, C] GELU'(pre-activation) of this layer's output, saved by the forward
  bf16* out;                   // [rows_out, C] d(pre-activation)
  const long long *off_out, *off_in;   // [U] first row of each utterance in this / the upper layer
  const int *L_out, *L_in;     // [U] valid rows
  int C, k, s, n_utts, max_L_out;
};
int conv_col2im_gelu_grad(const Col2imArgs& a, cudaStream_t stream);
struct Conv0BwdArgs {
  const float* x;              // normalised audio
  const long long* samp_off;
  const int* L0;
  const long long* out_off;
  const float* w;              // per-utterance conv0 weight inside the trainable vector: w + u*w_stride
  long long w_stride;
  const bf16* dy;              // [rows0, C] d(GroupNorm output) (already multiplied by GELU')
  const double* stats;         // [U][C][2] forward sum / sumsq
  const double* mom;           // [U][k + k(k+1)/2] audio moments
  float* part;                 // scratch [U][n_chunk][C][k+1] per-chunk partial sums
  int n_chunk;
  const float* P;              // trainable vectors (for gamma)
  float* G;                    // gradient vectors (same layout)
  long long pstride;
  long long g_off, b_off, w_off;
  int n_utts, C, k, stride, max_L0;
  int plain;                   // 1: no GroupNorm behind the conv (lv60 family): dy is d(conv output), the result is just
                               //    d w[c][j] = sum_t dy[t,c] x[s t + j] -> G[w_off ..], d bias[c] = sum_t dy[t,c] -> G[b_off ..]
};
int conv0_groupnorm_backward(const Conv0BwdArgs& a, cudaStream_t stream);
long long conv0_bwd_scratch_floats(int n_utts, int C, int k, int max_L0);
int conv0_bwd_chunks(int max_L0);
int colsum_per_utt(const float* x, const long long* tok_off, const int* T, float* G, long long gstride, long long g_off,
                   int C, int n_utts, cudaStream_t stream);
// the same over bf16 rows (conv bias gradient of the lv60 family: column sums of d z_l over an utterance's rows), fixed order
int colsum_per_utt_bf16(const bf16* x, const long long* row_off, const int* L, float* G, long long gstride, long long g_off,
                        int C, int n_utts, cudaStream_t stream);

// ---- trainall.cu (SUTA_FLAG_TRAIN_ALL) --------------------------------------------------------
// dst [C][R] bf16 = transpose of src [R][C] fp32, for every job of the list in ONE launch (the list rides in the kernel
// parameters: 4 matrices per encoder layer + lm_head, launched MAX jobs at a time)
struct TransposeJobs {
  static constexpr int MAX = 100;
  struct Job { const float* src; bf16* dst; int R, C, tile0; } job[MAX];
  int n = 0;
  void add(const float* src, bf16* dst, int R, int C) { job[n].src = src; job[n].dst = dst; job[n].R = R; job[n].C = C; job[n].tile0 = 0; ++n; }
};
int transpose_cast_bf16(TransposeJobs& jobs, cudaStream_t stream);
// positional conv weight_norm (HF/modeling_wav2vec2.py:344-352, dim = 2): g [K], v [H][CG][K] fp32 inside the trainable vector
// -> w_fwd [H][(tap, ci)], w_bwd [g*CG + ci][(K-1-tap, co)] bf16; scratch keeps ||v|| and g / ||v|| for the backward
long long posconv_weight_norm_scratch_floats(int K);
int posconv_weight_norm_forward(const float* g, const float* v, float* scratch, bf16* w_fwd, bf16* w_bwd, int H, int CG, int K,
                                cudaStream_t stream);
// dW fp32 [H][(tap, ci)] = gradient of the folded weight -> dg [K], dv [H][CG][K]
int posconv_weight_norm_backward(const float* v, const float* dW, float* scratch, float* dg, float* dv, int H, int CG, int K,
                                 cudaStream_t stream);

// ---- posconv_tc.cu --------------------------------------------------------------------------
// grouped positional conv on tcgen05 with a shared-memory-resident input window (CG = H/G in {48, 64})
bool posconv_tc_supported(int CG, int pos_k);
int posconv_tc(const bf16* xg, const bf16* w, const float* bias, float* out, int out_ld, int G, int CG, long long R, long long Rm,
               int pos_k, cudaStream_t stream);

// ---- posconv.cu -----------------------------------------------------------------------------
// scatter fp32 [M,H] tokens into the zero-padded per-group bf16 layout [G][R][CGP] used as the implicit-GEMM A operand
int posconv_pack(const float* h, const int* row_utt, const long long* tok_off, const long long* pad_off, bf16* xg,
                 long long M, int H, int G, int CGP, long long R, cudaStream_t stream);
// h_out = h + GELU(conv[pad_row(u,t)])   (conv already contains the bias)
int posconv_combine(const float* h, const float* conv, const int* row_utt, const long long* tok_off,
                    const long long* pad_off, float* h_out, long long M, int H, int row_shift, cudaStream_t stream);
// dC = d_out * GELU'(conv) scattered into the padded per-group layout (backward A operand)
int posconv_pack_grad(const float* d_out, const float* conv, const int* row_utt, const long long* tok_off,
                      const long long* pad_off, bf16* dg, long long M, int H, int G, int CGP, long long R, int row_shift,
                      cudaStream_t stream);
// d_h = d_out + dconv[pad_row(u,t)]   (residual path + conv path)
int posconv_combine_grad(const float* d_out, const float* dconv, const int* row_utt, const long long* tok_off,
                         const long long* pad_off, float* d_h, bf16* d_h_bf16, long long M, int H, int row_shift,
                         cudaStream_t stream);

// ---- loss.cu --------------------------------------------------------------------------------
struct LossArgs {
  const float* logits;         // [M, 32]
  const long long* tok_off;    // [U]
  const int* T;                // [U]
  float* dlogits_f32;          // optional [M,32]
  bf16* dlogits_bf16;          // optional [M,32]
  float* loss;                 // [U] total, then [U] entropy term, then [U] mcc term
  int n_utts;
  float em_coef, temp;
  int reweight, not_blank;
  float div_coef;              // REF/main.py:201-203 (0: term skipped)
};
int suta_loss_forward_backward(const LossArgs& a, cudaStream_t stream);
// out[row] = entropy of softmax(logits[row] / temp)   (REF/main.py:26-28), V = 32
int softmax_entropy_rows(const float* logits, long long rows, float temp, float* out, cudaStream_t stream);

// ---- ctc.cu ---------------------------------------------------------------------------------
// Pseudo-label CTC loss of the SDPL baseline (REF/main_SDPL.py:194-209), value + gradient, mixed into dlogits:
// dlogits <- (1 - pl_coef) * dlogits + pl_coef * d ctc / d logits;  loss[u] <- (1 - pl_coef) * loss[u] + pl_coef * ctc
struct CtcArgs {
  const float* logits;         // [M, 32]
  const long long* tok_off;    // [U]
  const int* T;                // [U]
  const int* collapsed;        // [M] greedy transcript ids (repeat-collapsed, blank dropped) of utterance u at tok_off[u]
  const int* collapsed_len;    // [U]
  float* alpha;                // scratch: utterance u uses ctc_alpha_floats(T_u) floats at alpha_off[u]
  const long long* alpha_off;  // [U]
  float* g;                    // scratch [M, 32]
  float* dlogits_f32;          // in/out [M, 32] (may be null: treated as zero)
  bf16* dlogits_bf16;          // out, optional
  float* loss;                 // [4][U]: row 0 (total) is mixed, row 3 receives the CTC term; optional
  int* target_len;             // [U] optional: pseudo-label length L
  int n_utts, max_states;      // max_states >= 2 * max(T) + 1
  float pl_coef;
};
int ctc_pseudo_label_loss(const CtcArgs& a, cudaStream_t stream);
long long ctc_alpha_floats(int T);

// ---- optim.cu -------------------------------------------------------------------------------
struct AdamArgs {
  float* P;                    // [U][n]
  const float* G;              // [U][n]
  float* Mom;                  // [U][n]
  float* Var;                  // [U][n]
  const unsigned char* mult;   // [n] multiplicity k of each element (REF/main.py:62-103 duplicates), 1..4
  long long n;
  int n_utts;
  int step_index;              // number of optimizer.step() calls already applied since reset
  float lr, beta1, beta2, eps, weight_decay;
  int kind;                    // 0 AdamW (decoupled decay), 1 SGD (L2 decay), 2 Adam (L2 decay: g += wd p, torch/optim/adam.py)
  bf16* shadow;                // optional bf16 copy of P (GEMM operands of trainable weights), same layout
  // optional bf16 copies of selected segments, each stored [U][size] contiguously (the stacked per-utterance B operands
  // of the conv / projection GEMMs): written by the same pass that updates P
  struct Seg { long long off, size; bf16* dst; };
  Seg seg[8];
  int n_seg;
};
int optimizer_step(const AdamArgs& a, cudaStream_t stream);
int params_reset(float* P, const float* P0, float* Mom, float* Var, bf16* shadow, long long n, int n_utts,
                 cudaStream_t stream);

// ---- decode.cu ------------------------------------------------------------------------------
// ids[row] = argmax(logits[row]); collapsed[u_off + j] = j-th kept id (repeat-collapsed, blank dropped); out_len[u]
int ctc_greedy_decode(const float* logits, const long long* tok_off, const int* T, int* ids, int* collapsed,
                      int* out_len, int n_utts, int V, cudaStream_t stream);
