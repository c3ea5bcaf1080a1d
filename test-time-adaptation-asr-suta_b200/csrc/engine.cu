// suta_b200 engine: host-side orchestration of one batched SUTA adaptation (C ABI in include/suta_b200.h).
//
// The reference adapts ONE utterance at a time (REF/main.py:319-402): reset -> forward -> {loss, backward,
// optimizer step, forward} x steps.  This engine runs the same arithmetic for a batch of independent utterances
// at once: frozen weights are shared, every utterance owns a private copy of the trainable vector (LayerNorm
// affine; plus the CNN front end under train_feature), tokens of all utterances are packed on one M axis so the
// tcgen05 GEMMs see M = sum(T_u), and every cross-frame operator (GroupNorm statistics, positional conv padding,
// attention, loss reductions) is confined to its own utterance.
//
// Work per adaptation step = 1 backward + 1 forward (the reference's "repeat_inference" forward of step i IS the
// training forward of step i+1, REF/main.py:181 vs :212-214); LayerNorm-only mode runs the CNN once per utterance.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/suta_b200.h"
#include "gemm_tc.cuh"
#include "kernels.cuh"

static thread_local char g_err[1024] = "";
extern "C" void suta_set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* suta_last_error(void) { return g_err; }
extern "C" int suta_abi_version(void) { return SUTA_ABI_VERSION; }
extern "C" int suta_device_sm_count(void) { return gemm_num_sms(); }

namespace {

struct Bump {            // bump allocator over the caller's workspace (pass 1: base == nullptr just sizes it)
  uint8_t* base = nullptr;
  size_t off = 0;
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct LayerBufs {
  bf16* qkv;      // [M,3H]
  bf16* attn;     // [M,H]
  float* lse;     // [heads,M]
  float* h1;      // [M,H] pre-LN1
  float *mean1, *rstd1;
  bf16* pre;      // [M,I] GELU'(FFN pre-activation)
  float* h2;      // [M,H] pre-LN2
  float *mean2, *rstd2;
};

struct Seg { int kind, module, index; long long off, size; };

}  // namespace

struct suta_engine {
  suta_model_cfg cfg{};
  int train_feature = 0;
  // the "lv60" family: conv_ln = every conv layer is Conv1d(+bias) -> LayerNorm(C) -> GELU with TRAINABLE LayerNorms (so the
  // backward spans the CNN even in LayerNorm-only mode, through the frozen shared conv weights); stable = pre-LN encoder;
  // cnn_bwd = the batch layout / tables / buffers a backward through the CNN needs (train_feature or conv_ln)
  int conv_ln = 0, stable = 0, cnn_bwd = 0;
  int pseudo_label = 0;                            // SUTA_FLAG_PSEUDO_LABEL: CTC scratch is part of every batch workspace
  // SUTA_FLAG_TRAIN_ALL (REF/main.py:96-100): every weight and bias of the model lives in the trainable vector (train_feature's
  // layout plus the segments below); the GEMM operands are bf16 copies of that vector (Pb) and transposed copies for the
  // dgrads, refreshed after every update; one utterance per batch (nothing is shared between utterances any more)
  int train_all = 0;
  std::vector<long long> wqkv_off, bqkv_off, wo_off, bo_off, w1_off, b1_off, w2_off, b2_off;
  long long lm_w_off = 0, lm_b_off = 0, pos_g_off = 0, pos_v_off = 0, pos_b_off = 0;
  struct TaLayer {
    bf16 *wqkv_t, *wo_t, *w1_t, *w2_t;             // [H,3H], [H,H], [H,I], [I,H]
    bf16 *xa, *xf, *gl;                            // saved GEMM inputs of the layer: LayerNorm outputs [M,H] x2, GELU output [M,I]
  };
  std::vector<TaLayer> ta;
  bf16 *Pb = nullptr, *lm_w_t_sh = nullptr, *pos_w_sh = nullptr, *pos_w_t_sh = nullptr, *x_last16 = nullptr, *xg_grad = nullptr;
  float *pos_dW = nullptr, *wn_scratch = nullptr;
  suta_weights w{};
  bool have_weights = false;
  std::vector<Seg> segs;
  long long n_params = 0;
  // segment offsets
  long long fp_g = 0, fp_b = 0, enc_g = 0, enc_b = 0;
  std::vector<long long> ln1_g, ln1_b, ln2_g, ln2_b;
  std::vector<long long> cln_g, cln_b;             // conv-layer LayerNorms (conv_ln)
  // train_feature segments (REF/main.py:88-94): GroupNorm affine, every conv weight, projection weight + bias
  long long ln_params = 0;                         // size of the LayerNorm-only prefix of the trainable vector
  long long gn_g = 0, gn_b = 0, proj_w_off = 0, proj_b_off = 0;
  long long conv_w_off[SUTA_MAX_CONV] = {};
  long long conv_w_size[SUTA_MAX_CONV] = {};
  long long conv_b_off[SUTA_MAX_CONV] = {};        // conv biases (lv60 family under train_feature)
  long long launches = 0;
  // optional per-launch GEMM timing (bench.py roofline leg): CUDA event pairs around every tcgen05 GEMM launch
  bool profile = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct ProfRec { std::string tag; double flops; bool is_gemm; double bytes; };
  std::vector<ProfRec> prof_recs;
  std::string prof_report;
  bool audio_normalized = false;
  double sumT2 = 0.0;                              // sum over utterances of T_u^2 (attention FLOP accounting)

  // ---- batch state ----
  int U = 0;
  long long M = 0, S = 0, R = 0, Rm = 0;
  int max_samples = 0, max_L0 = 0;
  std::vector<int> n_samples;
  std::vector<std::vector<int>> L;                 // [layer][u]
  std::vector<std::vector<long long>> off;         // [layer][u] first row in layer output
  std::vector<long long> rows_total;               // [layer]
  std::vector<long long> samp_off, tok_off, pad_off;
  std::vector<int> T;
  std::vector<int> n_mblk;                         // [layer]
  std::vector<int> n_dg_mblk;                      // [layer] M-blocks of the conv dgrad GEMM(s)
  std::vector<int> n_mpair;                        // [layer] 256-row pair tiles of the forward conv GEMM (train_feature)
  int4* d_mpair[SUTA_MAX_CONV] = {};
  int n_attn_blk = 0;
  bool frontend_done = false;
  int opt_steps = 0;

  // device tables
  long long *d_samp_off = nullptr, *d_tok_off = nullptr, *d_pad_off = nullptr, *d_off0 = nullptr;
  int *d_n_samples = nullptr, *d_T = nullptr, *d_L0 = nullptr, *d_row_utt = nullptr;
  int4* d_mblk[SUTA_MAX_CONV] = {};
  int4* d_attn_tab = nullptr;
  // train_feature only
  std::vector<long long> off64;                    // [u] first row of utterance u in the 128-row-aligned token slabs
  long long R64 = 0;
  int n_tok_mblk = 0;
  int4* d_tok_mblk = nullptr;                      // per-utterance M-blocks over packed tokens (per-utterance projection)
  int4* d_dgrad_mblk[SUTA_MAX_CONV] = {};          // per-utterance M-blocks over layer-l rows (conv dgrad)
  int4* d_ztab[SUTA_MAX_CONV + 1] = {};            // per-utterance reduction ranges of the wgrad GEMMs ([n_conv] = projection)
  long long* d_off[SUTA_MAX_CONV] = {};            // [U] row offsets per layer
  long long* d_dpre_off_last = nullptr;            // = off64 on device
  int* d_L[SUTA_MAX_CONV] = {};                    // [U] valid rows per layer
  bf16* conv_pre[SUTA_MAX_CONV] = {};              // GELU'(pre-activation) saved by the forward (bf16)
  bf16* conv_dpre[SUTA_MAX_CONV] = {};             // d(pre-activation), zero outside valid rows
  // conv_ln only: conv output before its LayerNorm (bf16, saved), per-row statistics, row -> utterance tables of the
  // layer's row space (-1 in the gaps; the last layer is token-packed and uses d_row_utt)
  bf16* conv_z[SUTA_MAX_CONV] = {};
  float *conv_mean[SUTA_MAX_CONV] = {}, *conv_rstd[SUTA_MAX_CONV] = {};
  int* d_conv_row_utt[SUTA_MAX_CONV] = {};
  bf16* w_shadow[SUTA_MAX_CONV] = {};              // bf16 copies of the per-utterance conv weights [U][Cout][k*Cin]
  bf16* proj_shadow = nullptr;                     // [U][H][C]
  bf16* zbuf = nullptr;                            // dgrad GEMM output [rows_l, k*Cin]
  bf16* dh0_pad = nullptr;                         // [R64, H]
  float* d_feat = nullptr;                         // [M, C]
  float* c0_scratch = nullptr;
  double* mom = nullptr;                           // [U][k + k(k+1)/2] audio moments (conv0 GroupNorm statistics / backward)
  bool moments_done = false;
  bool z0_done = false;                            // conv_ln: conv0 + bias of the current audio is in conv_z[0]
  int max_L[SUTA_MAX_CONV] = {};
  // device buffers
  float *wav = nullptr, *wav_norm = nullptr;
  double* stats = nullptr;
  bf16* conv_out[SUTA_MAX_CONV] = {};
  float *fp_mean = nullptr, *fp_rstd = nullptr;
  bf16* y_fp = nullptr;
  float* h0 = nullptr;
  bf16* xg = nullptr;
  float *cpos = nullptr, *dcpos = nullptr;
  float* hE = nullptr;
  float *enc_mean = nullptr, *enc_rstd = nullptr;
  float *fa = nullptr, *fb = nullptr;
  bf16 *b16 = nullptr, *gelu16 = nullptr, *dpre16 = nullptr, *dqkv16 = nullptr, *dO16 = nullptr;
  float* Dbuf = nullptr;
  float* d_yfp = nullptr;
  std::vector<LayerBufs> lb;
  float *logits = nullptr, *dlogits = nullptr, *losses = nullptr;
  bf16* dlogits16 = nullptr;
  float *P = nullptr, *G = nullptr, *Mom = nullptr, *Var = nullptr;
  float *ctc_alpha = nullptr, *ctc_g = nullptr;    // pseudo-label CTC: alpha lattices, d loss / d log-prob scratch
  long long* d_alpha_off = nullptr;
  std::vector<long long> alpha_off;
  int* ctc_tlen = nullptr;
  float* ln_part = nullptr;                        // dgamma/dbeta slots of the two-stage LayerNorm-backward reduction
  size_t ln_part_stride = 0;
  int *ids = nullptr, *collapsed = nullptr, *out_len = nullptr;
  // every device table of a batch lives in one contiguous region at the start of the workspace and is uploaded by ONE
  // host-to-device copy from this pinned mirror (no pageable copies, no stream synchronisation in suta_batch_begin)
  // CUDA graphs of the two launch chains of an adaptation step.  A small batch (the reference's operating point: one
  // utterance at a time) is launch-bound -- ~250 kernels of a few microseconds per step -- and its launch sequence does
  // not change between the steps of a batch: the second call with the same key records the chain on a side stream
  // (the caller's may be the legacy default stream, which cannot be captured), later calls replay it.
  struct GraphSlot {
    cudaGraphExec_t exec = nullptr;
    unsigned long long key = 0, seen_key = 0;      // what the recorded chain depends on besides the batch layout
    bool seen = false;
    long long launches = 0;                        // launch-counter increment of one replay
    bool post_frontend_done = false, post_moments_done = false, post_z0_done = false;   // host state after the chain (forward)
  };
  GraphSlot g_fwd, g_bwd;
  cudaStream_t cap_stream = nullptr;
  long long graph_replays = 0;                     // chains launched as graphs since the engine was created
  size_t tables_bytes = 0;
  uint8_t* h_stage = nullptr;
  size_t h_stage_cap = 0;
  cudaEvent_t stage_ev = nullptr;                  // recorded after the upload: the mirror may be rewritten once it fired
};

static void drop_graphs(suta_engine* e);

namespace {

int build_layout(suta_engine* e) {
  const suta_model_cfg& c = e->cfg;
  e->segs.clear();
  long long o = 0;
  auto add = [&](int kind, int module, int index, long long size) {
    e->segs.push_back({kind, module, index, o, size});
    long long r = o;
    o += size;
    return r;
  };
  const int C = c.conv_dim[c.n_conv - 1], H = c.hidden;
  e->fp_g = add(0, 0, 0, C);
  e->fp_b = add(1, 0, 0, C);
  e->enc_g = add(0, 1, 0, H);
  e->enc_b = add(1, 1, 0, H);
  e->ln1_g.resize(c.layers); e->ln1_b.resize(c.layers); e->ln2_g.resize(c.layers); e->ln2_b.resize(c.layers);
  for (int l = 0; l < c.layers; ++l) {
    e->ln1_g[l] = add(0, 2, l, H);
    e->ln1_b[l] = add(1, 2, l, H);
    e->ln2_g[l] = add(0, 3, l, H);
    e->ln2_b[l] = add(1, 3, l, H);
  }
  if (e->conv_ln) {                  // HF/modeling_wav2vec2.py:288: one LayerNorm(C_l) per conv layer, collected by REF/main.py:81-87
    e->cln_g.resize(c.n_conv); e->cln_b.resize(c.n_conv);
    for (int l = 0; l < c.n_conv; ++l) {
      e->cln_g[l] = add(0, 6, l, c.conv_dim[l]);
      e->cln_b[l] = add(1, 6, l, c.conv_dim[l]);
    }
  }
  e->ln_params = o;
  if (e->train_feature) {
    if (!e->conv_ln) {
      e->gn_g = add(2, 4, 0, c.conv_dim[0]);
      e->gn_b = add(3, 4, 0, c.conv_dim[0]);
    }
    for (int l = 0; l < c.n_conv; ++l) {
      e->conv_w_size[l] = (long long)c.conv_dim[l] * c.conv_kernel[l] * (l ? c.conv_dim[l - 1] : 1);
      e->conv_w_off[l] = add(4, 4, l, e->conv_w_size[l]);     // packed [Cout][(tap, Cin)]
    }
    if (e->conv_ln)                                           // HF:286 conv_bias=True: Conv1d biases are parameters of the feature extractor too
      for (int l = 0; l < c.n_conv; ++l) e->conv_b_off[l] = add(7, 4, l, c.conv_dim[l]);
    e->proj_w_off = add(5, 5, 0, (long long)H * C);
    e->proj_b_off = add(6, 5, 0, H);
  }
  if (e->train_all) {
    const int I = c.intermediate, K = c.pos_k, CG = H / c.pos_groups;
    e->pos_g_off = add(8, 14, 0, K);                               // weight_norm: g [1,1,K], v [H, H/G, K] (HF layout), bias [H]
    e->pos_v_off = add(9, 14, 0, (long long)H * CG * K);
    e->pos_b_off = add(6, 14, 0, H);
    auto v = [&](std::vector<long long>& x) { x.assign(c.layers, 0); };
    v(e->wqkv_off); v(e->bqkv_off); v(e->wo_off); v(e->bo_off); v(e->w1_off); v(e->b1_off); v(e->w2_off); v(e->b2_off);
    for (int l = 0; l < c.layers; ++l) {
      e->wqkv_off[l] = add(5, 7, l, (long long)H * H);             // q, k, v weights back to back = the fused [3H, H] operand
      add(5, 8, l, (long long)H * H);
      add(5, 9, l, (long long)H * H);
      e->bqkv_off[l] = add(6, 7, l, H);
      add(6, 8, l, H);
      add(6, 9, l, H);
      e->wo_off[l] = add(5, 10, l, (long long)H * H);
      e->bo_off[l] = add(6, 10, l, H);
      e->w1_off[l] = add(5, 11, l, (long long)I * H);
      e->b1_off[l] = add(6, 11, l, I);
      e->w2_off[l] = add(5, 12, l, (long long)H * I);
      e->b2_off[l] = add(6, 12, l, H);
    }
    e->lm_w_off = add(5, 13, 0, (long long)c.vocab * H);
    e->lm_b_off = add(6, 13, 0, c.vocab);
  }
  e->n_params = o;
  return SUTA_OK;
}

int check_cfg(const suta_model_cfg& c) {
  SUTA_CHECK_ARG(c.hidden > 0 && c.hidden % 64 == 0 && c.heads * 64 == c.hidden);
  SUTA_CHECK_ARG(c.layers > 0 && c.layers <= SUTA_MAX_LAYERS && c.intermediate % 64 == 0);
  SUTA_CHECK_ARG(c.vocab == 32);
  SUTA_CHECK_ARG(c.n_conv >= 2 && c.n_conv <= SUTA_MAX_CONV);
  for (int l = 0; l < c.n_conv; ++l) SUTA_CHECK_ARG(c.conv_dim[l] % 64 == 0 && c.conv_kernel[l] > 0 && c.conv_stride[l] > 0);
  SUTA_CHECK_ARG(c.pos_k > 0 && c.pos_k % 2 == 0 && c.pos_groups > 0 && c.hidden % c.pos_groups == 0);
  SUTA_CHECK_ARG((c.hidden / c.pos_groups) % 16 == 0);
  return SUTA_OK;
}

static bool use_posconv_tc(const suta_engine* e) {
  const suta_model_cfg& c = e->cfg;
  static const bool on = getenv("SUTA_NO_POSCONV_TC") == nullptr;
  return on && posconv_tc_supported(c.hidden / c.pos_groups, c.pos_k);
}

// Transposed-conv dgrad of layer l as GEMMs that write d(input) directly (no [rows, k*Cin] intermediate, no col2im):
// stride 2 with k = 2 (rows 2j, 2j+1 are the two halves of one N = 2*Cin output row) or k = 3 (even rows 2j =
// dY[j] W_0 + dY[j-1] W_2, odd rows 2j+1 = dY[j] W_1).
static bool dgrad_fused(const suta_engine* e, int l) {
  const suta_model_cfg& c = e->cfg;
  return e->cnn_bwd && l >= 1 && c.conv_stride[l] == 2 && (c.conv_kernel[l] == 2 || c.conv_kernel[l] == 3) &&
         c.conv_dim[l] % 64 == 0 && c.conv_dim[l - 1] % 64 == 0 &&
         (e->conv_ln || getenv("SUTA_NO_FUSED_DGRAD") == nullptr);   // the debug switch selects train_feature's col2im path only
}

// geometry of a batch; fills host vectors, returns workspace size through the bump allocator
int plan_batch(suta_engine* e, int U, const int32_t* n_samples) {
  const suta_model_cfg& c = e->cfg;
  SUTA_CHECK_ARG(U > 0 && n_samples);
  e->U = U;
  e->n_samples.assign(n_samples, n_samples + U);
  e->L.assign(c.n_conv, std::vector<int>(U));
  e->off.assign(c.n_conv, std::vector<long long>(U));
  e->rows_total.assign(c.n_conv, 0);
  e->samp_off.resize(U); e->tok_off.resize(U); e->pad_off.resize(U); e->T.resize(U);
  e->n_mblk.assign(c.n_conv, 0);
  e->n_dg_mblk.assign(c.n_conv, 0);
  e->n_mpair.assign(c.n_conv, 0);
  long long s = 0;
  e->max_samples = 0; e->max_L0 = 0;
  for (int u = 0; u < U; ++u) {
    int n = n_samples[u];
    int Lc = n;
    for (int l = 0; l < c.n_conv; ++l) {
      Lc = (Lc - c.conv_kernel[l]) / c.conv_stride[l] + 1;
      if (Lc < 1 || (l == 0 && n < c.conv_kernel[0])) {
        suta_set_last_error("utterance %d with %d samples is too short for the feature extractor", u, n);
        return SUTA_ERR_ARG;
      }
      e->L[l][u] = Lc;
    }
    e->samp_off[u] = s;
    s += (n + 3) & ~3;
    e->max_samples = n > e->max_samples ? n : e->max_samples;
    e->max_L0 = e->L[0][u] > e->max_L0 ? e->L[0][u] : e->max_L0;
  }
  e->S = s;
  for (int l = 0; l < c.n_conv; ++l) {
    long long r = 0;
    const bool last = l == c.n_conv - 1;
    for (int u = 0; u < U; ++u) {
      e->off[l][u] = r;
      // rows of the next layer's implicit-GEMM view start at off/stride: keep offsets multiples of 8; under
      // train_feature every utterance starts on a 128-row boundary, so (a) the weight-gradient GEMMs can reduce over
      // time in 64-row steps and (b) every 128-row GEMM tile owns its output rows outright (TMA-store epilogue)
      // (c) the parity-split conv dgrad (conv_backward) writes 256 rows of layer l per 128-row tile of layer l + 1 and
      // reads one zero row past the utterance's last gradient row: 256-row alignment with at least one spare row
      r += last ? e->L[l][u] : (e->cnn_bwd ? ((e->L[l][u] + 1 + 255) & ~255) : ((e->L[l][u] + 7) & ~7));
      if (l >= 1) {
        e->n_mblk[l] += ceil_div(e->L[l][u], 128);
        e->n_mpair[l] += ceil_div(e->L[l][u], 256);
        e->n_dg_mblk[l] += ceil_div(e->L[l][u] + (dgrad_fused(e, l) && !last ? 1 : 0), 128);
      }
      e->max_L[l] = u == 0 ? e->L[l][u] : (e->L[l][u] > e->max_L[l] ? e->L[l][u] : e->max_L[l]);
    }
    e->rows_total[l] = r;
    if (!last) SUTA_CHECK_ARG(8 % c.conv_stride[l + 1] == 0);
  }
  long long m = 0, p = c.pos_k / 2;
  int nblk = 0;
  for (int u = 0; u < U; ++u) {
    e->T[u] = e->L[c.n_conv - 1][u];
    e->tok_off[u] = m;
    e->pad_off[u] = p;
    m += e->T[u];
    p += e->T[u] + c.pos_k / 2;
    nblk += ceil_div(e->T[u], 128);
  }
  e->sumT2 = 0.0;
  for (int u = 0; u < U; ++u) e->sumT2 += (double)e->T[u] * e->T[u];
  e->M = m;
  e->off64.assign(U, 0);
  e->n_tok_mblk = 0;
  long long r64 = 0;
  for (int u = 0; u < U; ++u) {
    e->off64[u] = r64;
    r64 += (e->T[u] + 127) & ~127;
    e->n_tok_mblk += ceil_div(e->T[u], 128);
  }
  e->R64 = r64;
  e->R = p;                       // padded rows per group slab (leading, between-utterance and trailing zero rows)
  e->Rm = e->R - c.pos_k + 1;     // number of K-tap windows
  e->n_attn_blk = nblk;
  return SUTA_OK;
}

void carve(suta_engine* e, Bump& b) {
  const suta_model_cfg& c = e->cfg;
  const int U = e->U, H = c.hidden, I = c.intermediate, C = c.conv_dim[c.n_conv - 1], V = c.vocab;
  const long long M = e->M;
  e->d_samp_off = b.take<long long>(U); e->d_tok_off = b.take<long long>(U); e->d_pad_off = b.take<long long>(U);
  e->d_off0 = b.take<long long>(U);
  e->d_n_samples = b.take<int>(U); e->d_T = b.take<int>(U); e->d_L0 = b.take<int>(U);
  e->d_row_utt = b.take<int>(M);
  for (int l = 1; l < c.n_conv; ++l) e->d_mblk[l] = b.take<int4>(e->n_mblk[l]);
  e->d_attn_tab = b.take<int4>(e->n_attn_blk);
  if (e->cnn_bwd) {
    if (e->train_feature) {
      e->d_tok_mblk = b.take<int4>(e->n_tok_mblk);
      e->d_ztab[c.n_conv] = b.take<int4>(U);
    }
    e->d_dpre_off_last = b.take<long long>(U);
    for (int l = 0; l < c.n_conv; ++l) {
      e->d_off[l] = b.take<long long>(U);
      e->d_L[l] = b.take<int>(U);
      if (l >= 1) {
        e->d_mpair[l] = b.take<int4>(e->n_mpair[l]);
        e->d_dgrad_mblk[l] = b.take<int4>(e->n_dg_mblk[l]);
        if (e->train_feature) e->d_ztab[l] = b.take<int4>(U);
      }
    }
  }
  if (e->pseudo_label) e->d_alpha_off = b.take<long long>(U);
  e->tables_bytes = align_up(b.off, 256);          // everything above is uploaded by suta_batch_begin in one copy
  e->wav = b.take<float>(e->S); e->wav_norm = b.take<float>(e->S + 64);
  e->stats = b.take<double>((size_t)2 * U * (c.conv_dim[0] > 1 ? c.conv_dim[0] : 1) + 2 * U);
  e->mom = b.take<double>((size_t)U * (c.conv_kernel[0] + c.conv_kernel[0] * (c.conv_kernel[0] + 1) / 2));
  for (int l = 0; l < c.n_conv; ++l) e->conv_out[l] = b.take<bf16>((size_t)(e->rows_total[l] + 128) * c.conv_dim[l]);
  e->fp_mean = b.take<float>(M); e->fp_rstd = b.take<float>(M);
  e->y_fp = b.take<bf16>((size_t)M * C);
  e->h0 = b.take<float>((size_t)M * H);
  e->xg = b.take<bf16>((size_t)(e->R + 8) * H);
  e->cpos = b.take<float>((size_t)e->Rm * H);
  e->dcpos = b.take<float>((size_t)e->Rm * H);
  e->hE = b.take<float>((size_t)M * H);
  e->enc_mean = b.take<float>(M); e->enc_rstd = b.take<float>(M);
  e->fa = b.take<float>((size_t)M * H); e->fb = b.take<float>((size_t)M * H);
  e->b16 = b.take<bf16>((size_t)M * H);
  e->gelu16 = b.take<bf16>((size_t)M * I);
  e->dpre16 = b.take<bf16>((size_t)M * I);
  e->dqkv16 = b.take<bf16>((size_t)M * 3 * H);
  e->dO16 = b.take<bf16>((size_t)M * H);
  e->Dbuf = b.take<float>((size_t)M * c.heads);
  e->d_yfp = b.take<float>((size_t)M * C);
  e->lb.resize(c.layers);
  for (int l = 0; l < c.layers; ++l) {
    LayerBufs& x = e->lb[l];
    x.qkv = b.take<bf16>((size_t)M * 3 * H);
    x.attn = b.take<bf16>((size_t)M * H);
    x.lse = b.take<float>((size_t)M * c.heads);
    x.h1 = b.take<float>((size_t)M * H);
    x.mean1 = b.take<float>(M); x.rstd1 = b.take<float>(M);
    x.pre = b.take<bf16>((size_t)M * I);
    x.h2 = b.take<float>((size_t)M * H);
    x.mean2 = b.take<float>(M); x.rstd2 = b.take<float>(M);
  }
  e->logits = b.take<float>((size_t)M * V); e->dlogits = b.take<float>((size_t)M * V);
  e->dlogits16 = b.take<bf16>((size_t)M * V);
  e->losses = b.take<float>(4 * U);
  if (e->pseudo_label) {
    long long tot = 0;
    e->alpha_off.assign(U, 0);
    for (int u = 0; u < U; ++u) { e->alpha_off[u] = tot; tot += ctc_alpha_floats(e->T[u]); }
    e->ctc_alpha = b.take<float>((size_t)tot);
    e->ctc_g = b.take<float>((size_t)M * V);
    e->ctc_tlen = b.take<int>(U);
  }
  e->P = b.take<float>((size_t)U * e->n_params); e->G = b.take<float>((size_t)U * e->n_params);
  e->Mom = b.take<float>((size_t)U * e->n_params); e->Var = b.take<float>((size_t)U * e->n_params);
  e->ids = b.take<int>(M); e->collapsed = b.take<int>(M); e->out_len = b.take<int>(U);
  e->ln_part_stride = (size_t)layernorm_backward_scratch_floats(H > C ? H : C, U);
  e->ln_part = b.take<float>(e->ln_part_stride * (2 * c.layers + 2 + (e->conv_ln ? c.n_conv : 0)));   // one slot set per LayerNorm: reduced together
  if (e->conv_ln) {
    for (int l = 0; l < c.n_conv; ++l) {
      const bool last = l == c.n_conv - 1;
      const long long rows = last ? e->R64 : e->rows_total[l];
      e->conv_z[l] = b.take<bf16>((size_t)(e->rows_total[l] + 128) * c.conv_dim[l]);
      e->conv_mean[l] = b.take<float>((size_t)e->rows_total[l] + 128);
      e->conv_rstd[l] = b.take<float>((size_t)e->rows_total[l] + 128);
      if (!last) e->d_conv_row_utt[l] = b.take<int>((size_t)e->rows_total[l] + 128);
      // gradient of the layer's output, turned IN PLACE into the gradient of its pre-LayerNorm value (128 leading zero
      // rows: the even-row dgrad GEMM reads row -1 of the first utterance)
      e->conv_dpre[l] = b.take<bf16>((size_t)(rows + 256) * c.conv_dim[l]) + (size_t)128 * c.conv_dim[l];
    }
    e->d_feat = b.take<float>((size_t)M * C);
  }
  if (e->train_all) e->Pb = b.take<bf16>((size_t)e->n_params);      // U == 1: the conv / projection operand copies live in it too
  if (e->train_feature) {
    size_t zmax = 0;
    for (int l = 0; l < c.n_conv; ++l) {
      const bool last = l == c.n_conv - 1;
      const long long rows = last ? e->R64 : e->rows_total[l];
      if (!e->conv_ln) {         // (the LayerNorm feature extractor keeps z_l instead of GELU' and owns its d z_l buffers above)
        e->conv_pre[l] = b.take<bf16>((size_t)(e->rows_total[l] + 128) * c.conv_dim[l]);
        // 128 leading rows (zero): the even-row dgrad GEMM reads row -1 of the first utterance
        e->conv_dpre[l] = b.take<bf16>((size_t)(rows + 256) * c.conv_dim[l]) + (size_t)128 * c.conv_dim[l];
      }
      if (l >= 1) {
        e->w_shadow[l] = e->train_all ? e->Pb + e->conv_w_off[l] : b.take<bf16>((size_t)U * e->conv_w_size[l]);
        size_t z = dgrad_fused(e, l) ? 0 : (size_t)(rows + 128) * c.conv_kernel[l] * c.conv_dim[l - 1];
        zmax = z > zmax ? z : zmax;
      }
    }
    e->zbuf = b.take<bf16>(zmax);
    e->proj_shadow = e->train_all ? e->Pb + e->proj_w_off : b.take<bf16>((size_t)U * H * C);
    e->dh0_pad = b.take<bf16>((size_t)(e->R64 + 128) * H);
    if (!e->conv_ln) e->d_feat = b.take<float>((size_t)M * C);
    e->c0_scratch = b.take<float>((size_t)conv0_bwd_scratch_floats(U, c.conv_dim[0], c.conv_kernel[0], e->max_L0));
  }
  if (e->train_all) {
    const size_t K = c.pos_k, CG = H / c.pos_groups;
    e->ta.resize(c.layers);
    for (int l = 0; l < c.layers; ++l) {
      suta_engine::TaLayer& t = e->ta[l];
      t.wqkv_t = b.take<bf16>((size_t)3 * H * H); t.wo_t = b.take<bf16>((size_t)H * H);
      t.w1_t = b.take<bf16>((size_t)I * H); t.w2_t = b.take<bf16>((size_t)H * I);
      t.xa = b.take<bf16>((size_t)M * H); t.xf = b.take<bf16>((size_t)M * H); t.gl = b.take<bf16>((size_t)M * I);
    }
    e->x_last16 = b.take<bf16>((size_t)M * H);
    e->lm_w_t_sh = b.take<bf16>((size_t)H * V);
    e->pos_w_sh = b.take<bf16>((size_t)H * K * CG); e->pos_w_t_sh = b.take<bf16>((size_t)H * K * CG);
    e->xg_grad = b.take<bf16>((size_t)(e->R + 8) * H);
    e->pos_dW = b.take<float>((size_t)H * K * CG);
    e->wn_scratch = b.take<float>((size_t)posconv_weight_norm_scratch_floats(c.pos_k));
  }
  b.off = align_up(b.off, 256);
}

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Where the encoder's GEMM operands live: the frozen shared weights handed over by suta_engine_set_weights, or (train_all)
// the bf16 copies of the utterance's own trainable vector (Pb, transposed copies) and fp32 biases inside P (U == 1).
struct LayerW {
  const bf16 *wqkv, *wqkv_t, *wo, *wo_t, *w1, *w1_t, *w2, *w2_t;
  const float *bqkv, *bo, *b1, *b2;
};
inline const bf16* B16(const void* p) { return reinterpret_cast<const bf16*>(p); }
LayerW layer_w(const suta_engine* e, int l) {
  if (!e->train_all) {
    const suta_layer_weights& w = e->w.layer[l];
    return {B16(w.wqkv), B16(w.wqkv_t), B16(w.wo), B16(w.wo_t), B16(w.w1), B16(w.w1_t), B16(w.w2), B16(w.w2_t), w.bqkv, w.bo, w.b1, w.b2};
  }
  const suta_engine::TaLayer& t = e->ta[l];
  return {e->Pb + e->wqkv_off[l], t.wqkv_t, e->Pb + e->wo_off[l], t.wo_t, e->Pb + e->w1_off[l], t.w1_t, e->Pb + e->w2_off[l], t.w2_t,
          e->P + e->bqkv_off[l], e->P + e->bo_off[l], e->P + e->b1_off[l], e->P + e->b2_off[l]};
}
inline const bf16* lm_w(const suta_engine* e) { return e->train_all ? e->Pb + e->lm_w_off : B16(e->w.lm_w); }
inline const bf16* lm_w_t(const suta_engine* e) { return e->train_all ? e->lm_w_t_sh : B16(e->w.lm_w_t); }
inline const float* lm_b(const suta_engine* e) { return e->train_all ? e->P + e->lm_b_off : e->w.lm_b; }
inline const bf16* pos_w(const suta_engine* e) { return e->train_all ? e->pos_w_sh : B16(e->w.pos_w); }
inline const bf16* pos_w_t(const suta_engine* e) { return e->train_all ? e->pos_w_t_sh : B16(e->w.pos_w_t); }
inline const float* pos_b(const suta_engine* e) { return e->train_all ? e->P + e->pos_b_off : e->w.pos_b; }
// bf16 GEMM inputs the weight gradients of train_all need again in the backward: kept per layer instead of in the shared
// b16 / gelu16 scratch.  act_xa(l): input of layer l's q/k/v projections (l == layers: of lm_head); act_xf: of FFN1; act_gl: of FFN2
inline bf16* act_xa(const suta_engine* e, int l) { return !e->train_all ? e->b16 : (l < e->cfg.layers ? e->ta[l].xa : e->x_last16); }
inline bf16* act_xf(const suta_engine* e, int l) { return e->train_all ? e->ta[l].xf : e->b16; }
inline bf16* act_gl(const suta_engine* e, int l) { return e->train_all ? e->ta[l].gl : e->gelu16; }

// CUDA-event pair around one launch (or group of launches) when profiling is on; records (tag, flops) for the report
struct ProfScope {
  suta_engine* e;
  cudaStream_t st;
  bool on;
  ProfScope(suta_engine* e_, cudaStream_t st_, const std::string& tag, double flops, bool is_gemm, double bytes = 0.0)
      : e(e_), st(st_), on(e_->profile) {
    if (!on) return;
    if (e->ev_used + 2 > e->ev_pool.size()) {
      size_t old = e->ev_pool.size();
      e->ev_pool.resize(old + 1024);
      for (size_t i = old; i < e->ev_pool.size(); ++i) cudaEventCreate(&e->ev_pool[i]);
    }
    e->prof_recs.push_back({tag, flops, is_gemm, bytes});
    cudaEventRecord(e->ev_pool[e->ev_used], st);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(e->ev_pool[e->ev_used + 1], st);
    e->ev_used += 2;
  }
};
#define PROF_F(tag, flops, expr)                \
  do {                                          \
    ProfScope _ps(e, st, tag, flops, false);    \
    SUTA_TRY(expr);                             \
  } while (0)
#define PROF(tag, expr) PROF_F(tag, 0.0, expr)
// HBM-bound kernels: ALGORITHMIC bytes of the launch (what must cross HBM once)
#define PROF_B(tag, bytes, expr)                         \
  do {                                                   \
    ProfScope _ps(e, st, tag, 0.0, false, (double)(bytes)); \
    SUTA_TRY(expr);                                      \
  } while (0)

// ALGORITHMIC bytes of one GEMM launch: every operand element read once, every output element written once (an
// accumulating fp32 output is read and written), saved / consumed auxiliary tensors included
double gemm_algorithmic_bytes(const GemmProblem& p) {
  const double M = p.M, N = p.N, K = p.K > 0 ? p.K : 0, nz = p.nz > 0 ? p.nz : 1;
  double b = 0.0;
  if (p.ztab) {                     // per-utterance reductions over time (weight gradients): flops = 2 M N sum_z K_z
    const double ksum = p.flops > 0 ? p.flops / (2.0 * M * N) : 0.0;
    b += (M + N) * ksum * 2.0 + M * N * 4.0 * nz;
    return b;
  }
  b += M * K * 2.0 * (p.a_z_rows ? nz : 1.0);
  b += N * K * 2.0 * nz;
  const double out = M * N * (p.c_z_cols ? nz : 1.0);
  if (p.epi.out_f32) b += out * (p.epi.accumulate ? 8.0 : 4.0);
  if (p.epi.out_bf16) b += out * 2.0;
  if (p.epi.aux_out) b += out * 2.0;
  if (p.epi.aux_in) b += out * 2.0;
  if (p.epi.residual) b += out * 4.0;
  return b;
}

int gemm(suta_engine* e, const GemmProblem& p, cudaStream_t st) {
  e->launches += 1;
  // SUTA_GEMM_LOG=<file>: one line per tcgen05 GEMM launch, in launch order (joined with an ncu capture by
  // tools/ncu_traffic.py to compare DRAM traffic with the algorithmic bytes)
  static FILE* glog = getenv("SUTA_GEMM_LOG") ? fopen(getenv("SUTA_GEMM_LOG"), "w") : nullptr;
  static long long gidx = 0;
  if (glog) {
    fprintf(glog, "%lld\t%d\t%d\t%d\t%d\t%.0f\t%.0f\n", gidx++, p.M, p.N, p.K, p.nz, gemm_algorithmic_bytes(p),
            p.flops > 0.0 ? p.flops : 2.0 * (double)p.M * p.N * p.K * p.nz);
    fflush(glog);
  }
  if (!e->profile) return gemm_bf16_tc(p, st);
  // algorithmic FLOPs of this launch: valid rows only (the M-block table may pad), all z slices
  const double flops = p.flops > 0.0 ? p.flops : 2.0 * (double)p.M * p.N * p.K * p.nz;
  char tag[96];
  snprintf(tag, sizeof(tag), "gemm N%d K%d%s%s%s%s%s%s", p.N, p.K, p.nz > 1 ? " batched" : "", p.a.mn_major ? " wgrad" : "",
           (!p.a.mn_major && p.b.mn_major) ? " Bmn" : "", p.epi.act == 1 ? " gelu" : (p.epi.act == 2 ? " gelu'" : ""),
           (p.epi.residual || p.epi.accumulate) ? " +res" : "", p.epi.bias ? " +bias" : "");
  ProfScope ps(e, st, tag, flops, true);
  return gemm_bf16_tc(p, st);
}

GemmProblem dense(const bf16* A, long long M, int K, const bf16* B, int N) {
  GemmProblem p;
  p.a = {A, M, K};
  p.b = {B, N, K};
  p.M = (int)M; p.N = N; p.K = K;
  return p;
}

// train_all: gradient of one Linear of the encoder, written whole into the gradient vector (U == 1):
//   d W [n_out, n_in] = d Y^T X   (both operands MN-major, the reduction runs over the utterance's frames; rows past the
//                                  last frame are out of bounds of the tensor maps and read as zeros)
//   d b [n_out]       = column sums of d Y   (bf16 rows, fixed order)
int linear_param_grads(suta_engine* e, const bf16* dY, int n_out, const bf16* X, int n_in, long long w_off, long long b_off,
                       cudaStream_t st) {
  GemmProblem p;
  p.a = {dY, e->M, n_out, 1, n_out};
  p.b = {X, e->M, n_in, 1, n_in};
  p.M = n_out; p.N = n_in; p.K = 0; p.nz = 1;
  p.ztab = e->d_ztab[e->cfg.n_conv];               // U == 1: {0, 0, T}
  p.epi.out_f32 = e->G + w_off; p.epi.out_ld = n_in; p.out_z_stride = e->n_params;
  p.flops = 2.0 * n_out * n_in * (double)e->M;
  SUTA_TRY(gemm(e, p, st));
  PROF("colsum", colsum_per_utt_bf16(dY, e->d_tok_off, e->d_T, e->G, e->n_params, b_off, n_out, e->U, st));
  e->launches += 1;
  return SUTA_OK;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" int suta_engine_create(const suta_model_cfg* cfg, int flags, suta_engine** out) {
  SUTA_CHECK_ARG(cfg && out && (flags & ~(SUTA_FLAG_TRAIN_FEATURE | SUTA_FLAG_PSEUDO_LABEL | SUTA_FLAG_TRAIN_ALL)) == 0);
  SUTA_TRY(check_cfg(*cfg));
  suta_engine* e = new suta_engine();
  e->cfg = *cfg;
  e->train_all = (flags & SUTA_FLAG_TRAIN_ALL) ? 1 : 0;
  e->train_feature = (flags & (SUTA_FLAG_TRAIN_FEATURE | SUTA_FLAG_TRAIN_ALL)) ? 1 : 0;
  e->pseudo_label = (flags & SUTA_FLAG_PSEUDO_LABEL) ? 1 : 0;
  e->conv_ln = cfg->feat_norm_layer ? 1 : 0;
  e->stable = cfg->stable_layer_norm ? 1 : 0;
  e->cnn_bwd = e->train_feature || e->conv_ln;
  if (e->train_all && e->conv_ln != e->stable) {         // the two families HF ships: GroupNorm + post-LN, LayerNorm conv + pre-LN
    suta_set_last_error("SUTA_FLAG_TRAIN_ALL: feat_norm_layer and stable_layer_norm must be set together");
    delete e;
    return SUTA_ERR_ARG;
  }
  if (e->conv_ln)
    for (int l = 1; l < cfg->n_conv; ++l)
      if (!dgrad_fused(e, l)) {
        suta_set_last_error("LayerNorm feature extractor: conv layer %d needs stride 2 and kernel 2 or 3 (parity-split dgrad)", l);
        delete e;
        return SUTA_ERR_ARG;
      }
  build_layout(e);
  *out = e;
  return SUTA_OK;
}
extern "C" void suta_engine_destroy(suta_engine* e) {
  if (!e) return;
  drop_graphs(e);
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
  if (e->stage_ev) cudaEventDestroy(e->stage_ev);
  if (e->h_stage) cudaFreeHost(e->h_stage);
  delete e;
}
extern "C" int64_t suta_engine_param_count(const suta_engine* e) { return e ? e->n_params : 0; }
extern "C" int suta_engine_param_layout(const suta_engine* e, suta_param_seg* segs, int max_segs, int* n_segs) {
  SUTA_CHECK_ARG(e && n_segs);
  *n_segs = (int)e->segs.size();
  if (segs) {
    SUTA_CHECK_ARG(max_segs >= (int)e->segs.size());
    for (size_t i = 0; i < e->segs.size(); ++i)
      segs[i] = {e->segs[i].kind, e->segs[i].module, e->segs[i].index, e->segs[i].off, e->segs[i].size};
  }
  return SUTA_OK;
}
extern "C" int suta_engine_set_weights(suta_engine* e, const suta_weights* w) {
  SUTA_CHECK_ARG(e && w);
  const suta_model_cfg& c = e->cfg;
  SUTA_CHECK_ARG(w->conv0_w && (e->conv_ln || (w->gn_g && w->gn_b)) && w->proj_w && w->proj_w_t && w->proj_b);
  SUTA_CHECK_ARG(w->pos_w && w->pos_w_t && w->pos_b && w->lm_w && w->lm_w_t && w->lm_b && w->params0 && w->mult);
  for (int l = 1; l < c.n_conv; ++l) SUTA_CHECK_ARG(w->conv_w[l]);
  for (int l = 0; l < c.layers; ++l) {
    const suta_layer_weights& x = w->layer[l];
    SUTA_CHECK_ARG(x.wqkv && x.wqkv_t && x.wo && x.wo_t && x.w1 && x.w1_t && x.w2 && x.w2_t && x.bqkv && x.bo && x.b1 && x.b2);
  }
  e->w = *w;
  e->have_weights = true;
  return SUTA_OK;
}

extern "C" int64_t suta_batch_workspace_bytes(suta_engine* e, int n_utts, const int32_t* n_samples) {
  if (!e) return -1;
  drop_graphs(e);                 // planning invalidates the live batch
  if (plan_batch(e, n_utts, n_samples) != SUTA_OK) return -1;
  Bump b;
  carve(e, b);
  e->U = 0;   // planning only: the batch is not live until suta_batch_begin
  return (int64_t)b.off + 256;
}

extern "C" int suta_batch_begin(suta_engine* e, int n_utts, const int32_t* n_samples, void* workspace, int64_t bytes,
                                void* stream) {
  SUTA_CHECK_ARG(e && workspace && e->have_weights);
  if (e->train_all && n_utts != 1) {
    suta_set_last_error("SUTA_FLAG_TRAIN_ALL adapts one utterance per batch (got %d): every weight is the utterance's own", n_utts);
    return SUTA_ERR_ARG;
  }
  drop_graphs(e);                 // recorded chains point into the previous batch's layout
  SUTA_TRY(plan_batch(e, n_utts, n_samples));
  Bump b;
  b.base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(workspace), 256));
  carve(e, b);
  if ((int64_t)(b.off + (b.base - reinterpret_cast<uint8_t*>(workspace))) > bytes) {
    suta_set_last_error("workspace too small: need %zu bytes, got %lld", b.off + 256, (long long)bytes);
    e->U = 0;
    return SUTA_ERR_NOMEM;
  }
  const suta_model_cfg& c = e->cfg;
  const int U = e->U;
  cudaStream_t st = S(stream);
  // ---- host tables -> pinned mirror of the table region -> device, one asynchronous copy ----
  if (e->stage_ev) CUDA_TRY(cudaEventSynchronize(e->stage_ev));      // the previous batch's upload has left the mirror
  else CUDA_TRY(cudaEventCreateWithFlags(&e->stage_ev, cudaEventDisableTiming));
  if (e->h_stage_cap < e->tables_bytes) {
    if (e->h_stage) cudaFreeHost(e->h_stage);
    e->h_stage_cap = e->tables_bytes * 2;
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&e->h_stage), e->h_stage_cap, cudaHostAllocDefault));
  }
  auto up = [&](void* dst, const void* src, size_t n) {
    memcpy(e->h_stage + (reinterpret_cast<uint8_t*>(dst) - b.base), src, n);
  };
  std::vector<int> row_utt(e->M);
  for (int u = 0; u < U; ++u)
    for (int t = 0; t < e->T[u]; ++t) row_utt[e->tok_off[u] + t] = u;
  std::vector<int4> attn_tab;
  attn_tab.reserve(e->n_attn_blk);
  for (int u = 0; u < U; ++u)
    for (int m0 = 0; m0 < e->T[u]; m0 += 128) attn_tab.push_back(make_int4((int)e->tok_off[u], e->T[u], m0, 0));
  up(e->d_samp_off, e->samp_off.data(), sizeof(long long) * U);
  up(e->d_tok_off, e->tok_off.data(), sizeof(long long) * U);
  up(e->d_pad_off, e->pad_off.data(), sizeof(long long) * U);
  up(e->d_off0, e->off[0].data(), sizeof(long long) * U);
  up(e->d_n_samples, e->n_samples.data(), sizeof(int) * U);
  up(e->d_T, e->T.data(), sizeof(int) * U);
  up(e->d_L0, e->L[0].data(), sizeof(int) * U);
  up(e->d_row_utt, row_utt.data(), sizeof(int) * e->M);
  up(e->d_attn_tab, attn_tab.data(), sizeof(int4) * attn_tab.size());
  for (int l = 1; l < c.n_conv; ++l) {
    std::vector<int4> tab;
    tab.reserve(e->n_mblk[l]);
    const int s = c.conv_stride[l];
    for (int u = 0; u < U; ++u)
      for (int m0 = 0; m0 < e->L[l][u]; m0 += 128) {
        int rows = e->L[l][u] - m0 < 128 ? e->L[l][u] - m0 : 128;
        // b_row_off selects the utterance's own weights under train_feature (stacked [U][Cout] rows)
        tab.push_back(make_int4((int)(e->off[l - 1][u] / s + m0), (int)(e->off[l][u] + m0), rows,
                                e->train_feature ? u * c.conv_dim[l] : 0));
      }
    up(e->d_mblk[l], tab.data(), sizeof(int4) * tab.size());
    if (e->cnn_bwd) {                      // 256-row tiles for the CTA-pair kernel: utterance regions are 256-row aligned
      std::vector<int4> ptab;
      for (int u = 0; u < U; ++u)
        for (int m0 = 0; m0 < e->L[l][u]; m0 += 256)
          ptab.push_back(make_int4((int)(e->off[l - 1][u] / s + m0), (int)(e->off[l][u] + m0),
                                   e->L[l][u] - m0 < 256 ? e->L[l][u] - m0 : 256, e->train_feature ? u * c.conv_dim[l] : 0));
      up(e->d_mpair[l], ptab.data(), sizeof(int4) * ptab.size());
    }
  }
  if (e->cnn_bwd) {
    const int last = c.n_conv - 1;
    if (e->train_feature) {
      std::vector<int4> tab;
      for (int u = 0; u < U; ++u)
        for (int m0 = 0; m0 < e->T[u]; m0 += 128)
          tab.push_back(make_int4((int)(e->tok_off[u] + m0), (int)(e->tok_off[u] + m0), e->T[u] - m0 < 128 ? e->T[u] - m0 : 128,
                                  u * c.hidden));
      up(e->d_tok_mblk, tab.data(), sizeof(int4) * tab.size());
    }
    up(e->d_dpre_off_last, e->off64.data(), sizeof(long long) * U);
    for (int l = 0; l < c.n_conv; ++l) {
      up(e->d_off[l], e->off[l].data(), sizeof(long long) * U);
      up(e->d_L[l], e->L[l].data(), sizeof(int) * U);
      if (l >= 1) {
        // rows of d(pre-activation) of layer l: the layer's own (256-row-aligned) layout, except the last layer (token slab off64)
        std::vector<int4> dg, zt;
        for (int u = 0; u < U; ++u) {
          const long long dro = l == last ? e->off64[u] : e->off[l][u];
          const bool fused = dgrad_fused(e, l);
          // fused: one extra (zero) gradient row j = L so that the last input rows get their zeros written; the
          // output row index counts PAIRS of layer l-1 rows
          const int Lx = e->L[l][u] + (fused && l != last ? 1 : 0);
          for (int m0 = 0; m0 < Lx; m0 += 128)
            dg.push_back(make_int4((int)(dro + m0), fused ? (int)(e->off[l - 1][u] / 2 + m0) : (int)(dro + m0),
                                   Lx - m0 < 128 ? Lx - m0 : 128, e->train_feature ? u * c.conv_dim[l] : 0));
          zt.push_back(make_int4((int)dro, (int)(e->off[l - 1][u] / c.conv_stride[l]), e->L[l][u], 0));
        }
        up(e->d_dgrad_mblk[l], dg.data(), sizeof(int4) * dg.size());
        if (e->train_feature) up(e->d_ztab[l], zt.data(), sizeof(int4) * zt.size());
      }
      // gap rows between utterances must be exact zeros: they are reduced over by the weight-gradient GEMMs
      const long long rows = l == last ? e->R64 : e->rows_total[l];
      CUDA_TRY(cudaMemsetAsync(e->conv_out[l], 0, sizeof(bf16) * (size_t)(e->rows_total[l] + 128) * c.conv_dim[l], st));
      CUDA_TRY(cudaMemsetAsync(e->conv_dpre[l] - (size_t)128 * c.conv_dim[l], 0, sizeof(bf16) * (size_t)(rows + 256) * c.conv_dim[l], st));
      // GELU' of rows no forward tile covers is multiplied with zero gradients by the fused dgrad: must be finite
      if (e->train_feature && !e->conv_ln)
        CUDA_TRY(cudaMemsetAsync(e->conv_pre[l], 0, sizeof(bf16) * (size_t)(e->rows_total[l] + 128) * c.conv_dim[l], st));
      // conv_ln: rows of the pre-LayerNorm buffer that no conv tile writes are still streamed through the LayerNorm
      // kernels' shared-memory ring (and skipped): keep them finite
      if (e->conv_ln)
        CUDA_TRY(cudaMemsetAsync(e->conv_z[l], 0, sizeof(bf16) * (size_t)(e->rows_total[l] + 128) * c.conv_dim[l], st));
    }
    if (e->train_feature) {
      std::vector<int4> zt;
      for (int u = 0; u < U; ++u) zt.push_back(make_int4((int)e->off64[u], (int)e->tok_off[u], e->T[u], 0));
      up(e->d_ztab[c.n_conv], zt.data(), sizeof(int4) * zt.size());
      CUDA_TRY(cudaMemsetAsync(e->dh0_pad, 0, sizeof(bf16) * (size_t)(e->R64 + 128) * c.hidden, st));
    }
  }
  if (e->pseudo_label) up(e->d_alpha_off, e->alpha_off.data(), sizeof(long long) * U);
  CUDA_TRY(cudaMemcpyAsync(b.base, e->h_stage, e->tables_bytes, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaEventRecord(e->stage_ev, st));
  if (e->conv_ln)                      // row -> utterance tables of the conv layers' row spaces, built on the device
    for (int l = 0; l < c.n_conv - 1; ++l)
      SUTA_TRY(fill_row_utt(e->d_conv_row_utt[l], e->rows_total[l] + 128, e->d_off[l], e->d_L[l], U, e->max_L[l], st));
  // zero rows of the padded positional-conv slabs never get written afterwards
  CUDA_TRY(cudaMemsetAsync(e->xg, 0, sizeof(bf16) * (size_t)(e->R + 8) * c.hidden, st));
  if (e->train_all) CUDA_TRY(cudaMemsetAsync(e->xg_grad, 0, sizeof(bf16) * (size_t)(e->R + 8) * c.hidden, st));
  // conv buffers carry 128 slack rows read (never used) by partial implicit-GEMM tiles
  for (int l = 0; l < c.n_conv; ++l)
    CUDA_TRY(cudaMemsetAsync(e->conv_out[l] + (size_t)e->rows_total[l] * c.conv_dim[l], 0, sizeof(bf16) * 128 * c.conv_dim[l], st));
  e->frontend_done = false;
  e->moments_done = false; e->z0_done = false;
  e->opt_steps = 0;
  return SUTA_OK;
}

extern "C" int suta_batch_info(const suta_engine* e, int64_t* total_frames, int32_t* frames, int64_t* frame_off,
                               int64_t* sample_off, int64_t* total_samples) {
  SUTA_CHECK_ARG(e && e->U > 0);
  if (total_frames) *total_frames = e->M;
  if (total_samples) *total_samples = e->S;
  for (int u = 0; u < e->U; ++u) {
    if (frames) frames[u] = e->T[u];
    if (frame_off) frame_off[u] = e->tok_off[u];
    if (sample_off) sample_off[u] = e->samp_off[u];
  }
  return SUTA_OK;
}

extern "C" int suta_batch_set_audio(suta_engine* e, const float* wav, int flags, void* stream) {
  SUTA_CHECK_ARG(e && e->U > 0 && wav);
  const bool is_host = flags & 1;
  e->audio_normalized = (flags & 2) != 0;
  CUDA_TRY(cudaMemcpyAsync(e->audio_normalized ? e->wav_norm : e->wav, wav, sizeof(float) * e->S,
                           is_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, S(stream)));
  e->frontend_done = false;
  e->moments_done = false; e->z0_done = false;
  return SUTA_OK;
}

extern "C" int suta_profile(suta_engine* e, int enable, double* gemm_ms, int64_t* gemm_launches, double* gemm_flops) {
  SUTA_CHECK_ARG(e);
  if (gemm_ms || gemm_launches || gemm_flops) {
    CUDA_TRY(cudaDeviceSynchronize());
    double ms = 0.0, fl = 0.0;
    int64_t n = 0;
    e->prof_report.clear();
    std::vector<std::string> tags;
    std::vector<double> tms, tfl, tby;
    std::vector<long long> tn;
    for (size_t i = 0; i < e->prof_recs.size() && 2 * i + 1 < e->ev_used; ++i) {
      float t = 0.f;
      CUDA_TRY(cudaEventElapsedTime(&t, e->ev_pool[2 * i], e->ev_pool[2 * i + 1]));
      const auto& r = e->prof_recs[i];
      if (r.is_gemm) { ms += t; fl += r.flops; n += 1; }
      size_t k = 0;
      while (k < tags.size() && tags[k] != r.tag) ++k;
      if (k == tags.size()) { tags.push_back(r.tag); tms.push_back(0); tfl.push_back(0); tby.push_back(0); tn.push_back(0); }
      tms[k] += t; tfl[k] += r.flops; tby[k] += r.bytes; tn[k] += 1;
    }
    for (size_t k = 0; k < tags.size(); ++k) {
      char line[256];
      snprintf(line, sizeof(line), "%s\t%.4f\t%.6g\t%lld\t%.6g\n", tags[k].c_str(), tms[k], tfl[k], tn[k], tby[k]);
      e->prof_report += line;
    }
    if (gemm_ms) *gemm_ms = ms;
    if (gemm_launches) *gemm_launches = n;
    if (gemm_flops) *gemm_flops = fl;
  }
  e->ev_used = 0;
  e->prof_recs.clear();
  e->profile = enable != 0;
  return SUTA_OK;
}
// per-tag breakdown of the last suta_profile read-out: lines "tag<TAB>ms<TAB>flops<TAB>launches"
extern "C" const char* suta_profile_report(const suta_engine* e) { return e ? e->prof_report.c_str() : ""; }

// bf16 GEMM-operand copies of the per-utterance trainable matrices (train_feature): refreshed after every update
// train_all: everything derived from the trainable vector beyond its plain bf16 copy -- W^T of every encoder Linear and of
// lm_head (B operands of the dgrads) and the folded weight_norm weight of the positional conv in its two layouts
static int refresh_train_all(suta_engine* e, cudaStream_t st) {
  const suta_model_cfg& c = e->cfg;
  const int H = c.hidden, I = c.intermediate;
  TransposeJobs jobs;                         // <= 100 per launch: 24 layers x 4 + lm_head = 97
  auto flush = [&]() -> int {
    if (jobs.n == 0) return SUTA_OK;
    e->launches += 1;
    PROF("transpose_cast", transpose_cast_bf16(jobs, st));
    jobs.n = 0;
    return SUTA_OK;
  };
  for (int l = 0; l < c.layers; ++l) {
    const suta_engine::TaLayer& t = e->ta[l];
    if (jobs.n + 4 > TransposeJobs::MAX) SUTA_TRY(flush());
    jobs.add(e->P + e->wqkv_off[l], t.wqkv_t, 3 * H, H);
    jobs.add(e->P + e->wo_off[l], t.wo_t, H, H);
    jobs.add(e->P + e->w1_off[l], t.w1_t, I, H);
    jobs.add(e->P + e->w2_off[l], t.w2_t, H, I);
  }
  if (jobs.n + 1 > TransposeJobs::MAX) SUTA_TRY(flush());
  jobs.add(e->P + e->lm_w_off, e->lm_w_t_sh, c.vocab, H);
  SUTA_TRY(flush());
  PROF("weight_norm_fwd", posconv_weight_norm_forward(e->P + e->pos_g_off, e->P + e->pos_v_off, e->wn_scratch, e->pos_w_sh, e->pos_w_t_sh, H,
                                                      H / c.pos_groups, c.pos_k, st));
  e->launches += 3;
  e->frontend_done = false;
  return SUTA_OK;
}

static int refresh_shadows(suta_engine* e, cudaStream_t st) {
  if (!e->train_feature) return SUTA_OK;
  const suta_model_cfg& c = e->cfg;
  if (e->train_all) {            // one cast of the whole vector (the conv / projection operand copies are slices of Pb)
    PROF("cast_shadow", cast_params_bf16(e->P, e->n_params, 0, e->n_params, e->U, e->Pb, st));
    e->launches += 1;
    return refresh_train_all(e, st);
  }
  for (int l = 1; l < c.n_conv; ++l) {
    PROF("cast_shadow", cast_params_bf16(e->P, e->n_params, e->conv_w_off[l], e->conv_w_size[l], e->U, e->w_shadow[l], st));
    e->launches += 1;
  }
  PROF("cast_shadow", cast_params_bf16(e->P, e->n_params, e->proj_w_off, (long long)c.hidden * c.conv_dim[c.n_conv - 1], e->U,
                            e->proj_shadow, st));
  e->launches += 1;
  e->frontend_done = false;      // the CNN output depends on the updated weights
  return SUTA_OK;
}

// REF/data.py:23 on the device: adds sigma * N(0,1) to the RAW waveform the last suta_batch_set_audio copied in (before
// the normalisation of suta_frontend).  utt_ids (host, optional): a stable id per utterance, so that an utterance gets
// the same noise whatever batch it is adapted in; default = position in the batch.
extern "C" int suta_batch_add_noise(suta_engine* e, float sigma, uint64_t seed, const int32_t* utt_ids, void* stream) {
  SUTA_CHECK_ARG(e && e->U > 0 && sigma >= 0.f);
  if (e->audio_normalized) {
    suta_set_last_error("suta_batch_add_noise: the audio was handed over already normalised; noise goes on the raw waveform");
    return SUTA_ERR_ARG;
  }
  cudaStream_t st = S(stream);
  int* d_ids = nullptr;
  if (utt_ids) {                   // ids ride in the (otherwise unused until decode) argmax-id buffer
    d_ids = e->ids;
    CUDA_TRY(cudaMemcpyAsync(d_ids, utt_ids, sizeof(int) * e->U, cudaMemcpyHostToDevice, st));
  }
  e->launches += 1;
  PROF("audio_noise", audio_add_noise(e->wav, e->d_samp_off, e->d_n_samples, d_ids, e->U, e->max_samples, sigma, seed, st));
  e->frontend_done = false;
  e->moments_done = false; e->z0_done = false;
  return SUTA_OK;
}

extern "C" int suta_reset(suta_engine* e, void* stream) {
  SUTA_CHECK_ARG(e && e->U > 0);
  e->launches += 1;
  e->opt_steps = 0;
  cudaStream_t st = S(stream);
  PROF("reset", params_reset(e->P, e->w.params0, e->Mom, e->Var, nullptr, e->n_params, e->U, st));
  if (e->conv_ln) e->frontend_done = false;        // the CNN output depends on the (trainable) conv LayerNorms
  return refresh_shadows(e, S(stream));
}

// The caller wrote the trainable vectors directly (model.load_state_dict, carrying a continual model into a new batch):
// refresh everything derived from them.
extern "C" int suta_params_written(suta_engine* e, void* stream) {
  SUTA_CHECK_ARG(e && e->U > 0);
  e->frontend_done = false;
  return refresh_shadows(e, S(stream));
}

// LayerNorm feature extractor (feat_extract_norm == "layer", HF/modeling_wav2vec2.py:275-299): every layer is
// Conv1d(+bias, frozen) -> LayerNorm(C) with the utterance's own gamma/beta (trainable, REF/main.py:81-87) -> GELU.  The
// pre-LayerNorm value z_l and the row statistics are kept for the backward; runs on every forward (the LayerNorms move).
static int frontend_layer_norm(suta_engine* e, cudaStream_t st) {
  const suta_model_cfg& c = e->cfg;
  UttParams prm{e->P, e->n_params};
  const int last = c.n_conv - 1;
  const bool tf = e->train_feature != 0;
  if (tf) {                // per-utterance taps and bias inside the trainable vector: every forward
    PROF_B("conv0_bias", (double)e->S * 4 + (double)e->rows_total[0] * c.conv_dim[0] * 2,
           conv0_bias(e->wav_norm, e->d_samp_off, e->d_L0, e->d_off0, e->P + e->conv_w_off[0], e->P + e->conv_b_off[0], e->conv_z[0], e->U,
                      c.conv_dim[0], c.conv_kernel[0], c.conv_stride[0], e->max_L0, st, e->n_params, e->n_params));
    e->launches += 1;
  } else if (!e->z0_done) {  // conv0 and its bias are frozen: z_0 depends on the audio only, once per batch (only its LayerNorm moves)
    PROF_B("conv0_bias", (double)e->S * 4 + (double)e->rows_total[0] * c.conv_dim[0] * 2,
           conv0_bias(e->wav_norm, e->d_samp_off, e->d_L0, e->d_off0, e->w.conv0_w, e->w.conv_b[0], e->conv_z[0], e->U, c.conv_dim[0],
                      c.conv_kernel[0], c.conv_stride[0], e->max_L0, st));
    e->launches += 1;
    e->z0_done = true;
  }
  for (int l = 0; l < c.n_conv; ++l) {
    const int Cout = c.conv_dim[l];
    long long rows_valid = 0;
    for (int u = 0; u < e->U; ++u) rows_valid += e->L[l][u];
    if (l >= 1) {
      const int Cin = c.conv_dim[l - 1], k = c.conv_kernel[l], s = c.conv_stride[l];
      GemmProblem p;
      const long long rows_in = e->rows_total[l - 1] + 128;      // overlapping-row view, as in suta_frontend
      p.a = {e->conv_out[l - 1], (rows_in - k) / s + 1, (long long)s * Cin};
      p.b = {reinterpret_cast<const bf16*>(e->w.conv_w[l]), Cout, (long long)k * Cin};
      p.M = (int)rows_valid; p.N = Cout; p.K = k * Cin;
      p.mblk = e->d_mblk[l]; p.num_mblk = e->n_mblk[l];
      p.epi.bias = e->w.conv_b[l];
      p.epi.out_bf16 = e->conv_z[l]; p.epi.out_ld = Cout;
      if (tf) {                  // the utterance's own weights (stacked bf16 copies) and bias: row-masked epilogue, the gap
        p.b = {e->w_shadow[l], (long long)e->U * Cout, (long long)k * Cin};       // rows of z_l keep their zeros
        p.epi.bias = e->P + e->conv_b_off[l]; p.epi.bias_utt_stride = e->n_params;
      } else if (l < last) {     // 256-row-aligned utterances: tiles own their rows (the last layer packs tokens densely)
        p.tiles_own_rows = 1;
        p.out_rows = e->rows_total[l] + 128;
        p.mpair = e->d_mpair[l]; p.num_mpair = e->n_mpair[l];
      }
      SUTA_TRY(gemm(e, p, st));
    }
    const long long rows = l == last ? e->M : e->rows_total[l];
    PROF_B("ln_gelu_fwd C", (double)rows_valid * Cout * (2 + 2),
           layernorm_forward(nullptr, e->conv_z[l], l == last ? e->d_row_utt : e->d_conv_row_utt[l], prm, (int)e->cln_g[l],
                             (int)e->cln_b[l], nullptr, e->conv_out[l], e->conv_mean[l], e->conv_rstd[l], rows, Cout, 1e-5f, st,
                             nullptr, LN_GELU));
    e->launches += 1;
  }
  e->frontend_done = true;
  return SUTA_OK;
}

extern "C" int suta_frontend(suta_engine* e, void* stream) {
  SUTA_CHECK_ARG(e && e->U > 0);
  const suta_model_cfg& c = e->cfg;
  cudaStream_t st = S(stream);
  if (!e->audio_normalized && !e->moments_done)     // once per set_audio
    PROF("normalize", normalize_audio(e->wav, e->wav_norm, e->d_samp_off, e->d_n_samples, e->U, e->max_samples, e->stats, st));
  if (e->conv_ln) {
    if (!e->moments_done) e->launches += 2;
    e->moments_done = true;
    return frontend_layer_norm(e, st);
  }
  if (!e->moments_done) {      // second moments of the conv0 input windows: depend on the audio only, once per batch
    PROF("conv0_moments", audio_conv0_moments(e->wav_norm, e->d_samp_off, e->d_L0, e->mom, c.conv_kernel[0], c.conv_stride[0], e->U,
                                              e->max_L0, st));
    e->moments_done = true;
    e->launches += 2;
  }
  Conv0Args a{};
  a.mom = e->mom;
  a.x = e->wav_norm; a.samp_off = e->d_samp_off; a.L0 = e->d_L0; a.out_off = e->d_off0;
  a.w = e->w.conv0_w; a.w_stride = 0;
  a.gn_shared_g = e->w.gn_g; a.gn_shared_b = e->w.gn_b;
  a.stats = e->stats + 2 * e->U;
  a.out = e->conv_out[0]; a.pre_out = nullptr;
  if (e->train_feature) {        // per-utterance conv0 weight and GroupNorm affine live in the trainable vector
    a.w = e->P + e->conv_w_off[0]; a.w_stride = e->n_params;
    a.gn = UttParams{e->P, e->n_params}; a.g_off = (int)e->gn_g; a.b_off = (int)e->gn_b;
    a.pre_out = e->conv_pre[0];
  }
  a.n_utts = e->U; a.C = c.conv_dim[0]; a.k = c.conv_kernel[0]; a.stride = c.conv_stride[0]; a.max_L0 = e->max_L0;
  PROF_B("conv0_fwd", (double)e->S * 4 + (double)e->rows_total[0] * c.conv_dim[0] * 2 * (e->train_feature ? 2 : 1),
         conv0_groupnorm_gelu(a, st));
  e->launches += 2;
  for (int l = 1; l < c.n_conv; ++l) {
    const int Cin = c.conv_dim[l - 1], Cout = c.conv_dim[l], k = c.conv_kernel[l], s = c.conv_stride[l];
    GemmProblem p;
    // overlapping-row view: row r = input rows [s*r, s*r + k) flattened = k*Cin contiguous elements
    long long rows_in = e->rows_total[l - 1] + 128;
    p.a = {e->conv_out[l - 1], (rows_in - k) / s + 1, (long long)s * Cin};
    p.b = {reinterpret_cast<const bf16*>(e->w.conv_w[l]), Cout, (long long)k * Cin};
    p.M = 0; p.N = Cout; p.K = k * Cin;
    p.mblk = e->d_mblk[l]; p.num_mblk = e->n_mblk[l];
    { long long rows = 0; for (int u = 0; u < e->U; ++u) rows += e->L[l][u]; p.M = (int)rows; }   // valid rows (FLOP accounting); tiles come from the table
    p.epi.act = 1;
    p.epi.out_bf16 = e->conv_out[l]; p.epi.out_ld = Cout;
    if (e->train_feature) {
      p.b = {e->w_shadow[l], (long long)e->U * Cout, (long long)k * Cin};
      p.epi.aux_out = e->conv_pre[l]; p.epi.aux_ld = Cout;
      if (l < c.n_conv - 1) {    // 256-row-aligned utterances: tiles own their rows (the last layer packs tokens densely)
        p.tiles_own_rows = 1;
        p.out_rows = e->rows_total[l] + 128;
        p.mpair = e->d_mpair[l]; p.num_mpair = e->n_mpair[l];
      }
    }
    SUTA_TRY(gemm(e, p, st));
  }
  e->frontend_done = true;
  return SUTA_OK;
}

// Pre-LN encoder (do_stable_layer_norm, HF/modeling_wav2vec2.py:638-645, :790): per layer
//   r' = r + out_proj(attention(LN1(r))),   r'' = r' + FFN(LN2(r')),   logits = lm_head(LN_enc(r_final)).
// Same buffers and the same in-place residual scheme as the post-LN path: lb[l].h1 = r entering layer l (LN1's input, kept
// for the backward), lb[l].h2 = r'; each LayerNorm writes its normalised bf16 output for the GEMM AND seeds the next
// residual buffer with its INPUT plus the bias of the GEMM that then accumulates onto it (LN_KEEP_INPUT).
static int forward_stable(suta_engine* e, cudaStream_t st) {
  const suta_model_cfg& c = e->cfg;
  const int H = c.hidden, I = c.intermediate, V = c.vocab;
  const long long M = e->M;
  UttParams prm{e->P, e->n_params};
  for (int l = 0; l < c.layers; ++l) {
    const LayerW w = layer_w(e, l);
    LayerBufs& x = e->lb[l];
    PROF_B("ln_fwd", (double)M * H * (4 + 4 + 2), layernorm_forward(x.h1, nullptr, e->d_row_utt, prm, (int)e->ln1_g[l], (int)e->ln1_b[l], x.h2, act_xa(e, l),
                               x.mean1, x.rstd1, M, H, c.ln_eps, st, w.bo, LN_KEEP_INPUT));
    {
      GemmProblem p = dense(act_xa(e, l), M, H, reinterpret_cast<const bf16*>(w.wqkv), 3 * H);
      p.epi.bias = w.bqkv; p.epi.out_bf16 = x.qkv; p.epi.out_ld = 3 * H;
      SUTA_TRY(gemm(e, p, st));
    }
    PROF_F("attn_fwd", 4.0 * H * e->sumT2, attention_forward(x.qkv, x.attn, x.lse, e->d_attn_tab, e->n_attn_blk, H, c.heads, M, st));
    {
      GemmProblem p = dense(x.attn, M, H, reinterpret_cast<const bf16*>(w.wo), H);
      p.epi.accumulate = 1; p.epi.out_f32 = x.h2; p.epi.out_ld = H;
      SUTA_TRY(gemm(e, p, st));
    }
    float* next = l + 1 < c.layers ? e->lb[l + 1].h1 : e->hE;        // hE: the residual stream after the last layer
    PROF_B("ln_fwd", (double)M * H * (4 + 4 + 2), layernorm_forward(x.h2, nullptr, e->d_row_utt, prm, (int)e->ln2_g[l], (int)e->ln2_b[l], next, act_xf(e, l),
                               x.mean2, x.rstd2, M, H, c.ln_eps, st, w.b2, LN_KEEP_INPUT));
    {
      GemmProblem p = dense(act_xf(e, l), M, H, reinterpret_cast<const bf16*>(w.w1), I);
      p.epi.bias = w.b1; p.epi.act = 1; p.epi.aux_out = x.pre; p.epi.aux_ld = I; p.epi.out_bf16 = act_gl(e, l); p.epi.out_ld = I;
      SUTA_TRY(gemm(e, p, st));
    }
    {
      GemmProblem p = dense(act_gl(e, l), M, I, reinterpret_cast<const bf16*>(w.w2), H);
      p.epi.accumulate = 1; p.epi.out_f32 = next; p.epi.out_ld = H;
      SUTA_TRY(gemm(e, p, st));
    }
    e->launches += 3;
  }
  // encoder.layer_norm AFTER the layers (HF:790): only the bf16 operand of lm_head is needed
  PROF_B("ln_fwd", (double)M * H * (4 + 2), layernorm_forward(e->hE, nullptr, e->d_row_utt, prm, (int)e->enc_g, (int)e->enc_b, nullptr, act_xa(e, c.layers),
                             e->enc_mean, e->enc_rstd, M, H, c.ln_eps, st));
  e->launches += 1;
  {
    GemmProblem p = dense(act_xa(e, c.layers), M, H, lm_w(e), V);
    p.epi.bias = lm_b(e); p.epi.out_f32 = e->logits; p.epi.out_ld = V;
    SUTA_TRY(gemm(e, p, st));
  }
  return SUTA_OK;
}

static int forward_eager(suta_engine* e, void* stream) {
  SUTA_CHECK_ARG(e && e->U > 0);
  if (!e->frontend_done) SUTA_TRY(suta_frontend(e, stream));
  const suta_model_cfg& c = e->cfg;
  cudaStream_t st = S(stream);
  const int H = c.hidden, I = c.intermediate, C = c.conv_dim[c.n_conv - 1], V = c.vocab, G = c.pos_groups, CG = H / G;
  const long long M = e->M;
  UttParams prm{e->P, e->n_params};
  const bf16* feat = e->conv_out[c.n_conv - 1];

  // feature projection: LayerNorm(C) -> Linear(C->H)        HF/modeling_wav2vec2.py:429-434
  PROF_B("ln_fwd C", (double)M * C * (2 + 2), layernorm_forward(nullptr, feat, e->d_row_utt, prm, (int)e->fp_g, (int)e->fp_b, nullptr, e->y_fp, e->fp_mean,
                             e->fp_rstd, M, C, c.ln_eps, st));
  {
    GemmProblem p = dense(e->y_fp, M, C, reinterpret_cast<const bf16*>(e->w.proj_w), H);
    p.epi.bias = e->w.proj_b; p.epi.out_f32 = e->h0; p.epi.out_ld = H;
    if (e->train_feature) {      // per-utterance projection weight / bias
      p.b = {e->proj_shadow, (long long)e->U * H, C};
      p.mblk = e->d_tok_mblk; p.num_mblk = e->n_tok_mblk;
      p.epi.bias = e->P + e->proj_b_off; p.epi.bias_utt_stride = e->n_params;
    }
    SUTA_TRY(gemm(e, p, st));
  }
  // positional conv embedding + GELU + residual               HF/modeling_wav2vec2.py:360-368, :690-691
  PROF("posconv_pack", posconv_pack(e->h0, e->d_row_utt, e->d_tok_off, e->d_pad_off, e->xg, M, H, G, CG, e->R, st));
  if (use_posconv_tc(e)) {
    PROF("posconv_tc", posconv_tc(e->xg, pos_w(e), pos_b(e), e->cpos, H, G, CG, e->R, e->Rm, c.pos_k, st));
    e->launches += 1;
  } else {
    GemmProblem p;
    p.a = {e->xg, (long long)G * e->R - c.pos_k + 1, CG};
    p.b = {pos_w(e), H, (long long)c.pos_k * CG};
    p.M = (int)e->Rm; p.N = CG; p.K = c.pos_k * CG;
    p.nz = G; p.a_z_rows = e->R; p.b_z_rows = CG; p.c_z_cols = CG;
    p.epi.bias = pos_b(e); p.epi.out_f32 = e->cpos; p.epi.out_ld = H;
    SUTA_TRY(gemm(e, p, st));
  }
  // pre-LN ("stable") encoder: no LayerNorm here -- h0 + pos-conv IS the residual stream entering layer 0 (HF:760-762)
  PROF("posconv_combine", posconv_combine(e->h0, e->cpos, e->d_row_utt, e->d_tok_off, e->d_pad_off, e->stable ? e->lb[0].h1 : e->hE, M, H,
                                          -(c.pos_k / 2), st));
  if (e->stable) {
    e->launches += 4;
    return forward_stable(e, st);
  }
  // encoder.layer_norm                                          HF/modeling_wav2vec2.py:692
  // The residual stream is updated IN PLACE: every LayerNorm writes its fp32 output straight into the buffer that holds
  // the next pre-LayerNorm sum (lb[l].h1 / h2, kept per layer for the backward), and the following GEMM accumulates
  // "+= x W^T + b" into it with a TMA reduce-add -- the GEMM epilogue never loads the residual.
  // (the bias of that GEMM is added by the LayerNorm too, so its fp32 epilogue needs no bias and keeps 4 stages)
  PROF_B("ln_fwd", (double)M * H * (4 + 4 + 2), layernorm_forward(e->hE, nullptr, e->d_row_utt, prm, (int)e->enc_g, (int)e->enc_b, e->lb[0].h1, act_xa(e, 0), e->enc_mean,
                             e->enc_rstd, M, H, c.ln_eps, st, layer_w(e, 0).bo));
  e->launches += 5;
  for (int l = 0; l < c.layers; ++l) {
    const LayerW w = layer_w(e, l);
    LayerBufs& x = e->lb[l];
    {  // q,k,v projections fused to one N=3H GEMM            HF:500-507
      GemmProblem p = dense(act_xa(e, l), M, H, reinterpret_cast<const bf16*>(w.wqkv), 3 * H);
      p.epi.bias = w.bqkv; p.epi.out_bf16 = x.qkv; p.epi.out_ld = 3 * H;
      SUTA_TRY(gemm(e, p, st));
    }
    PROF_F("attn_fwd", 4.0 * H * e->sumT2, attention_forward(x.qkv, x.attn, x.lse, e->d_attn_tab, e->n_attn_blk, H, c.heads, M, st));
    {  // out_proj + residual                                    HF:546, :597
      GemmProblem p = dense(x.attn, M, H, reinterpret_cast<const bf16*>(w.wo), H);
      p.epi.accumulate = 1; p.epi.out_f32 = x.h1; p.epi.out_ld = H;      // + bo: already in h1 (added by the LayerNorm)
      SUTA_TRY(gemm(e, p, st));
    }
    PROF_B("ln_fwd", (double)M * H * (4 + 4 + 2), layernorm_forward(x.h1, nullptr, e->d_row_utt, prm, (int)e->ln1_g[l], (int)e->ln1_b[l], x.h2, act_xf(e, l),
                               x.mean1, x.rstd1, M, H, c.ln_eps, st, w.b2));
    {  // intermediate_dense + GELU (pre-activation kept for the backward)      HF:565-566
      GemmProblem p = dense(act_xf(e, l), M, H, reinterpret_cast<const bf16*>(w.w1), I);
      p.epi.bias = w.b1; p.epi.act = 1; p.epi.aux_out = x.pre; p.epi.aux_ld = I; p.epi.out_bf16 = act_gl(e, l); p.epi.out_ld = I;
      SUTA_TRY(gemm(e, p, st));
    }
    {  // output_dense + residual                                HF:569, :600
      GemmProblem p = dense(act_gl(e, l), M, I, reinterpret_cast<const bf16*>(w.w2), H);
      p.epi.accumulate = 1; p.epi.out_f32 = x.h2; p.epi.out_ld = H;      // + b2: already in h2
      SUTA_TRY(gemm(e, p, st));
    }
    PROF_B("ln_fwd", (double)M * H * (4 + 4 + 2), layernorm_forward(x.h2, nullptr, e->d_row_utt, prm, (int)e->ln2_g[l], (int)e->ln2_b[l],
                               l + 1 < c.layers ? e->lb[l + 1].h1 : e->fa, act_xa(e, l + 1), x.mean2, x.rstd2, M, H, c.ln_eps, st,
                               l + 1 < c.layers ? layer_w(e, l + 1).bo : nullptr));
    e->launches += 3;
  }
  {  // lm_head                                                   HF:1708
    GemmProblem p = dense(act_xa(e, c.layers), M, H, lm_w(e), V);
    p.epi.bias = lm_b(e); p.epi.out_f32 = e->logits; p.epi.out_ld = V;
    SUTA_TRY(gemm(e, p, st));
  }
  return SUTA_OK;
}

static int loss_backward_eager(suta_engine* e, const suta_hyper* h, void* stream) {
  SUTA_CHECK_ARG(e && e->U > 0 && h);
  const suta_model_cfg& c = e->cfg;
  cudaStream_t st = S(stream);
  const int H = c.hidden, I = c.intermediate, C = c.conv_dim[c.n_conv - 1], V = c.vocab, G = c.pos_groups, CG = H / G;
  const long long M = e->M;
  UttParams prm{e->P, e->n_params};

  LossArgs la{};
  la.logits = e->logits; la.tok_off = e->d_tok_off; la.T = e->d_T;
  la.dlogits_f32 = e->dlogits; la.dlogits_bf16 = e->dlogits16; la.loss = e->losses; la.n_utts = e->U;
  la.em_coef = h->em_coef; la.temp = h->temp; la.reweight = h->reweight; la.not_blank = h->not_blank;
  la.div_coef = h->div_coef;
  PROF("loss", suta_loss_forward_backward(la, st));
  e->launches += 1;
  if (h->pl_coef > 0.f) {          // SDPL: CTC loss against the greedy transcript of these logits (REF/main_SDPL.py:176,194-209)
    if (!e->pseudo_label) {
      suta_set_last_error("pl_coef > 0 needs an engine created with SUTA_FLAG_PSEUDO_LABEL (the CTC scratch is part of the workspace)");
      return SUTA_ERR_ARG;
    }
    SUTA_TRY(ctc_greedy_decode(e->logits, e->d_tok_off, e->d_T, e->ids, e->collapsed, e->out_len, e->U, V, st));
    CtcArgs ca{};
    ca.logits = e->logits; ca.tok_off = e->d_tok_off; ca.T = e->d_T; ca.collapsed = e->collapsed; ca.collapsed_len = e->out_len;
    ca.alpha = e->ctc_alpha; ca.alpha_off = e->d_alpha_off; ca.g = e->ctc_g;
    ca.dlogits_f32 = e->dlogits; ca.dlogits_bf16 = e->dlogits16; ca.loss = e->losses; ca.target_len = e->ctc_tlen;
    ca.n_utts = e->U; ca.pl_coef = h->pl_coef;
    int maxT = 0;
    for (int u = 0; u < e->U; ++u) maxT = e->T[u] > maxT ? e->T[u] : maxT;
    ca.max_states = 2 * maxT + 1;
    PROF("ctc_pseudo_label", ctc_pseudo_label_loss(ca, st));
    e->launches += 2;
  }
  // every gradient segment is written whole by exactly one kernel of this backward (no accumulation into G, no atomics)

  // two fp32 gradient streams, updated in place like the forward's residual stream: LayerNorm backward writes d(input)
  // into the other buffer and the following dgrad GEMM accumulates its product onto it (TMA reduce-add)
  LnReduceBatch lnred;                 // second stage of every LayerNorm's dgamma/dbeta reduction, launched once at the end
  auto ln_slot = [&]() { return e->ln_part + e->ln_part_stride * (size_t)lnred.n; };
  float* da = e->fa;   // gradient w.r.t. the current LayerNorm output
  float* db = e->fb;   // gradient w.r.t. the pre-LayerNorm sum
  {  // lm_head dgrad
    GemmProblem p = dense(e->dlogits16, M, V, lm_w_t(e), H);
    p.epi.out_f32 = da; p.epi.out_ld = H;
    SUTA_TRY(gemm(e, p, st));
  }
  if (e->train_all) SUTA_TRY(linear_param_grads(e, e->dlogits16, V, act_xa(e, c.layers), H, e->lm_w_off, e->lm_b_off, st));
  if (e->stable) {
    // Pre-LN encoder (HF:638-645, :790).  db carries d(residual stream); every branch gradient comes back through its
    // LayerNorm's backward, which adds it onto db in place (dx_add = dx = db) and emits the bf16 operand of the next dgrad.
    PROF_B("ln_bwd", (double)M * H * (4 + 4 + 4 + 2), layernorm_backward(da, e->hE, nullptr, e->enc_mean, e->enc_rstd, e->d_row_utt, prm, (int)e->enc_g, (int)e->enc_b,
                                e->G, db, e->b16, M, H, e->d_tok_off, e->d_T, e->U, ln_slot(), st, &lnred.item[lnred.n]));
    lnred.n += 1;
    for (int l = c.layers - 1; l >= 0; --l) {
      const LayerW w = layer_w(e, l);
      LayerBufs& x = e->lb[l];
      // (train_all: b16 = bf16 of d(residual stream) = the gradient of output_dense's result; after the LayerNorm backward
      //  below, of out_proj's)
      if (e->train_all) SUTA_TRY(linear_param_grads(e, e->b16, H, act_gl(e, l), I, e->w2_off[l], e->b2_off[l], st));
      {  // output_dense dgrad, times GELU'(pre)
        GemmProblem p = dense(e->b16, M, H, reinterpret_cast<const bf16*>(w.w2_t), I);
        p.epi.act = 2; p.epi.aux_in = x.pre; p.epi.aux_ld = I; p.epi.out_bf16 = e->dpre16; p.epi.out_ld = I;
        SUTA_TRY(gemm(e, p, st));
      }
      {  // intermediate_dense dgrad -> d(final_layer_norm output)
        GemmProblem p = dense(e->dpre16, M, I, reinterpret_cast<const bf16*>(w.w1_t), H);
        p.epi.out_f32 = da; p.epi.out_ld = H;
        SUTA_TRY(gemm(e, p, st));
      }
      if (e->train_all) SUTA_TRY(linear_param_grads(e, e->dpre16, I, act_xf(e, l), H, e->w1_off[l], e->b1_off[l], st));
      PROF_B("ln_bwd", (double)M * H * (4 + 4 + 4 + 4 + 2), layernorm_backward(da, x.h2, nullptr, x.mean2, x.rstd2, e->d_row_utt, prm, (int)e->ln2_g[l], (int)e->ln2_b[l],
                                  e->G, db, e->b16, M, H, e->d_tok_off, e->d_T, e->U, ln_slot(), st, &lnred.item[lnred.n], db));
      lnred.n += 1;
      if (e->train_all) SUTA_TRY(linear_param_grads(e, e->b16, H, x.attn, H, e->wo_off[l], e->bo_off[l], st));
      {  // out_proj dgrad
        GemmProblem p = dense(e->b16, M, H, reinterpret_cast<const bf16*>(w.wo_t), H);
        p.epi.out_bf16 = e->dO16; p.epi.out_ld = H;
        SUTA_TRY(gemm(e, p, st));
      }
      PROF_F("attn_bwd", 8.0 * H * e->sumT2, attention_backward(x.qkv, x.attn, e->dO16, x.lse, e->Dbuf, e->dqkv16, e->d_attn_tab, e->n_attn_blk, H, c.heads, M, st));
      {  // q,k,v dgrad -> d(layer_norm output)
        GemmProblem p = dense(e->dqkv16, M, 3 * H, reinterpret_cast<const bf16*>(w.wqkv_t), H);
        p.epi.out_f32 = da; p.epi.out_ld = H;
        SUTA_TRY(gemm(e, p, st));
      }
      if (e->train_all) SUTA_TRY(linear_param_grads(e, e->dqkv16, 3 * H, act_xa(e, l), H, e->wqkv_off[l], e->bqkv_off[l], st));
      PROF_B("ln_bwd", (double)M * H * (4 + 4 + 4 + 4 + 2), layernorm_backward(da, x.h1, nullptr, x.mean1, x.rstd1, e->d_row_utt, prm, (int)e->ln1_g[l], (int)e->ln1_b[l],
                                  e->G, db, l > 0 ? e->b16 : nullptr, M, H, e->d_tok_off, e->d_T, e->U, ln_slot(), st, &lnred.item[lnred.n], db));
      lnred.n += 1;
      e->launches += 5;
    }
  }
  for (int l = e->stable ? -1 : c.layers - 1; l >= 0; --l) {
    const LayerW w = layer_w(e, l);
    LayerBufs& x = e->lb[l];
    PROF_B("ln_bwd", (double)M * H * (4 + 4 + 4 + 2), layernorm_backward(da, x.h2, nullptr, x.mean2, x.rstd2, e->d_row_utt, prm, (int)e->ln2_g[l], (int)e->ln2_b[l],
                                e->G, db, e->b16, M, H, e->d_tok_off, e->d_T, e->U, ln_slot(), st, &lnred.item[lnred.n]));
    lnred.n += 1;
    if (e->train_all) SUTA_TRY(linear_param_grads(e, e->b16, H, act_gl(e, l), I, e->w2_off[l], e->b2_off[l], st));      // output_dense
    {  // output_dense dgrad, times GELU'(pre)
      GemmProblem p = dense(e->b16, M, H, reinterpret_cast<const bf16*>(w.w2_t), I);
      p.epi.act = 2; p.epi.aux_in = x.pre; p.epi.aux_ld = I; p.epi.out_bf16 = e->dpre16; p.epi.out_ld = I;
      SUTA_TRY(gemm(e, p, st));
    }
    {  // intermediate_dense dgrad + residual path
      GemmProblem p = dense(e->dpre16, M, I, reinterpret_cast<const bf16*>(w.w1_t), H);
      p.epi.accumulate = 1; p.epi.out_f32 = db; p.epi.out_ld = H;
      SUTA_TRY(gemm(e, p, st));
    }
    if (e->train_all) SUTA_TRY(linear_param_grads(e, e->dpre16, I, act_xf(e, l), H, e->w1_off[l], e->b1_off[l], st));   // intermediate_dense
    PROF_B("ln_bwd", (double)M * H * (4 + 4 + 4 + 2), layernorm_backward(db, x.h1, nullptr, x.mean1, x.rstd1, e->d_row_utt, prm, (int)e->ln1_g[l], (int)e->ln1_b[l],
                                e->G, da, e->b16, M, H, e->d_tok_off, e->d_T, e->U, ln_slot(), st, &lnred.item[lnred.n]));
    lnred.n += 1;
    if (e->train_all) SUTA_TRY(linear_param_grads(e, e->b16, H, x.attn, H, e->wo_off[l], e->bo_off[l], st));            // out_proj
    {  // out_proj dgrad
      GemmProblem p = dense(e->b16, M, H, reinterpret_cast<const bf16*>(w.wo_t), H);
      p.epi.out_bf16 = e->dO16; p.epi.out_ld = H;
      SUTA_TRY(gemm(e, p, st));
    }
    PROF_F("attn_bwd", 8.0 * H * e->sumT2, attention_backward(x.qkv, x.attn, e->dO16, x.lse, e->Dbuf, e->dqkv16, e->d_attn_tab, e->n_attn_blk, H, c.heads, M, st));
    {  // q,k,v dgrad + residual path
      GemmProblem p = dense(e->dqkv16, M, 3 * H, reinterpret_cast<const bf16*>(w.wqkv_t), H);
      p.epi.accumulate = 1; p.epi.out_f32 = da; p.epi.out_ld = H;
      SUTA_TRY(gemm(e, p, st));
    }
    if (e->train_all) SUTA_TRY(linear_param_grads(e, e->dqkv16, 3 * H, act_xa(e, l), H, e->wqkv_off[l], e->bqkv_off[l], st));   // q, k, v
    e->launches += 5;             // 2 x LayerNorm backward, attention backward (3 launches)
  }
  // encoder.layer_norm (post-LN: in front of the layers)
  if (!e->stable) {
    PROF_B("ln_bwd", (double)M * H * (4 + 4 + 4), layernorm_backward(da, e->hE, nullptr, e->enc_mean, e->enc_rstd, e->d_row_utt, prm, (int)e->enc_g, (int)e->enc_b,
                                e->G, db, nullptr, M, H, e->d_tok_off, e->d_T, e->U, ln_slot(), st, &lnred.item[lnred.n]));
    lnred.n += 1;
  }
  // positional conv: d h0 = d hE + conv^T (d hE * GELU'(cpos))
  // (train_all keeps the forward's padded input slabs in xg for the weight gradient and packs the gradient beside them)
  bf16* xgg = e->train_all ? e->xg_grad : e->xg;
  PROF("posconv_pack_grad", posconv_pack_grad(db, e->cpos, e->d_row_utt, e->d_tok_off, e->d_pad_off, xgg, M, H, G, CG, e->R, -(c.pos_k / 2), st));
  if (e->train_all) {
    // d w[co][tap][ci] = sum_t d conv[t, co] * x_pad[t + tap, ci]: per group, d conv^T (rows K/2 .. of the gradient slab) times
    // the overlapping K-tap windows of the input slab, both MN-major over the utterance's frames; d bias = column sums;
    // then through the weight_norm parametrisation (HF:344-352) into d g, d v
    const int K = c.pos_k;
    for (int g = 0; g < G; ++g) {
      GemmProblem p;
      p.a = {xgg + ((long long)g * e->R + K / 2) * CG, M, CG, 1, CG};
      p.b = {e->xg + (long long)g * e->R * CG, M, CG, 1, (long long)K * CG};
      p.M = CG; p.N = K * CG; p.K = 0; p.nz = 1;
      p.ztab = e->d_ztab[c.n_conv];
      p.epi.out_f32 = e->pos_dW + (long long)g * CG * K * CG; p.epi.out_ld = K * CG;
      p.flops = 2.0 * CG * K * CG * (double)M;
      SUTA_TRY(gemm(e, p, st));
      PROF("colsum", colsum_per_utt_bf16(xgg + (long long)g * e->R * CG, e->d_pad_off, e->d_T, e->G, e->n_params, e->pos_b_off + (long long)g * CG,
                                         CG, e->U, st));
    }
    PROF("weight_norm_bwd", posconv_weight_norm_backward(e->P + e->pos_v_off, e->pos_dW, e->wn_scratch, e->G + e->pos_g_off, e->G + e->pos_v_off,
                                                         H, CG, K, st));
    e->launches += G + 3;
  }
  if (use_posconv_tc(e)) {
    PROF("posconv_tc", posconv_tc(xgg, pos_w_t(e), nullptr, e->dcpos, H, G, CG, e->R, e->Rm, c.pos_k, st));
    e->launches += 1;
  } else {
    GemmProblem p;
    p.a = {xgg, (long long)G * e->R - c.pos_k + 1, CG};
    p.b = {pos_w_t(e), H, (long long)c.pos_k * CG};
    p.M = (int)e->Rm; p.N = CG; p.K = c.pos_k * CG;
    p.nz = G; p.a_z_rows = e->R; p.b_z_rows = CG; p.c_z_cols = CG;
    p.epi.out_f32 = e->dcpos; p.epi.out_ld = H;
    SUTA_TRY(gemm(e, p, st));
  }
  PROF("posconv_combine_grad", posconv_combine_grad(db, e->dcpos, e->d_row_utt, e->d_tok_off, e->d_pad_off, e->train_feature ? da : nullptr,
                                e->b16, M, H, -(c.pos_k / 2 - 1), st));
  {  // projection dgrad
    GemmProblem p = dense(e->b16, M, H, reinterpret_cast<const bf16*>(e->w.proj_w_t), C);
    p.epi.out_f32 = e->d_yfp; p.epi.out_ld = C;
    if (e->train_feature) {      // d y = d h0 * W_u : W_u [H (reduction rows), C] is MN-major for this product
      p.b = {e->proj_shadow, (long long)e->U * H, C, 1, C};
      p.mblk = e->d_tok_mblk; p.num_mblk = e->n_tok_mblk;
    }
    SUTA_TRY(gemm(e, p, st));
  }
  e->launches += 4;
  if (e->conv_ln) {
    // ============ LayerNorm feature extractor: the trainable conv LayerNorms put the whole CNN on the backward path ==========
    const int last = c.n_conv - 1;
    const bool tf = e->train_feature != 0;
    if (tf) {  // d W_proj[u] = d h0[u]^T y_fp[u], d b_proj[u] = column sums of d h0[u]  (as in the GroupNorm family below)
      PROF("gelu_grad_pad", gelu_grad_to_padded(da, nullptr, e->dh0_pad, e->d_row_utt, e->d_tok_off, e->d_dpre_off_last, M, H, st));
      GemmProblem p;
      p.a = {e->dh0_pad, e->R64 + 128, H, 1, H};
      p.b = {e->y_fp, M, C, 1, C};
      p.M = H; p.N = C; p.K = 0; p.nz = e->U;
      p.ztab = e->d_ztab[c.n_conv];
      p.epi.out_f32 = e->G + e->proj_w_off; p.epi.out_ld = C; p.out_z_stride = e->n_params;
      p.flops = 2.0 * H * C * (double)M;
      SUTA_TRY(gemm(e, p, st));
      PROF("colsum", colsum_per_utt(da, e->d_tok_off, e->d_T, e->G, e->n_params, e->proj_b_off, H, e->U, st));
      e->launches += 2;
    }
    // feature_projection.layer_norm with input gradient (its input, the last conv layer's output, is bf16)
    PROF_B("ln_bwd C", (double)M * C * (4 + 2 + 4), layernorm_backward(e->d_yfp, nullptr, e->conv_out[last], e->fp_mean, e->fp_rstd, e->d_row_utt, prm, (int)e->fp_g,
                                (int)e->fp_b, e->G, e->d_feat, nullptr, M, C, e->d_tok_off, e->d_T, e->U, ln_slot(), st, &lnred.item[lnred.n]));
    lnred.n += 1;
    // last conv layer: GELU(LayerNorm(z)) backward, in place on the dense fp32 gradient, then into the 128-row-aligned slab
    PROF_B("ln_gelu_bwd C", (double)M * C * (4 + 2 + 4), layernorm_gelu_backward(e->d_feat, nullptr, e->conv_z[last], e->conv_mean[last], e->conv_rstd[last], e->d_row_utt, prm,
                                   (int)e->cln_g[last], (int)e->cln_b[last], e->G, e->d_feat, nullptr, M, C, e->d_tok_off, e->d_T, e->U,
                                   ln_slot(), st, &lnred.item[lnred.n]));
    lnred.n += 1;
    PROF("gelu_grad_pad", gelu_grad_to_padded(e->d_feat, nullptr, e->conv_dpre[last], e->d_row_utt, e->d_tok_off, e->d_dpre_off_last, M, C, st));
    e->launches += 3;
    for (int l = last; l >= 1; --l) {
      const int Cin = c.conv_dim[l - 1], Cout = c.conv_dim[l], k = c.conv_kernel[l];
      const long long dpre_rows = (l == last ? e->R64 : e->rows_total[l]) + 128;
      long long rows_valid = 0, rows_below = 0;
      for (int u = 0; u < e->U; ++u) { rows_valid += e->L[l][u]; rows_below += e->L[l - 1][u]; }
      if (tf) {  // d W_l[u] = d z_l[u]^T im2col(a_{l-1}[u]) (reduction over time), d b_l[u] = column sums of d z_l[u]
        const int s = c.conv_stride[l];
        GemmProblem p;
        p.a = {e->conv_dpre[l], dpre_rows, Cout, 1, Cout};
        p.b = {e->conv_out[l - 1], (e->rows_total[l - 1] + 128 - k) / s + 1, (long long)s * Cin, 1, (long long)k * Cin};
        p.M = Cout; p.N = k * Cin; p.K = 0; p.nz = e->U;
        p.ztab = e->d_ztab[l];
        p.epi.out_f32 = e->G + e->conv_w_off[l]; p.epi.out_ld = k * Cin; p.out_z_stride = e->n_params;
        p.flops = 2.0 * Cout * k * Cin * (double)rows_valid;
        SUTA_TRY(gemm(e, p, st));
        PROF("colsum", colsum_per_utt_bf16(e->conv_dpre[l], l == last ? e->d_dpre_off_last : e->d_off[l], e->d_L[l], e->G, e->n_params,
                                           e->conv_b_off[l], Cout, e->U, st));
        e->launches += 1;
      }
      // d a_{l-1} = transposed conv of d z_l through W_l (frozen and shared, or the utterance's own under train_feature), split
      // by the parity of the input row exactly as in the GroupNorm family below -- no GELU' factor here: the LayerNorm
      // backward that follows applies it
      GemmProblem p;
      p.b = {reinterpret_cast<const bf16*>(e->w.conv_w[l]), Cout, (long long)k * Cin, 1, (long long)k * Cin};
      if (tf) p.b = {e->w_shadow[l], (long long)e->U * Cout, (long long)k * Cin, 1, (long long)k * Cin};
      p.M = (int)rows_valid; p.mblk = e->d_dgrad_mblk[l]; p.num_mblk = e->n_dg_mblk[l];
      p.tiles_own_rows = 1; p.out_rows = (e->rows_total[l - 1] + 128) / 2;
      p.epi.out_ld = 2 * Cin;
      if (k == 2) {
        p.a = {e->conv_dpre[l], dpre_rows, Cout};
        p.N = 2 * Cin; p.K = Cout;
        p.epi.out_bf16 = e->conv_dpre[l - 1];
        SUTA_TRY(gemm(e, p, st));
      } else {
        p.a = {e->conv_dpre[l] - Cout, dpre_rows + 1, Cout};
        p.N = Cin; p.K = 2 * Cout; p.b_kwrap = Cout; p.b_tap_col[0] = 2 * Cin; p.b_tap_col[1] = 0;
        p.epi.out_bf16 = e->conv_dpre[l - 1];
        p.flops = 2.0 * (2 * Cout) * Cin * (double)rows_valid;
        SUTA_TRY(gemm(e, p, st));
        p.a = {e->conv_dpre[l], dpre_rows, Cout};
        p.K = Cout; p.b_tap_col[0] = Cin; p.b_tap_col[1] = Cin;
        p.epi.out_bf16 = e->conv_dpre[l - 1] + Cin;
        p.flops = 2.0 * Cout * Cin * (double)rows_valid;
        SUTA_TRY(gemm(e, p, st));
      }
      // layer l-1: d z = LayerNorm-backward(d a * GELU'), in place (valid rows only; the gaps stay zero); layer 0 has
      // nothing below it that is trainable: parameter gradients only
      PROF_B("ln_gelu_bwd C", (double)rows_below * Cin * (2 + 2 + (l > 1 || tf ? 2 : 0)),
             layernorm_gelu_backward(nullptr, e->conv_dpre[l - 1], e->conv_z[l - 1], e->conv_mean[l - 1], e->conv_rstd[l - 1],
                                     e->d_conv_row_utt[l - 1], prm, (int)e->cln_g[l - 1], (int)e->cln_b[l - 1], e->G, nullptr,
                                     l > 1 || tf ? e->conv_dpre[l - 1] : nullptr, e->rows_total[l - 1], Cin, e->d_off[l - 1], e->d_L[l - 1],
                                     e->U, ln_slot(), st, &lnred.item[lnred.n]));
      lnred.item[lnred.n].tok_off = e->d_off[l - 1];
      lnred.item[lnred.n].T = e->d_L[l - 1];
      lnred.n += 1;
      e->launches += 1;
    }
    if (tf) {  // layer 0: d w_0[c][j] = sum_t d z_0[t,c] x[s t + j], d b_0[c] = sum_t d z_0[t,c]
      Conv0BwdArgs ba{};
      ba.x = e->wav_norm; ba.samp_off = e->d_samp_off; ba.L0 = e->d_L0; ba.out_off = e->d_off0;
      ba.dy = e->conv_dpre[0]; ba.part = e->c0_scratch; ba.n_chunk = conv0_bwd_chunks(e->max_L0);
      ba.G = e->G; ba.pstride = e->n_params; ba.b_off = e->conv_b_off[0]; ba.w_off = e->conv_w_off[0];
      ba.n_utts = e->U; ba.C = c.conv_dim[0]; ba.k = c.conv_kernel[0]; ba.stride = c.conv_stride[0]; ba.max_L0 = e->max_L0;
      ba.plain = 1;
      PROF_B("conv0_bwd", (double)e->S * 4 + (double)e->rows_total[0] * c.conv_dim[0] * 2, conv0_groupnorm_backward(ba, st));
      e->launches += 2;
    }
    PROF("ln_bwd_reduce", layernorm_backward_reduce(lnred, e->d_tok_off, e->d_T, e->U, e->G, e->n_params, st));
    e->launches += 1;
    return SUTA_OK;
  }
  if (!e->train_feature) {
    // feature_projection.layer_norm: parameter gradients only (the CNN below it is frozen)
    SUTA_TRY(layernorm_backward(e->d_yfp, nullptr, e->conv_out[c.n_conv - 1], e->fp_mean, e->fp_rstd, e->d_row_utt, prm,
                                (int)e->fp_g, (int)e->fp_b, e->G, nullptr, nullptr, M, C, e->d_tok_off, e->d_T, e->U, ln_slot(), st,
                                &lnred.item[lnred.n]));
    lnred.n += 1;
    PROF("ln_bwd_reduce", layernorm_backward_reduce(lnred, e->d_tok_off, e->d_T, e->U, e->G, e->n_params, st));
    return SUTA_OK;
  }

  // ================= train_feature: projection weight/bias, then the whole CNN (REF/main.py:88-94) =================
  const int last = c.n_conv - 1;
  {  // d W_proj[u] = d h0[u]^T y_fp[u]  (reduction over the utterance's frames; both operands MN-major)
    PROF("gelu_grad_pad", gelu_grad_to_padded(da, nullptr, e->dh0_pad, e->d_row_utt, e->d_tok_off, e->d_dpre_off_last, M, H, st));
    GemmProblem p;
    p.a = {e->dh0_pad, e->R64 + 128, H, 1, H};
    p.b = {e->y_fp, M, C, 1, C};
    p.M = H; p.N = C; p.K = 0; p.nz = e->U;
    p.ztab = e->d_ztab[c.n_conv];
    p.epi.out_f32 = e->G + e->proj_w_off; p.epi.out_ld = C; p.out_z_stride = e->n_params;
    p.flops = 2.0 * H * C * (double)M;
    SUTA_TRY(gemm(e, p, st));
    PROF("colsum", colsum_per_utt(da, e->d_tok_off, e->d_T, e->G, e->n_params, e->proj_b_off, H, e->U, st));
  }
  // feature_projection.layer_norm with input gradient
  PROF_B("ln_bwd C", (double)M * C * (4 + 2 + 4), layernorm_backward(e->d_yfp, nullptr, e->conv_out[last], e->fp_mean, e->fp_rstd, e->d_row_utt, prm, (int)e->fp_g,
                              (int)e->fp_b, e->G, e->d_feat, nullptr, M, C, e->d_tok_off, e->d_T, e->U, ln_slot(), st, &lnred.item[lnred.n]));
  lnred.n += 1;
  PROF("ln_bwd_reduce", layernorm_backward_reduce(lnred, e->d_tok_off, e->d_T, e->U, e->G, e->n_params, st));
  // d(pre-activation) of the last conv layer, in the 128-row-aligned token slab
  PROF("gelu_grad_pad", gelu_grad_to_padded(e->d_feat, e->conv_pre[last], e->conv_dpre[last], e->d_row_utt, e->d_tok_off,
                               e->d_dpre_off_last, M, C, st));
  e->launches += 4;
  for (int l = last; l >= 1; --l) {
    const int Cin = c.conv_dim[l - 1], Cout = c.conv_dim[l], k = c.conv_kernel[l], s = c.conv_stride[l];
    const long long dpre_rows = (l == last ? e->R64 : e->rows_total[l]) + 128;
    long long rows_valid = 0;
    for (int u = 0; u < e->U; ++u) rows_valid += e->L[l][u];
    {  // d W_l[u] = d pre_l[u]^T  im2col(x_{l-1}[u])   (reduction over time)
      GemmProblem p;
      p.a = {e->conv_dpre[l], dpre_rows, Cout, 1, Cout};
      p.b = {e->conv_out[l - 1], (e->rows_total[l - 1] + 128 - k) / s + 1, (long long)s * Cin, 1, (long long)k * Cin};
      p.M = Cout; p.N = k * Cin; p.K = 0; p.nz = e->U;
      p.ztab = e->d_ztab[l];
      p.epi.out_f32 = e->G + e->conv_w_off[l]; p.epi.out_ld = k * Cin; p.out_z_stride = e->n_params;
      p.flops = 2.0 * Cout * k * Cin * (double)rows_valid;
      SUTA_TRY(gemm(e, p, st));
    }
    if (dgrad_fused(e, l)) {
      // d pre_{l-1} = GELU'(pre_{l-1}) * (transposed conv of d pre_l), written in place by the GEMM epilogue.  The output
      // (and the saved GELU') is addressed as [pairs of rows][2*Cin]: row j of that view = input rows 2j, 2j+1.
      GemmProblem p;
      p.b = {e->w_shadow[l], (long long)e->U * Cout, (long long)k * Cin, 1, (long long)k * Cin};
      p.M = (int)rows_valid; p.mblk = e->d_dgrad_mblk[l]; p.num_mblk = e->n_dg_mblk[l];
      p.tiles_own_rows = 1; p.out_rows = (e->rows_total[l - 1] + 128) / 2;
      p.epi.out_ld = 2 * Cin; p.epi.act = 2; p.epi.aux_ld = 2 * Cin;
      if (k == 2) {
        p.a = {e->conv_dpre[l], dpre_rows, Cout};
        p.N = 2 * Cin; p.K = Cout;
        p.epi.out_bf16 = e->conv_dpre[l - 1]; p.epi.aux_in = e->conv_pre[l - 1];
        SUTA_TRY(gemm(e, p, st));
      } else {
        // even rows: k runs over (dY[j-1], dY[j]) = one 2*Cout-long overlapping row starting one row early; taps (2, 0)
        p.a = {e->conv_dpre[l] - Cout, dpre_rows + 1, Cout};
        p.N = Cin; p.K = 2 * Cout; p.b_kwrap = Cout; p.b_tap_col[0] = 2 * Cin; p.b_tap_col[1] = 0;
        p.epi.out_bf16 = e->conv_dpre[l - 1]; p.epi.aux_in = e->conv_pre[l - 1];
        p.flops = 2.0 * (2 * Cout) * Cin * (double)rows_valid;
        SUTA_TRY(gemm(e, p, st));
        // odd rows: tap 1
        p.a = {e->conv_dpre[l], dpre_rows, Cout};
        p.K = Cout; p.b_tap_col[0] = Cin; p.b_tap_col[1] = Cin;
        p.epi.out_bf16 = e->conv_dpre[l - 1] + Cin; p.epi.aux_in = e->conv_pre[l - 1] + Cin;
        p.flops = 2.0 * Cout * Cin * (double)rows_valid;
        SUTA_TRY(gemm(e, p, st));
      }
      continue;
    }
    {  // Z = d pre_l * W_l[u]  -> [rows_l, (tap, Cin)]
      GemmProblem p;
      p.a = {e->conv_dpre[l], dpre_rows, Cout};
      p.b = {e->w_shadow[l], (long long)e->U * Cout, (long long)k * Cin, 1, (long long)k * Cin};
      p.M = (int)rows_valid; p.N = k * Cin; p.K = Cout;
      p.mblk = e->d_dgrad_mblk[l]; p.num_mblk = e->n_dg_mblk[l];
      p.tiles_own_rows = 1; p.out_rows = dpre_rows;
      p.epi.out_bf16 = e->zbuf; p.epi.out_ld = k * Cin;
      SUTA_TRY(gemm(e, p, st));
    }
    Col2imArgs ca{};
    ca.Z = e->zbuf; ca.pre = e->conv_pre[l - 1]; ca.out = e->conv_dpre[l - 1];
    ca.off_out = e->d_off[l - 1]; ca.off_in = l == last ? e->d_dpre_off_last : e->d_off[l];
    ca.L_out = e->d_L[l - 1]; ca.L_in = e->d_L[l];
    ca.C = Cin; ca.k = k; ca.s = s; ca.n_utts = e->U; ca.max_L_out = e->max_L[l - 1];
    PROF("col2im", conv_col2im_gelu_grad(ca, st));
    e->launches += 1;
  }
  Conv0BwdArgs ba{};
  ba.x = e->wav_norm; ba.samp_off = e->d_samp_off; ba.L0 = e->d_L0; ba.out_off = e->d_off0;
  ba.w = e->P + e->conv_w_off[0]; ba.w_stride = e->n_params;
  ba.dy = e->conv_dpre[0]; ba.stats = e->stats + 2 * e->U;
  ba.mom = e->mom; ba.part = e->c0_scratch; ba.n_chunk = conv0_bwd_chunks(e->max_L0);
  ba.P = e->P; ba.G = e->G; ba.pstride = e->n_params;
  ba.g_off = e->gn_g; ba.b_off = e->gn_b; ba.w_off = e->conv_w_off[0];
  ba.n_utts = e->U; ba.C = c.conv_dim[0]; ba.k = c.conv_kernel[0]; ba.stride = c.conv_stride[0]; ba.max_L0 = e->max_L0;
  PROF_B("conv0_bwd", (double)e->S * 4 + (double)e->rows_total[0] * c.conv_dim[0] * 2, conv0_groupnorm_backward(ba, st));
  e->launches += 2;
  return SUTA_OK;
}

// ---- CUDA graphs of the forward / backward launch chains (small batches) ---------------------------------------
static void drop_graphs(suta_engine* e) {
  for (suta_engine::GraphSlot* g : {&e->g_fwd, &e->g_bwd}) {
    if (g->exec) cudaGraphExecDestroy(g->exec);
    *g = suta_engine::GraphSlot();
  }
}
// Launch-bound batches only: at the benched batch sizes (M ~ 20 k tokens, ~90 us per kernel) the stream launches keep
// the GPU fed and the chains stay eager.  SUTA_NO_GRAPH=1 switches the graphs off, SUTA_GRAPH_MAX_TOKENS moves the bound.
static bool graph_eligible(const suta_engine* e) {
  static const bool off = getenv("SUTA_NO_GRAPH") != nullptr || getenv("SUTA_GEMM_LOG") != nullptr;
  static const long long max_tokens = getenv("SUTA_GRAPH_MAX_TOKENS") ? atoll(getenv("SUTA_GRAPH_MAX_TOKENS")) : 4096;
  return !off && !e->profile && e->M <= max_tokens;
}
// run `chain` (a launch sequence on the stream it is given) eagerly the first time a key is seen, record + replay it the
// second time, replay it afterwards
template <typename F>
static int run_chain(suta_engine* e, suta_engine::GraphSlot& g, unsigned long long key, void* stream, bool is_fwd, F chain) {
  if (!graph_eligible(e)) return chain(stream);
  cudaStream_t st = S(stream);
  if (!(g.exec && g.key == key)) {
    if (!(g.seen && g.seen_key == key)) {        // first call with this key: eager (also runs every one-time kernel set-up)
      if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
      g.seen = true; g.seen_key = key;
      return chain(stream);
    }
    if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    if (!e->cap_stream) CUDA_TRY(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
    const bool fd = e->frontend_done, md = e->moments_done, zd = e->z0_done;
    const long long l0 = e->launches;
    CUDA_TRY(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = chain(e->cap_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(e->cap_stream, &graph);
    if (rc != SUTA_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    CUDA_TRY(ce);
    const cudaError_t ci = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    CUDA_TRY(ci);
    g.key = key;
    g.launches = e->launches - l0;
    g.post_frontend_done = e->frontend_done; g.post_moments_done = e->moments_done; g.post_z0_done = e->z0_done;
    e->frontend_done = fd; e->moments_done = md; e->z0_done = zd;      // nothing ran yet: the replay below does
    e->launches = l0;
  }
  CUDA_TRY(cudaGraphLaunch(g.exec, st));
  e->launches += g.launches;
  e->graph_replays += 1;
  if (is_fwd) { e->frontend_done = g.post_frontend_done; e->moments_done = g.post_moments_done; e->z0_done = g.post_z0_done; }
  return SUTA_OK;
}

extern "C" int suta_forward(suta_engine* e, void* stream) {
  SUTA_CHECK_ARG(e && e->U > 0);
  // the chain depends on which parts of the front end are still valid (LayerNorm-only mode runs the CNN once per batch)
  const unsigned long long key = 1ull | (e->frontend_done ? 2ull : 0ull) | (e->moments_done ? 4ull : 0ull) | (e->z0_done ? 8ull : 0ull);
  return run_chain(e, e->g_fwd, key, stream, true, [&](void* s2) { return forward_eager(e, s2); });
}

extern "C" int suta_loss_backward(suta_engine* e, const suta_hyper* h, void* stream) {
  SUTA_CHECK_ARG(e && e->U > 0 && h);
  // the loss kernels take their hyper-parameters by value: part of the recorded chain (the optimizer's are not: it stays eager)
  unsigned long long key = 1469598103934665603ull;
  auto mix = [&](const void* p, size_t n) {
    for (size_t i = 0; i < n; ++i) key = (key ^ reinterpret_cast<const unsigned char*>(p)[i]) * 1099511628211ull;
  };
  mix(&h->em_coef, sizeof(float)); mix(&h->temp, sizeof(float)); mix(&h->reweight, sizeof(int32_t)); mix(&h->not_blank, sizeof(int32_t));
  mix(&h->div_coef, sizeof(float)); mix(&h->pl_coef, sizeof(float));
  const suta_hyper hc = *h;
  return run_chain(e, e->g_bwd, key | 1ull, stream, false, [&, hc](void* s2) { return loss_backward_eager(e, &hc, s2); });
}

extern "C" int suta_optimizer_step(suta_engine* e, const suta_hyper* h, void* stream) {
  SUTA_CHECK_ARG(e && e->U > 0 && h);
  AdamArgs a{};
  a.P = e->P; a.G = e->G; a.Mom = e->Mom; a.Var = e->Var; a.mult = e->w.mult;
  a.n = e->n_params; a.n_utts = e->U; a.step_index = e->opt_steps;
  a.lr = h->lr; a.beta1 = h->beta1; a.beta2 = h->beta2; a.eps = h->eps; a.weight_decay = h->weight_decay;
  a.kind = h->opt_kind; a.shadow = nullptr;
  if (e->train_all) {
    a.shadow = e->Pb;            // the update writes the bf16 copy of the whole vector
    e->frontend_done = false;
  } else if (e->train_feature) {        // the update also refreshes the bf16 GEMM-operand copies of the trainable matrices
    const suta_model_cfg& c = e->cfg;
    for (int l = 1; l < c.n_conv; ++l) a.seg[a.n_seg++] = {e->conv_w_off[l], e->conv_w_size[l], e->w_shadow[l]};
    a.seg[a.n_seg++] = {e->proj_w_off, (long long)c.hidden * c.conv_dim[c.n_conv - 1], e->proj_shadow};
    e->frontend_done = false;    // the CNN output depends on the updated weights
  }
  if (e->conv_ln) e->frontend_done = false;
  cudaStream_t st = S(stream);
  double shadow_elems = 0.0;
  for (int i = 0; i < a.n_seg; ++i) shadow_elems += (double)a.seg[i].size;
  if (a.shadow) shadow_elems = (double)e->n_params;
  PROF_B("adam", (double)e->U * ((double)e->n_params * 28.0 + shadow_elems * 2.0), optimizer_step(a, st));
  e->opt_steps += 1;
  e->launches += 1;
  if (e->train_all) SUTA_TRY(refresh_train_all(e, st));
  return SUTA_OK;
}

extern "C" int suta_adapt_step(suta_engine* e, const suta_hyper* h, void* stream) {
  SUTA_TRY(suta_loss_backward(e, h, stream));
  SUTA_TRY(suta_optimizer_step(e, h, stream));
  return suta_forward(e, stream);
}

extern "C" int suta_decode(suta_engine* e, void* stream) {
  SUTA_CHECK_ARG(e && e->U > 0);
  e->launches += 1;
  return ctc_greedy_decode(e->logits, e->d_tok_off, e->d_T, e->ids, e->collapsed, e->out_len, e->U, e->cfg.vocab, S(stream));
}

extern "C" float* suta_logits(const suta_engine* e) { return e->logits; }
extern "C" float* suta_dlogits(const suta_engine* e) { return e->dlogits; }
extern "C" float* suta_losses(const suta_engine* e) { return e->losses; }
extern "C" float* suta_params(const suta_engine* e) { return e->P; }
extern "C" float* suta_grads(const suta_engine* e) { return e->G; }
extern "C" int32_t* suta_argmax_ids(const suta_engine* e) { return e->ids; }
extern "C" int32_t* suta_collapsed_ids(const suta_engine* e) { return e->collapsed; }
extern "C" int32_t* suta_collapsed_len(const suta_engine* e) { return e->out_len; }
extern "C" int64_t suta_launch_count(const suta_engine* e) { return e->launches; }
extern "C" int64_t suta_graph_replays(const suta_engine* e) { return e ? e->graph_replays : 0; }
extern "C" float* suta_adam_exp_avg(const suta_engine* e) { return e->Mom; }
extern "C" float* suta_adam_exp_avg_sq(const suta_engine* e) { return e->Var; }
extern "C" int suta_opt_steps(const suta_engine* e) { return e ? e->opt_steps : 0; }
extern "C" int suta_set_opt_steps(suta_engine* e, int steps) {
  SUTA_CHECK_ARG(e && steps >= 0);
  e->opt_steps = steps;
  return SUTA_OK;
}

// dtype: 0 fp32, 1 bf16
extern "C" const void* suta_debug_buffer(const suta_engine* e, const char* name, int64_t* rows, int64_t* cols, int* dtype) {
  if (!e || e->U <= 0 || !name) return nullptr;
  const suta_model_cfg& c = e->cfg;
  const int H = c.hidden, C = c.conv_dim[c.n_conv - 1];
  auto ret = [&](const void* p, long long r, long long k, int dt) { *rows = r; *cols = k; *dtype = dt; return p; };
  std::string n(name);
  if (n == "wav_norm") return ret(e->wav_norm, 1, e->S, 0);
  if (n == "wav") return ret(e->wav, 1, e->S, 0);
  if (n == "feat") return ret(e->conv_out[c.n_conv - 1], e->M, C, 1);
  if (n == "h0") return ret(e->h0, e->M, H, 0);
  if (n == "cpos") return ret(e->cpos, e->Rm, H, 0);
  if (n == "hE") return ret(e->hE, e->M, H, 0);
  if (n == "x_final") return ret(e->fa, e->M, H, 0);
  if (n == "d_yfp") return ret(e->d_yfp, e->M, C, 0);
  if (n.rfind("dpre", 0) == 0 && e->train_feature) {      // d(pre-activation) of conv layer l (train_feature backward)
    int l = atoi(n.c_str() + 4);
    if (l >= 0 && l < c.n_conv) return ret(e->conv_dpre[l], l == c.n_conv - 1 ? e->R64 : e->rows_total[l], c.conv_dim[l], 1);
  }
  if (n.rfind("cpre", 0) == 0 && e->train_feature) {      // GELU'(pre-activation) saved by the forward
    int l = atoi(n.c_str() + 4);
    if (l >= 0 && l < c.n_conv) return ret(e->conv_pre[l], e->rows_total[l], c.conv_dim[l], 1);
  }
  if (n == "d_feat" && e->train_feature) return ret(e->d_feat, e->M, C, 0);
  if (n == "dh0_pad" && e->train_feature) return ret(e->dh0_pad, e->R64, H, 1);
  if (n.rfind("conv", 0) == 0) {
    int l = atoi(n.c_str() + 4);
    if (l >= 0 && l < c.n_conv) return ret(e->conv_out[l], e->rows_total[l], c.conv_dim[l], 1);
  }
  if (n.rfind("h1_", 0) == 0) { int l = atoi(n.c_str() + 3); if (l >= 0 && l < c.layers) return ret(e->lb[l].h1, e->M, H, 0); }
  if (n.rfind("h2_", 0) == 0) { int l = atoi(n.c_str() + 3); if (l >= 0 && l < c.layers) return ret(e->lb[l].h2, e->M, H, 0); }
  if (n.rfind("qkv_", 0) == 0) { int l = atoi(n.c_str() + 4); if (l >= 0 && l < c.layers) return ret(e->lb[l].qkv, e->M, 3 * H, 1); }
  if (n.rfind("attn_", 0) == 0) { int l = atoi(n.c_str() + 5); if (l >= 0 && l < c.layers) return ret(e->lb[l].attn, e->M, H, 1); }
  return nullptr;
}

// ---- single operators ----------------------------------------------------------------------------------
extern "C" int suta_op_gemm(const void* a, int64_t a_rows, int64_t a_row_stride, const void* b, int64_t b_rows,
                            int64_t b_row_stride, int M, int N, int K, float* out_f32, void* out_bf16, int out_ld,
                            const float* bias, const float* residual, int res_ld, int act, const void* aux_in, void* aux_out,
                            int aux_ld, void* stream) {
  GemmProblem p;
  p.a = {reinterpret_cast<const bf16*>(a), a_rows, a_row_stride};
  p.b = {reinterpret_cast<const bf16*>(b), b_rows, b_row_stride};
  p.M = M; p.N = N; p.K = K;
  p.epi.out_f32 = out_f32; p.epi.out_bf16 = reinterpret_cast<bf16*>(out_bf16); p.epi.out_ld = out_ld;
  p.epi.bias = bias; p.epi.residual = residual; p.epi.res_ld = res_ld; p.epi.act = act & 3; p.epi.accumulate = (act >> 2) & 1;
  p.epi.aux_in = reinterpret_cast<const bf16*>(aux_in); p.epi.aux_out = reinterpret_cast<bf16*>(aux_out); p.epi.aux_ld = aux_ld;
  return gemm_bf16_tc(p, S(stream));
}
// D[M,N] = A * B^T-style contraction with per-operand memory order: a_mn / b_mn = 1 means the operand is stored
// [K rows][M|N contiguous] (the layouts of weight-gradient GEMMs); fp32 output.
extern "C" int suta_op_gemm_mn(const void* a, int64_t a_rows, int64_t a_row_stride, int a_mn, const void* b, int64_t b_rows,
                               int64_t b_row_stride, int b_mn, int M, int N, int K, float* out_f32, int out_ld, void* stream) {
  GemmProblem p;
  p.a = {reinterpret_cast<const bf16*>(a), a_rows, a_row_stride, a_mn, a_mn ? M : 0};
  p.b = {reinterpret_cast<const bf16*>(b), b_rows, b_row_stride, b_mn, b_mn ? N : 0};
  p.M = M; p.N = N; p.K = K;
  p.epi.out_f32 = out_f32; p.epi.out_ld = out_ld;
  return gemm_bf16_tc(p, S(stream));
}
extern "C" int suta_op_layernorm_fwd(const float* x_f32, const void* x_bf16, const int32_t* row_utt, const float* P,
                                     int64_t pstride, int g_off, int b_off, float* y_f32, void* y_bf16, float* mean,
                                     float* rstd, int64_t M, int N, float eps, void* stream) {
  return layernorm_forward(x_f32, reinterpret_cast<const bf16*>(x_bf16), row_utt, UttParams{P, pstride}, g_off, b_off, y_f32,
                           reinterpret_cast<bf16*>(y_bf16), mean, rstd, M, N, eps, S(stream));
}
extern "C" int suta_op_layernorm_bwd(const float* dy, const float* x_f32, const void* x_bf16, const float* mean,
                                     const float* rstd, const int32_t* row_utt, const float* P, int64_t pstride, int g_off,
                                     int b_off, float* G, float* dx_f32, void* dx_bf16, int64_t M, int N,
                                     const int64_t* tok_off, const int32_t* T, int n_utts, float* scratch, void* stream) {
  return layernorm_backward(dy, x_f32, reinterpret_cast<const bf16*>(x_bf16), mean, rstd, row_utt, UttParams{P, pstride}, g_off,
                            b_off, G, dx_f32, reinterpret_cast<bf16*>(dx_bf16), M, N,
                            reinterpret_cast<const long long*>(tok_off), T, n_utts, scratch, S(stream));
}
extern "C" int suta_op_layernorm_fwd_mode(const float* x_f32, const void* x_bf16, const int32_t* row_utt, const float* P,
                                          int64_t pstride, int g_off, int b_off, float* y_f32, void* y_bf16, float* mean,
                                          float* rstd, int64_t M, int N, float eps, const float* y32_bias, int mode, void* stream) {
  return layernorm_forward(x_f32, reinterpret_cast<const bf16*>(x_bf16), row_utt, UttParams{P, pstride}, g_off, b_off, y_f32,
                           reinterpret_cast<bf16*>(y_bf16), mean, rstd, M, N, eps, S(stream), y32_bias, mode);
}
extern "C" int suta_op_layernorm_bwd_mode(const float* dy_f32, const void* dy_bf16, const float* x_f32, const void* x_bf16,
                                          const float* mean, const float* rstd, const int32_t* row_utt, const float* P,
                                          int64_t pstride, int g_off, int b_off, float* G, float* dx_f32, void* dx_bf16,
                                          const float* dx_add, int mode, int64_t M, int N, const int64_t* tok_off,
                                          const int32_t* T, int n_utts, float* scratch, void* stream) {
  const long long* to = reinterpret_cast<const long long*>(tok_off);
  if (mode == LN_GELU)
    return layernorm_gelu_backward(dy_f32, reinterpret_cast<const bf16*>(dy_bf16), reinterpret_cast<const bf16*>(x_bf16), mean, rstd,
                                   row_utt, UttParams{P, pstride}, g_off, b_off, G, dx_f32, reinterpret_cast<bf16*>(dx_bf16), M, N, to,
                                   T, n_utts, scratch, S(stream));
  SUTA_CHECK_ARG(dy_f32 && !dy_bf16 && (mode == LN_PLAIN || (mode == LN_KEEP_INPUT && dx_add)));
  return layernorm_backward(dy_f32, x_f32, reinterpret_cast<const bf16*>(x_bf16), mean, rstd, row_utt, UttParams{P, pstride}, g_off,
                            b_off, G, dx_f32, reinterpret_cast<bf16*>(dx_bf16), M, N, to, T, n_utts, scratch, S(stream), nullptr,
                            mode == LN_KEEP_INPUT ? dx_add : nullptr);
}
extern "C" int64_t suta_op_layernorm_bwd_scratch_floats(int N, int n_utts) { return layernorm_backward_scratch_floats(N, n_utts); }
extern "C" int suta_op_attention_fwd(const void* qkv, void* O, float* lse, const int32_t* blk_tab, int n_blk, int H, int heads,
                                     int64_t M, void* stream) {
  return attention_forward(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<bf16*>(O), lse,
                           reinterpret_cast<const int4*>(blk_tab), n_blk, H, heads, M, S(stream));
}
extern "C" int suta_op_attention_bwd(const void* qkv, const void* O, const void* dO, const float* lse, float* D, void* dqkv,
                                     const int32_t* blk_tab, int n_blk, int H, int heads, int64_t M, void* stream) {
  return attention_backward(reinterpret_cast<const bf16*>(qkv), reinterpret_cast<const bf16*>(O),
                            reinterpret_cast<const bf16*>(dO), lse, D, reinterpret_cast<bf16*>(dqkv),
                            reinterpret_cast<const int4*>(blk_tab), n_blk, H, heads, M, S(stream));
}
extern "C" int suta_op_loss(const float* logits, const int64_t* tok_off, const int32_t* T, int n_utts, float em_coef,
                            float temp, int reweight, int not_blank, float div_coef, float* loss, float* dlogits_f32,
                            void* dlogits_bf16, void* stream) {
  LossArgs la{};
  la.logits = logits; la.tok_off = reinterpret_cast<const long long*>(tok_off); la.T = T; la.n_utts = n_utts;
  la.em_coef = em_coef; la.temp = temp; la.reweight = reweight; la.not_blank = not_blank; la.div_coef = div_coef;
  la.loss = loss; la.dlogits_f32 = dlogits_f32; la.dlogits_bf16 = reinterpret_cast<bf16*>(dlogits_bf16);
  return suta_loss_forward_backward(la, S(stream));
}
extern "C" int suta_op_ctc_pseudo_label(const float* logits, const int64_t* tok_off, const int32_t* T, int n_utts,
                                        const int32_t* collapsed, const int32_t* collapsed_len, float* alpha, float* g, float* loss,
                                        float* dlogits_f32, int32_t* target_len, void* stream) {
  SUTA_CHECK_ARG(n_utts > 0 && n_utts <= 4096 && T && tok_off && loss && dlogits_f32);
  // host copies of T: lattice offsets (this entry point is for tests; the engine keeps its tables on the device)
  std::vector<int> hT(n_utts);
  CUDA_TRY(cudaMemcpyAsync(hT.data(), T, sizeof(int) * n_utts, cudaMemcpyDeviceToHost, S(stream)));
  CUDA_TRY(cudaStreamSynchronize(S(stream)));
  std::vector<long long> aoff(n_utts);
  long long tot = 0, M = 0;
  int maxT = 0;
  for (int u = 0; u < n_utts; ++u) { aoff[u] = tot; tot += ctc_alpha_floats(hT[u]); M += hT[u]; maxT = hT[u] > maxT ? hT[u] : maxT; }
  // the offsets ride at the end of the alpha scratch (caller sizes it sum_u T_u (2 T_u + 1) floats + 2 n_utts more)
  long long* d_aoff = reinterpret_cast<long long*>(alpha + ((tot + 1) & ~1LL));
  CUDA_TRY(cudaMemcpyAsync(d_aoff, aoff.data(), sizeof(long long) * n_utts, cudaMemcpyHostToDevice, S(stream)));
  CUDA_TRY(cudaMemsetAsync(dlogits_f32, 0, sizeof(float) * M * 32, S(stream)));
  CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float) * 4 * n_utts, S(stream)));
  CtcArgs ca{};
  ca.logits = logits; ca.tok_off = reinterpret_cast<const long long*>(tok_off); ca.T = T; ca.collapsed = collapsed;
  ca.collapsed_len = collapsed_len; ca.alpha = alpha; ca.alpha_off = d_aoff; ca.g = g; ca.dlogits_f32 = dlogits_f32;
  ca.loss = loss; ca.target_len = target_len; ca.n_utts = n_utts; ca.max_states = 2 * maxT + 1; ca.pl_coef = 1.0f;
  SUTA_TRY(ctc_pseudo_label_loss(ca, S(stream)));
  CUDA_TRY(cudaStreamSynchronize(S(stream)));      // aoff (host) must outlive the upload
  return SUTA_OK;
}
extern "C" int suta_op_softmax_entropy(const float* logits, int64_t rows, float temp, float* out, void* stream) {
  return softmax_entropy_rows(logits, rows, temp, out, S(stream));
}
extern "C" int suta_op_adam(float* P, const float* G, float* Mom, float* Var, const uint8_t* mult, int64_t n, int n_utts,
                            int step_index, const suta_hyper* h, void* shadow_bf16, void* stream) {
  SUTA_CHECK_ARG(h);
  AdamArgs a{};
  a.P = P; a.G = G; a.Mom = Mom; a.Var = Var; a.mult = mult; a.n = n; a.n_utts = n_utts; a.step_index = step_index;
  a.lr = h->lr; a.beta1 = h->beta1; a.beta2 = h->beta2; a.eps = h->eps; a.weight_decay = h->weight_decay;
  a.kind = h->opt_kind; a.shadow = reinterpret_cast<bf16*>(shadow_bf16);
  return optimizer_step(a, S(stream));
}
extern "C" int suta_op_decode(const float* logits, const int64_t* tok_off, const int32_t* T, int n_utts, int V, int32_t* ids,
                              int32_t* collapsed, int32_t* out_len, void* stream) {
  return ctc_greedy_decode(logits, reinterpret_cast<const long long*>(tok_off), T, ids, collapsed, out_len, n_utts, V, S(stream));
}
