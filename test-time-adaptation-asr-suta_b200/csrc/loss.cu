// Fused SUTA loss: value and d loss / d logits in one kernel.
// Restates REF/main.py:181-203 (forward_and_adapt's loss assembly), :26-28 (softmax_entropy), :30-44
// (mcc_loss, incl. the detached entropy re-weighting and the keepdim-less normalisation) and :46-60 (div_loss:
// minus the entropy of softmax(mean-over-time raw logits[1:]), only when div_coef > 0) for a batch of
// independent utterances; the gradient is the closed form verified against the reference's autograd
// (oracle/suta_oracle.py: suta_loss_grad_closed).
//
// One CTA per utterance, one warp per frame, lane == class (vocabulary is 32 = REF/vocab.json,
// class_num=32 at REF/main.py:30).  Pass 1 accumulates the 32x32 class-confusion matrix (lane b keeps
// column b in registers), pass 2 recomputes the softmax and writes the gradient.  Latency-bound (< 3 MB).
#include "kernels.cuh"

namespace {

constexpr int V = 32;
constexpr int WARPS = 8;

struct RowStats {
  float p, logp, H;
  int argmax;
};

__device__ __forceinline__ RowStats row_softmax(float x, float inv_temp, int lane) {
  RowStats r;
  // argmax of the RAW logits, first index on ties (torch.argmax)
  float best = x;
  int bi = lane;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ob = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  r.argmax = bi;
  float z = x * inv_temp;
  float zmax = warp_max(z);
  float e = __expf(z - zmax);
  float se = warp_sum(e);
  r.p = e / se;
  r.logp = (z - zmax) - __logf(se);
  r.H = -warp_sum(r.p * r.logp);
  return r;
}

__global__ void __launch_bounds__(WARPS * 32)
suta_loss_kernel(LossArgs a) {
  __shared__ float sC[WARPS][V][V + 1];     // per-warp partial confusion matrices, then sC[0] = C
  __shared__ float sGG[V][V + 1];           // G + G^T
  __shared__ float s_red[WARPS][4];         // nM, sum H over M, sum w
  __shared__ float s_scal[4];               // nM, sumH, sumW, (unused)
  __shared__ float s_r[V], s_col[V];
  __shared__ float s_xsum[WARPS][V];        // per-warp column sums of the raw logits (div_loss)
  __shared__ float s_gdiv[V];               // d div_loss / d logits[t][c] (same for every frame)
  const int u = blockIdx.x;
  const int T = a.T[u];
  const long long off = a.tok_off[u];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float inv_temp = 1.0f / a.temp;
  const bool use_mcc = (1.0f - a.em_coef) > 0.f;
  const bool use_em = a.em_coef > 0.f;

  // ---------------- pass 1 ----------------
  float colacc[V];
#pragma unroll
  for (int i = 0; i < V; ++i) colacc[i] = 0.f;
  float nM = 0.f, sumH = 0.f, sumW = 0.f, xsum = 0.f;
  for (int t = warp; t < T; t += WARPS) {
    float x = a.logits[(off + t) * V + lane];
    xsum += x;
    RowStats r = row_softmax(x, inv_temp, lane);
    bool inM = a.not_blank ? (r.argmax != 0) : true;
    if (inM) { nM += 1.f; sumH += r.H; }
    float w = a.reweight ? (1.0f + __expf(-r.H)) : 1.0f;
    sumW += w;
    if (use_mcc) {
      float wp = w * r.p;
#pragma unroll
      for (int i = 0; i < V; ++i) colacc[i] += __shfl_sync(0xffffffffu, wp, i) * r.p;   // C[i][lane]
    }
  }
#pragma unroll
  for (int i = 0; i < V; ++i) sC[warp][i][lane] = colacc[i];
  if (lane == 0) { s_red[warp][0] = nM; s_red[warp][1] = sumH; s_red[warp][2] = sumW; }
  s_xsum[warp][lane] = xsum;
  __syncthreads();
  // div_loss (REF/main.py:46-60, called with the raw logits at :202): d = -H(q), q = softmax(mean_t x[t][1:]);
  // d d / d x[t][c] = q_c (log q_c + H(q)) / T for c >= 1 -- every warp computes it redundantly (lane == class)
  float d_loss = 0.f;
  if (a.div_coef > 0.f) {
    float m = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) m += s_xsum[w][lane];
    m = (lane == 0) ? -INFINITY : m / (float)T;            // the blank column is dropped (:53-54)
    const float mmax = warp_max(m);
    const float em = (lane == 0) ? 0.f : __expf(m - mmax);
    const float sm = warp_sum(em);
    const float q = em / sm;
    const float logq = (lane == 0) ? 0.f : (m - mmax) - __logf(sm);
    const float Hq = -warp_sum(q * logq);
    d_loss = -Hq;
    if (warp == 0) s_gdiv[lane] = (lane == 0) ? 0.f : a.div_coef * q * (logq + Hq) / (float)T;
  }
  for (int idx = threadIdx.x; idx < V * V; idx += blockDim.x) {
    int i = idx / V, j = idx % V;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) s += sC[w][i][j];
    sC[0][i][j] = s;          // element (i,j) of every partial is touched by this thread only
  }
  if (threadIdx.x < 3) {
    float s = 0.f;
    for (int w = 0; w < WARPS; ++w) s += s_red[w][threadIdx.x];
    s_scal[threadIdx.x] = s;
  }
  __syncthreads();
  const float nM_tot = s_scal[0];
  const float e_loss = s_scal[1] / nM_tot;                 // NaN when no frame is selected (REF/main.py:190)
  const float wscale = a.reweight ? ((float)T / s_scal[2]) : 1.0f;   // w <- T*w/sum(w) (REF/main.py:36)

  float mcc = 0.f;
  if (use_mcc) {
    // C <- wscale*C ; r_j = sum_k C[j][k] ; col_a = sum_i C[i][a]
    for (int idx = threadIdx.x; idx < V * V; idx += blockDim.x) sC[0][idx / V][idx % V] *= wscale;
    __syncthreads();
    if (warp == 0) {
      float rs = 0.f, cs = 0.f;
#pragma unroll
      for (int k = 0; k < V; ++k) { rs += sC[0][lane][k]; cs += sC[0][k][lane]; }
      s_r[lane] = rs;
      s_col[lane] = cs;
    }
    __syncthreads();
    // mcc = (sum_{ij} C[i][j]/r[j] - sum_i C[i][i]/r[i]) / V      (REF/main.py:41-42)
    float part = 0.f;
    for (int idx = threadIdx.x; idx < V * V; idx += blockDim.x) {
      int i = idx / V, j = idx % V;
      float ch = sC[0][i][j] / s_r[j];
      part += (i == j) ? 0.f : ch;
    }
    part = warp_sum(part);
    if (lane == 0) s_red[warp][3] = part;
    // G[a][b] = (1/V)(1/r_b - col_a/r_a^2 - delta_ab/r_a + C[a][a]/r_a^2);  GG = G + G^T
    for (int idx = threadIdx.x; idx < V * V; idx += blockDim.x) {
      int i = idx / V, j = idx % V;
      float ri = s_r[i], rj = s_r[j];
      float gij = 1.0f / rj - s_col[i] / (ri * ri) + sC[0][i][i] / (ri * ri) - (i == j ? 1.0f / ri : 0.f);
      float gji = 1.0f / ri - s_col[j] / (rj * rj) + sC[0][j][j] / (rj * rj) - (i == j ? 1.0f / rj : 0.f);
      sGG[i][j] = (gij + gji) * (1.0f / V);
    }
    __syncthreads();
    for (int w = 0; w < WARPS; ++w) mcc += s_red[w][3];
    mcc *= (1.0f / V);
  }
  if (threadIdx.x == 0) {
    float loss = 0.f;
    if (use_em) loss += a.em_coef * e_loss;
    if (use_mcc) loss += (1.0f - a.em_coef) * mcc;
    if (a.div_coef > 0.f) loss += a.div_coef * d_loss;
    a.loss[u] = loss;
    a.loss[a.n_utts + u] = use_em ? e_loss : 0.f;
    a.loss[2 * a.n_utts + u] = mcc;
  }
  if (!a.dlogits_f32 && !a.dlogits_bf16) return;

  // ---------------- pass 2: gradient ----------------
  float gg[V];                       // row `lane` of GG
#pragma unroll
  for (int k = 0; k < V; ++k) gg[k] = use_mcc ? sGG[lane][k] : 0.f;
  const float em_scale = (use_em && nM_tot > 0.f) ? a.em_coef / nM_tot : 0.f;   // empty selection: autograd gives 0
  const float mcc_scale = use_mcc ? (1.0f - a.em_coef) : 0.f;
  __syncthreads();                   // s_gdiv (written by warp 0) is read by every warp below
  const float gdiv = a.div_coef > 0.f ? s_gdiv[lane] : 0.f;
  for (int t = warp; t < T; t += WARPS) {
    float x = a.logits[(off + t) * V + lane];
    RowStats r = row_softmax(x, inv_temp, lane);
    bool inM = a.not_blank ? (r.argmax != 0) : true;
    float g = 0.f;
    if (inM) g += em_scale * (-r.p * (r.logp + r.H));
    if (use_mcc) {
      float w = a.reweight ? (1.0f + __expf(-r.H)) * wscale : 1.0f;
      float gp = 0.f;
#pragma unroll
      for (int k = 0; k < V; ++k) gp += gg[k] * __shfl_sync(0xffffffffu, r.p, k);
      gp *= w;
      float s = warp_sum(r.p * gp);
      g += mcc_scale * r.p * (gp - s);
    }
    g = g * inv_temp + gdiv;
    if (a.dlogits_f32) a.dlogits_f32[(off + t) * V + lane] = g;
    if (a.dlogits_bf16) a.dlogits_bf16[(off + t) * V + lane] = __float2bfloat16(g);
  }
}

__global__ void entropy_rows_kernel(const float* __restrict__ logits, long long rows, float inv_temp, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  RowStats r = row_softmax(logits[row * V + lane], inv_temp, lane);
  if (lane == 0) out[row] = r.H;
}

}  // namespace

int softmax_entropy_rows(const float* logits, long long rows, float temp, float* out, cudaStream_t stream) {
  SUTA_CHECK_ARG(logits && out && temp > 0.f);
  if (rows <= 0) return SUTA_OK;
  entropy_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(logits, rows, 1.0f / temp, out);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

int suta_loss_forward_backward(const LossArgs& a, cudaStream_t stream) {
  SUTA_CHECK_ARG(a.n_utts > 0 && a.logits && a.loss && a.temp > 0.f);
  suta_loss_kernel<<<a.n_utts, WARPS * 32, 0, stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}
