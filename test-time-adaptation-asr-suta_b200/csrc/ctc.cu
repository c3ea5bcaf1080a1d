// Pseudo-label CTC loss of the SDPL baseline, forward + backward in one kernel.
// Restates REF/main_SDPL.py:194-209 (pseudo_labeling_loss) + :176 (loss * (1 - pl_coef) + pl * pl_coef) for a batch of
// independent utterances, with the arithmetic of torch.nn.CTCLoss on the CPU (aten/native/LossCTC.cpp), which is what
// the reference's autograd executes:
//   target   = greedy transcript of the logits: argmax, collapse repeats, drop blank (csrc/decode.cu), minus leading and
//              trailing word delimiters (str.strip() in batch_decode), re-encoded character by character
//   lp[t,c]  = logits[t,c] - logsumexp_t' logits[t',c]          -- log_softmax(1): over the TIME axis (:204)
//   nll      = -log sum over alignments prod_t exp(lp[t, l'_t]),   loss = nll / max(L, 1)   (reduction = 'mean')
//   d loss / d lp[t,c]  = (exp(lp[t,c]) - exp(lcab[t,c] + nll - lp[t,c])) / max(L, 1),  lcab = log sum_{s: l'_s = c} alpha_t(s) beta_t(s)
//              -- the formula LossCTC.cpp uses (it assumes lp is a log_softmax over classes; it is applied as is here too)
//   d loss / d logits[t,c] = g[t,c] - exp(lp[t,c]) * sum_t' g[t',c]                          (log_softmax backward over time)
// One CTA per utterance: the recursions are sequential in t and parallel over the 2L+1 extended-target states.  alpha is
// kept in global memory ([T][S] per utterance), beta only as two rows in shared memory.  Every reduction has a fixed
// order (per-class state lists are built by one thread per class), so the result is bit-reproducible.
#include "kernels.cuh"

namespace {

constexpr int V = 32;
constexpr int CTC_THREADS = 512;
constexpr int CTC_WARPS = CTC_THREADS / 32;
constexpr float NEG = -1e30f;          // log(0): finite, so that NEG - NEG never produces NaN
constexpr int DELIM = 4;               // word delimiter '|' (REF/vocab.json)

__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  if (m <= -1e29f) return NEG;
  return m + __logf(__expf(a - m) + __expf(b - m));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m <= -1e29f) return NEG;
  return m + __logf(__expf(a - m) + __expf(b - m) + __expf(c - m));
}

__global__ void __launch_bounds__(CTC_THREADS)
ctc_pseudo_label_kernel(CtcArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int u = blockIdx.x;
  const int T = a.T[u];
  const long long off = a.tok_off[u];
  const float* x = a.logits + off * V;
  const int* ids_raw = a.collapsed + off;
  const int n_raw = a.collapsed_len[u];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* A = a.alpha + a.alpha_off[u];                 // [T][S]

  __shared__ int s_first, s_last;
  __shared__ float s_lse[V], s_colsum[V], s_nll;
  __shared__ float s_part[CTC_WARPS][V][2];            // per-warp (max, sum) of the column logsumexp
  __shared__ int s_cls_ptr[V + 1];
  __shared__ float s_row[2][V];                        // lp[t][:] of the current / next frame

  // ---- target: strip leading / trailing delimiters ----
  if (tid == 0) { s_first = n_raw; s_last = -1; }
  __syncthreads();
  for (int i = tid; i < n_raw; i += CTC_THREADS)
    if (ids_raw[i] != DELIM) { atomicMin(&s_first, i); atomicMax(&s_last, i); }
  __syncthreads();
  const int L = s_last >= s_first ? s_last - s_first + 1 : 0;
  const int S = 2 * L + 1;
  const int* ids = ids_raw + s_first;
  const int Smax = a.max_states;
  float* rowA = reinterpret_cast<float*>(smem_raw);            // [2][Smax] alpha / beta rows
  float* term = rowA + 2 * Smax;                               // [Smax]
  int* cls_idx = reinterpret_cast<int*>(term + Smax);          // [Smax] states grouped by class
  uint8_t* lab = reinterpret_cast<uint8_t*>(cls_idx + Smax);   // [Smax] label of every extended state
  for (int s = tid; s < S; s += CTC_THREADS) lab[s] = (s & 1) ? (uint8_t)ids[s >> 1] : (uint8_t)0;

  // ---- column logsumexp over time (online max / sum per warp, merged in warp order) ----
  {
    float m = NEG, sum = 0.f;
    for (int t = warp; t < T; t += CTC_WARPS) {
      const float v = x[(long long)t * V + lane];
      const float mn = fmaxf(m, v);
      sum = sum * __expf(m - mn) + __expf(v - mn);
      m = mn;
    }
    s_part[warp][lane][0] = m;
    s_part[warp][lane][1] = sum;
  }
  // ---- per-class state lists: one thread per class scans the target in order ----
  if (tid < V) {
    int n = 0;
    if (tid == 0) n = L + 1;
    else for (int i = 0; i < L; ++i) n += ids[i] == tid;
    s_cls_ptr[tid + 1] = n;
  }
  __syncthreads();
  if (tid == 0) {
    s_cls_ptr[0] = 0;
    for (int c = 0; c < V; ++c) s_cls_ptr[c + 1] += s_cls_ptr[c];
  }
  if (warp == 1) {
    float m = NEG, sum = 0.f;
    for (int w = 0; w < CTC_WARPS; ++w) {
      const float mw = s_part[w][lane][0], sw = s_part[w][lane][1];
      const float mn = fmaxf(m, mw);
      sum = sum * __expf(m - mn) + sw * __expf(mw - mn);
      m = mn;
    }
    s_lse[lane] = m + __logf(sum);
  }
  __syncthreads();
  if (tid < V) {
    int p = s_cls_ptr[tid];
    if (tid == 0) for (int s = 0; s < S; s += 2) cls_idx[p++] = s;
    else for (int i = 0; i < L; ++i) if (ids[i] == tid) cls_idx[p++] = 2 * i + 1;
  }
  if (tid < V) s_row[0][tid] = x[tid] - s_lse[tid];
  __syncthreads();

  // ---- alpha ----
  float* prev = rowA;
  float* cur = rowA + Smax;
  for (int s = tid; s < S; s += CTC_THREADS) {
    const float v = s < 2 ? s_row[0][lab[s]] : NEG;
    prev[s] = v;
    A[s] = v;
  }
  __syncthreads();
  for (int t = 1; t < T; ++t) {
    const float* lp = s_row[(t - 1) & 1];
    float* lpn = s_row[t & 1];
    if (tid < V) lpn[tid] = x[(long long)t * V + tid] - s_lse[tid];
    __syncthreads();
    for (int s = tid; s < S; s += CTC_THREADS) {
      const int c = lab[s];
      const float a0 = prev[s];
      const float a1 = s >= 1 ? prev[s - 1] : NEG;
      const float a2 = (s >= 2 && c != 0 && c != lab[s - 2]) ? prev[s - 2] : NEG;
      const float v = lse3(a0, a1, a2);
      const float r = v <= -1e29f ? NEG : v + lpn[c];
      cur[s] = r;
      A[(long long)t * S + s] = r;
    }
    __syncthreads();
    float* tmp = prev; prev = cur; cur = tmp;
    (void)lp;
  }
  if (tid == 0) {
    const float l = lse2(prev[S - 1], S > 1 ? prev[S - 2] : NEG);
    s_nll = -l;
  }
  __syncthreads();
  const float nll = s_nll;
  const float invL = 1.0f / (float)(L > 0 ? L : 1);

  // ---- beta, occupancies, g = d loss / d lp; running column sums of g ----
  // prev / cur now hold beta rows; s_row[(T-1)&1] holds lp[T-1]
  float colsum0 = 0.f, colsum1 = 0.f;                 // lane 0 of warp w: classes w and w + 16
  for (int t = T - 1; t >= 0; --t) {
    const float* lp = s_row[t & 1];
    if (t > 0 && tid < V) s_row[(t - 1) & 1][tid] = x[(long long)(t - 1) * V + tid] - s_lse[tid];
    for (int s = tid; s < S; s += CTC_THREADS) {
      const int c = lab[s];
      float b;
      if (t == T - 1) {
        b = s >= S - 2 ? lp[c] : NEG;
      } else {
        const float b0 = prev[s];
        const float b1 = s + 1 < S ? prev[s + 1] : NEG;
        const float b2 = (s + 2 < S && c != 0 && c != lab[s + 2]) ? prev[s + 2] : NEG;
        const float v = lse3(b0, b1, b2);
        b = v <= -1e29f ? NEG : v + lp[c];
      }
      cur[s] = b;
      const float ab = A[(long long)t * S + s] + b;
      term[s] = ab <= -1e29f ? 0.f : __expf(ab + nll - lp[c]);
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = warp + h * CTC_WARPS;
      float occ = 0.f;
      for (int i = s_cls_ptr[c] + lane; i < s_cls_ptr[c + 1]; i += 32) occ += term[cls_idx[i]];
      occ = warp_sum(occ);
      if (lane == 0) {
        const float g = (__expf(lp[c]) - occ) * invL;
        a.g[(off + t) * V + c] = g;
        if (h == 0) colsum0 += g; else colsum1 += g;
      }
    }
    __syncthreads();
    float* tmp = prev; prev = cur; cur = tmp;
  }
  if (lane == 0) {
    s_colsum[warp] = colsum0;
    s_colsum[warp + CTC_WARPS] = colsum1;
  }
  __syncthreads();

  // ---- d loss / d logits through the log_softmax over time; mix with the SUTA gradient (REF/main_SDPL.py:176) ----
  const float pl = a.pl_coef, keep = 1.0f - a.pl_coef;
  for (int i = tid; i < T * V; i += CTC_THREADS) {
    const int c = i & (V - 1);
    const long long idx = off * V + i;
    const float p_t = __expf(x[i] - s_lse[c]);
    const float gx = a.g[idx] - p_t * s_colsum[c];
    const float d = keep * (a.dlogits_f32 ? a.dlogits_f32[idx] : 0.f) + pl * gx;
    if (a.dlogits_f32) a.dlogits_f32[idx] = d;
    if (a.dlogits_bf16) a.dlogits_bf16[idx] = __float2bfloat16(d);
  }
  if (tid == 0) {
    const float ctc = nll * invL;
    if (a.loss) {
      // NaN * 0 = NaN in the reference as well (an all-blank utterance under --non_blank poisons the sum)
      a.loss[u] = a.loss[u] * keep + ctc * pl;
      a.loss[3 * a.n_utts + u] = ctc;
    }
    if (a.target_len) a.target_len[u] = L;
  }
}

}  // namespace

long long ctc_alpha_floats(int T) { return (long long)T * (2LL * T + 1); }

int ctc_pseudo_label_loss(const CtcArgs& a, cudaStream_t stream) {
  SUTA_CHECK_ARG(a.n_utts > 0 && a.logits && a.collapsed && a.collapsed_len && a.alpha && a.alpha_off && a.g && a.max_states > 0);
  SUTA_CHECK_ARG(a.pl_coef > 0.f && a.pl_coef <= 1.f);
  const size_t smem = (size_t)a.max_states * (2 * 4 + 4 + 4 + 1) + 16;
  static size_t attr = 0;
  if (smem > attr) {
    CUDA_TRY(cudaFuncSetAttribute(ctc_pseudo_label_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  ctc_pseudo_label_kernel<<<a.n_utts, CTC_THREADS, smem, stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}
