// Greedy CTC decode on the device: argmax over the vocabulary, collapse repeats, drop blank.
// Restates torch.argmax (REF/main.py:333) + the groupby/filter of HF/tokenization_wav2vec2.py:311,317.
// The id -> character mapping (HF/tokenization_wav2vec2.py:320-322) stays on the host.  Integer work: bit-exact.
#include "kernels.cuh"

namespace {

__global__ void __launch_bounds__(256)
decode_kernel(const float* __restrict__ logits, const long long* __restrict__ tok_off, const int* __restrict__ Tlen,
              int* __restrict__ ids, int* __restrict__ collapsed, int* __restrict__ out_len, int V) {
  const int u = blockIdx.x;
  const int T = Tlen[u];
  const long long off = tok_off[u];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // 1. argmax per frame, first index on ties
  for (int t = warp; t < T; t += nw) {
    const float* row = logits + (off + t) * V;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = lane; c < V; c += 32) {
      float x = row[c];
      if (x > best || (x == best && c < bi) || (x != x && !(best != best))) { best = x; bi = c; }   // NaN wins like torch
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      bool take = (ob > best) || (ob == best && oi < bi) || (ob != ob && (best == best || oi < bi));
      if (best != best && !(ob != ob)) take = false;
      if (take) { best = ob; bi = oi; }
    }
    if (lane == 0) ids[off + t] = bi;
  }
  __syncthreads();
  // 2. keep[t] = ids[t] != blank && ids[t] != ids[t-1]; compact with a block scan over per-thread chunks
  __shared__ int s_cnt[256];
  const int per = (T + blockDim.x - 1) / blockDim.x;
  const int b = threadIdx.x * per, e = min(T, b + per);
  int cnt = 0;
  for (int t = b; t < e; ++t) {
    int id = ids[off + t];
    cnt += (id != 0 && (t == 0 || id != ids[off + t - 1]));
  }
  s_cnt[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = 1; o < (int)blockDim.x; o <<= 1) {
    int v = threadIdx.x >= o ? s_cnt[threadIdx.x - o] : 0;
    __syncthreads();
    s_cnt[threadIdx.x] += v;
    __syncthreads();
  }
  int pos = s_cnt[threadIdx.x] - cnt;
  for (int t = b; t < e; ++t) {
    int id = ids[off + t];
    if (id != 0 && (t == 0 || id != ids[off + t - 1])) collapsed[off + pos++] = id;
  }
  if (threadIdx.x == blockDim.x - 1) out_len[u] = s_cnt[threadIdx.x];
}

}  // namespace

int ctc_greedy_decode(const float* logits, const long long* tok_off, const int* T, int* ids, int* collapsed,
                      int* out_len, int n_utts, int V, cudaStream_t stream) {
  SUTA_CHECK_ARG(n_utts > 0 && V > 0);
  decode_kernel<<<n_utts, 256, 0, stream>>>(logits, tok_off, T, ids, collapsed, out_len, V);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}
