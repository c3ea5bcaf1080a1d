// Waveform front end: per-utterance input normalisation and the first feature-extractor layer.
//   normalize_audio       HF/feature_extraction_wav2vec2.py:95   (x - mean) / sqrt(var + 1e-7)
//   conv0_groupnorm_gelu  HF/modeling_wav2vec2.py:319-323        Conv1d(1->C,k=10,s=5,no bias) + GroupNorm(C groups,
//                         i.e. per-channel statistics over TIME) + erf-GELU      
// Cin = 1, so this layer is HBM-bound (writes L0*C bf16, reads 4 B per sample): direct convolution on CUDA cores,
// channels-last bf16 output so the next layer's implicit GEMM can read overlapping rows straight through TMA.
// In a batch, GroupNorm statistics are taken over each utterance's own valid frames only.
//
// The statistics never touch the conv output: with Sx[j] = sum_t x[s t + j] and R[j][j'] = sum_t x[s t + j] x[s t + j']
// (k + k(k+1)/2 numbers per utterance, computed ONCE per batch from the audio by audio_conv0_moments),
//   sum_t y[t,c] = w_c . Sx        sum_t y[t,c]^2 = w_c^T R w_c
// so a weight update (train_feature) only costs a 512 x 100-MAC kernel, not a pass over the waveform.
#include <stdio.h>
#include <stdlib.h>

#include "kernels.cuh"

namespace {

constexpr int NORM_CHUNK = 16384;

__global__ void audio_stats_kernel(const float* __restrict__ wav, const long long* __restrict__ samp_off,
                                   const int* __restrict__ n_samples, double* __restrict__ stats) {
  const int u = blockIdx.y;
  const int n = n_samples[u];
  const int i0 = blockIdx.x * NORM_CHUNK;
  if (i0 >= n) return;
  const float* x = wav + samp_off[u];
  double s = 0.0, ss = 0.0;
  for (int i = i0 + threadIdx.x; i < min(n, i0 + NORM_CHUNK); i += blockDim.x) {
    float v = x[i];
    s += v;
    ss += (double)v * v;
  }
  __shared__ double sh[2][8];
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += sh[0][w]; b += sh[1][w]; }
    atomicAdd(&stats[2 * u], a);
    atomicAdd(&stats[2 * u + 1], b);
  }
}

__global__ void audio_apply_kernel(const float* __restrict__ wav, float* __restrict__ out,
                                   const long long* __restrict__ samp_off, const int* __restrict__ n_samples,
                                   const double* __restrict__ stats) {
  const int u = blockIdx.y;
  const int n = n_samples[u];
  const int i0 = blockIdx.x * NORM_CHUNK;
  if (i0 >= n) return;
  const double mean = stats[2 * u] / n;
  const double var = stats[2 * u + 1] / n - mean * mean;
  const float m = (float)mean, r = (float)(1.0 / sqrt(var + 1e-7));
  const float* x = wav + samp_off[u];
  float* y = out + samp_off[u];
  for (int i = i0 + threadIdx.x; i < min(n, i0 + NORM_CHUNK); i += blockDim.x) y[i] = (x[i] - m) * r;
}

// REF/data.py:23: wav += extra_noise * randn_like(wav), before the processor's normalisation.  Counter-based
// Philox4x32-10 keyed by (seed, utterance id) and counted by sample index: the noise of an utterance depends on
// nothing but (seed, its id, the sample position) -- not on the batch it happens to be adapted in.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;      // (0, 1]
  const float u2 = (float)b * 2.3283064365386963e-10f;               // [0, 1)
  const float r = sqrtf(-2.0f * __logf(u1));
  float sn, cs;
  __sincosf(6.283185307179586f * u2, &sn, &cs);
  return make_float2(r * cs, r * sn);
}

__global__ void audio_noise_kernel(float* __restrict__ wav, const long long* __restrict__ samp_off,
                                   const int* __restrict__ n_samples, const int* __restrict__ utt_id, float sigma,
                                   unsigned long long seed) {
  const int u = blockIdx.y;
  const int n = n_samples[u];
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const uint4 r = philox4x32_10(make_uint4((uint32_t)(i >> 2), 0u, (uint32_t)(utt_id ? utt_id[u] : u), 0u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
  float* x = wav + samp_off[u] + i;
  const float z[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (i + j < n) x[j] += sigma * z[j];
}

constexpr int C0_TT_DEFAULT = 128;      // frames per CTA
constexpr int C0_MAXK = 16;

// MODE 0: accumulate per-channel sum / sum of squares;  MODE 1: normalise + GELU + store
template <int MODE, int C0_TT, int MINB>
__global__ void __launch_bounds__(256, MINB)
conv0_kernel(Conv0Args a, double* __restrict__ stats) {
  extern __shared__ float sx[];                  // C0_TT*stride + k samples
  const int u = blockIdx.y;
  const int L0 = a.L0[u];
  const int t0 = blockIdx.x * C0_TT;
  if (t0 >= L0) return;
  const int nt = min(C0_TT, L0 - t0);
  const float* x = a.x + a.samp_off[u] + (long long)t0 * a.stride;
  const int nx = (nt - 1) * a.stride + a.k;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) sx[i] = x[i];
  __syncthreads();
  const float* w = a.w + (a.w_stride ? (long long)u * a.w_stride : 0);
  for (int cp = threadIdx.x; 2 * cp < a.C; cp += blockDim.x) {
    const int c = 2 * cp;
    float w0[C0_MAXK], w1[C0_MAXK];
#pragma unroll
    for (int j = 0; j < C0_MAXK; ++j) {
      w0[j] = j < a.k ? w[(long long)c * a.k + j] : 0.f;
      w1[j] = j < a.k ? w[(long long)(c + 1) * a.k + j] : 0.f;
    }
    if (MODE == 0) {
      float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
      for (int t = 0; t < nt; ++t) {
        const float* xs = sx + t * a.stride;
        float y0 = 0.f, y1 = 0.f;
#pragma unroll
        for (int j = 0; j < C0_MAXK; ++j)
          if (j < a.k) { y0 += w0[j] * xs[j]; y1 += w1[j] * xs[j]; }
        s0 += y0; q0 += y0 * y0; s1 += y1; q1 += y1 * y1;
      }
      double* st = stats + ((long long)u * a.C + c) * 2;
      atomicAdd(st + 0, (double)s0); atomicAdd(st + 1, (double)q0);
      atomicAdd(st + 2, (double)s1); atomicAdd(st + 3, (double)q1);
    } else {
      const double* st = stats + ((long long)u * a.C + c) * 2;
      const double m0 = st[0] / L0, m1 = st[2] / L0;
      const float mean0 = (float)m0, mean1 = (float)m1;
      const float r0 = (float)(1.0 / sqrt(st[1] / L0 - m0 * m0 + 1e-5));
      const float r1 = (float)(1.0 / sqrt(st[3] / L0 - m1 * m1 + 1e-5));
      float g0, g1, b0, b1;
      if (a.gn.P) {
        const float* P = a.gn.P + (long long)u * a.gn.stride;
        g0 = P[a.g_off + c]; g1 = P[a.g_off + c + 1]; b0 = P[a.b_off + c]; b1 = P[a.b_off + c + 1];
      } else {
        g0 = a.gn_shared_g[c]; g1 = a.gn_shared_g[c + 1]; b0 = a.gn_shared_b[c]; b1 = a.gn_shared_b[c + 1];
      }
      // GroupNorm affine folded into the taps: y = sum_j (w_j r g) x_j + (b - mean r g); channel pairs ride in packed
      // fp32 registers (FFMA2), so a frame costs 10 packed FMAs + one packed GELU/GELU' for two channels
      const float s0 = r0 * g0, s1 = r1 * g1;
      const float c0 = b0 - mean0 * s0, c1 = b1 - mean1 * s1;
      const uint64_t bias2 = pk2(c0, c1);
      uint64_t w2[C0_MAXK];
#pragma unroll
      for (int j = 0; j < C0_MAXK; ++j) w2[j] = pk2(w0[j] * s0, w1[j] * s1);
      const long long row0 = a.out_off[u] + t0;
      auto emit = [&](int t, uint64_t y) {
        const long long o = (row0 + t) * a.C + c;
        float y0, y1, g0v, g1v, d0v, d1v;
        upk2(y, y0, y1);
        gelu_erf_both2(y0, y1, g0v, g1v, d0v, d1v);
        if (a.pre_out) *reinterpret_cast<uint32_t*>(a.pre_out + o) = pack_bf16x2(d0v, d1v);
        *reinterpret_cast<uint32_t*>(a.out + o) = pack_bf16x2(g0v, g1v);
      };
      int t = 0;
      if (a.stride == 5 && a.k == 10) {          // wav2vec2: 4 frames = 20 samples = five 16-byte words; 7 vector loads per 4 frames
        for (; t + 4 <= nt; t += 4) {
          float xw[28];
#pragma unroll
          for (int v = 0; v < 7; ++v) {
            const float4 f = *reinterpret_cast<const float4*>(sx + 5 * t + 4 * v);
            xw[4 * v] = f.x; xw[4 * v + 1] = f.y; xw[4 * v + 2] = f.z; xw[4 * v + 3] = f.w;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint64_t y = bias2;
#pragma unroll
            for (int j = 0; j < 10; ++j) y = fma2(w2[j], dup2(xw[5 * q + j]), y);   // FFMA2 takes the sample as a broadcast operand
            emit(t + q, y);
          }
        }
      }
      for (; t < nt; ++t) {
        const float* xs = sx + t * a.stride;
        uint64_t y = bias2;
#pragma unroll
        for (int j = 0; j < C0_MAXK; ++j)
          if (j < a.k) y = fma2(w2[j], dup2(xs[j]), y);
        emit(t, y);
      }
    }
  }
}

// ---- audio moments: Sx[j] and R[j][j'] (j <= j') per utterance, double precision ---------------------------------
constexpr int MOM_TT = 2048;    // frames per CTA

__global__ void __launch_bounds__(160)
moments_kernel(const float* __restrict__ xin, const long long* __restrict__ samp_off, const int* __restrict__ L0p,
               double* __restrict__ mom, int k, int stride) {
  extern __shared__ float sx[];
  const int u = blockIdx.y;
  const int L0 = L0p[u];
  const int t0 = blockIdx.x * MOM_TT;
  if (t0 >= L0) return;
  const int nt = min(MOM_TT, L0 - t0);
  const float* x = xin + samp_off[u] + (long long)t0 * stride;
  const int nx = (nt - 1) * stride + k;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) sx[i] = x[i];
  __syncthreads();
  const int npair = k + k * (k + 1) / 2;
  for (int p = threadIdx.x; p < npair; p += blockDim.x) {
    int j = p, j2 = -1;                        // p < k: Sx[j];  else the (j, j2) entry of R
    if (p >= k) {
      int q = p - k;
      j = 0;
      while (q >= k - j) { q -= k - j; ++j; }
      j2 = j + q;
    }
    double acc = 0.0;
    if (j2 < 0) {
      for (int t = 0; t < nt; ++t) acc += (double)sx[t * stride + j];
    } else {
      for (int t = 0; t < nt; ++t) acc += (double)sx[t * stride + j] * (double)sx[t * stride + j2];
    }
    atomicAdd(mom + (long long)u * npair + p, acc);
  }
}

// per (utterance, channel): sum and sum of squares of the conv0 output from the audio moments and the current weights
__global__ void conv0_stats_kernel(Conv0Args a) {
  const int u = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C) return;
  const int k = a.k;
  const int npair = k + k * (k + 1) / 2;
  const double* mom = a.mom + (long long)u * npair;
  const float* w = a.w + (a.w_stride ? (long long)u * a.w_stride : 0) + (long long)c * k;
  double wj[C0_MAXK];
#pragma unroll
  for (int j = 0; j < C0_MAXK; ++j) wj[j] = j < k ? (double)w[j] : 0.0;
  double s = 0.0, q = 0.0;
  int p = k;
  for (int j = 0; j < k; ++j) {
    s += wj[j] * mom[j];
    for (int j2 = j; j2 < k; ++j2, ++p) q += (j2 == j ? 1.0 : 2.0) * wj[j] * wj[j2] * mom[p];
  }
  double* st = a.stats + ((long long)u * a.C + c) * 2;
  st[0] = s;
  st[1] = q;
}

}  // namespace

int normalize_audio(const float* wav, float* out, const long long* samp_off, const int* n_samples, int n_utts,
                    int max_samples, double* stats_scratch, cudaStream_t stream) {
  SUTA_CHECK_ARG(n_utts > 0 && max_samples > 0 && stats_scratch);
  CUDA_TRY(cudaMemsetAsync(stats_scratch, 0, sizeof(double) * 2 * n_utts, stream));
  dim3 grid(ceil_div(max_samples, NORM_CHUNK), n_utts);
  audio_stats_kernel<<<grid, 256, 0, stream>>>(wav, samp_off, n_samples, stats_scratch);
  audio_apply_kernel<<<grid, 256, 0, stream>>>(wav, out, samp_off, n_samples, stats_scratch);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

int audio_add_noise(float* wav, const long long* samp_off, const int* n_samples, const int* utt_id, int n_utts,
                    int max_samples, float sigma, unsigned long long seed, cudaStream_t stream) {
  SUTA_CHECK_ARG(wav && n_utts > 0 && max_samples > 0 && sigma >= 0.f);
  if (sigma == 0.f) return SUTA_OK;
  audio_noise_kernel<<<dim3(ceil_div(ceil_div(max_samples, 4), 256), n_utts), 256, 0, stream>>>(wav, samp_off, n_samples, utt_id,
                                                                                               sigma, seed);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

int audio_conv0_moments(const float* x, const long long* samp_off, const int* L0, double* mom, int k, int stride, int n_utts,
                        int max_L0, cudaStream_t stream) {
  SUTA_CHECK_ARG(k <= C0_MAXK && n_utts > 0 && max_L0 > 0);
  const int npair = k + k * (k + 1) / 2;
  CUDA_TRY(cudaMemsetAsync(mom, 0, sizeof(double) * (size_t)npair * n_utts, stream));
  dim3 grid(ceil_div(max_L0, MOM_TT), n_utts);
  size_t smem = sizeof(float) * (MOM_TT * stride + k);
  moments_kernel<<<grid, 160, smem, stream>>>(x, samp_off, L0, mom, k, stride);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

int conv0_groupnorm_gelu(const Conv0Args& a, cudaStream_t stream) {
  SUTA_CHECK_ARG(a.k <= C0_MAXK && a.C % 2 == 0 && a.n_utts > 0 && a.mom);
  conv0_stats_kernel<<<dim3(ceil_div(a.C, 128), a.n_utts), 128, 0, stream>>>(a);
  // frames per CTA x minimum resident CTAs were swept on the B200 (128/256/512 x 1/3/4: 9.4-10.0 ms per step, all within
  // 6 %): the kernel is bound by instruction throughput (packed FMA + MUFU + bf16 packing), not by occupancy or latency
  dim3 grid(ceil_div(a.max_L0, C0_TT_DEFAULT), a.n_utts);
  size_t smem = sizeof(float) * (C0_TT_DEFAULT * a.stride + a.k + 32);       // slack for the 16-byte window loads
  conv0_kernel<1, C0_TT_DEFAULT, 3><<<grid, 256, smem, stream>>>(a, a.stats);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

// ---- LayerNorm feature extractor (feat_extract_norm == "layer", HF/modeling_wav2vec2.py:275-299) ----------------------
// Layer 0 there is Conv1d(1 -> C, k, stride, bias) with NO statistics over time: z0[t, c] = b[c] + sum_j w[c][j] x[s t + j],
// written channels-last in bf16; the per-frame LayerNorm(C) + GELU that follows is norm.cu's LN_GELU forward.
namespace {

constexpr int C0B_ROWS = 128;      // output frames per CTA

__global__ void __launch_bounds__(256)
conv0_bias_kernel(const float* __restrict__ x, const long long* __restrict__ samp_off, const int* __restrict__ L0,
                  const long long* __restrict__ out_off, const float* __restrict__ w, const float* __restrict__ bias,
                  bf16* __restrict__ out, int C, int k, int stride, long long w_stride, long long b_stride) {
  extern __shared__ float wsm[];                 // [k][C] taps (transposed: a channel octet reads 8 consecutive floats), then [C] bias
  const int u = blockIdx.y;
  w += (long long)u * w_stride;
  if (bias) bias += (long long)u * b_stride;
  const int L = L0[u];
  const int t0 = blockIdx.x * C0B_ROWS;
  if (t0 >= L) return;
  for (int i = threadIdx.x; i < C * k; i += blockDim.x) wsm[(i % k) * C + i / k] = w[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) wsm[k * C + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int n_oct = C >> 3;
  const int oct = threadIdx.x % n_oct, rsub = threadIdx.x / n_oct, rstep = blockDim.x / n_oct;
  const float* xu = x + samp_off[u];
  const int t1 = min(L, t0 + C0B_ROWS);
  for (int t = t0 + rsub; t < t1; t += rstep) {
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = wsm[k * C + oct * 8 + c];
    for (int j = 0; j < k; ++j) {
      const float xv = __ldg(xu + (long long)t * stride + j);
      const float4 w0 = *reinterpret_cast<const float4*>(wsm + j * C + oct * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(wsm + j * C + oct * 8 + 4);
      acc[0] = fmaf(w0.x, xv, acc[0]); acc[1] = fmaf(w0.y, xv, acc[1]); acc[2] = fmaf(w0.z, xv, acc[2]); acc[3] = fmaf(w0.w, xv, acc[3]);
      acc[4] = fmaf(w1.x, xv, acc[4]); acc[5] = fmaf(w1.y, xv, acc[5]); acc[6] = fmaf(w1.z, xv, acc[6]); acc[7] = fmaf(w1.w, xv, acc[7]);
    }
    *reinterpret_cast<uint4*>(out + (out_off[u] + t) * C + oct * 8) =
        make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
  }
}

// The usual shape (k = 10): every thread keeps the K taps + bias of its 8 channels in registers and walks rows; the x
// window of a row is one broadcast load per tap for the 32 lanes (= 32 channel octets) of a warp.  80 FMA per 16-byte
// store: about the HBM time of the output at C = 512.
template <int K>
__global__ void __launch_bounds__(256)
conv0_bias_reg_kernel(const float* __restrict__ x, const long long* __restrict__ samp_off, const int* __restrict__ L0,
                      const long long* __restrict__ out_off, const float* __restrict__ w, const float* __restrict__ bias,
                      bf16* __restrict__ out, int C, int stride, int rows_per_cta, long long w_stride, long long b_stride) {
  const int u = blockIdx.y;
  w += (long long)u * w_stride;
  if (bias) bias += (long long)u * b_stride;
  const int L = L0[u];
  const int t0 = blockIdx.x * rows_per_cta;
  if (t0 >= L) return;
  const int n_oct = C >> 3;
  const int oct = threadIdx.x % n_oct, rsub = threadIdx.x / n_oct, rstep = blockDim.x / n_oct;
  float wr[8][K], br[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    br[c] = bias ? __ldg(bias + oct * 8 + c) : 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) wr[c][j] = __ldg(w + (oct * 8 + c) * K + j);
  }
  const float* xu = x + samp_off[u];
  bf16* ou = out + out_off[u] * C + oct * 8;
  const int t1 = min(L, t0 + rows_per_cta);
  for (int t = t0 + rsub; t < t1; t += rstep) {
    float xv[K];
#pragma unroll
    for (int j = 0; j < K; ++j) xv[j] = __ldg(xu + (long long)t * stride + j);
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      acc[c] = br[c];
#pragma unroll
      for (int j = 0; j < K; ++j) acc[c] = fmaf(wr[c][j], xv[j], acc[c]);
    }
    *reinterpret_cast<uint4*>(ou + (long long)t * C) =
        make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
  }
}

// row -> utterance table of a conv layer's row space (gap rows keep the -1 of the caller's memset)
__global__ void fill_row_utt_kernel(int* __restrict__ row_utt, const long long* __restrict__ off, const int* __restrict__ L) {
  const int u = blockIdx.y;
  const long long o = off[u];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < L[u]; t += gridDim.x * blockDim.x) row_utt[o + t] = u;
}

}  // namespace

int conv0_bias(const float* x, const long long* samp_off, const int* L0, const long long* out_off, const float* w,
               const float* bias, bf16* out, int n_utts, int C, int k, int stride, int max_L0, cudaStream_t stream,
               long long w_stride, long long b_stride) {
  SUTA_CHECK_ARG(C % 8 == 0 && C <= 2048 && 256 % (C / 8) == 0 && k > 0 && k <= 32);
  if (k == 10) {
    // ~4 CTAs per SM and utterance column; at least 256 rows per CTA so the 88 register-resident weights amortise
    int rows = ceil_div(max_L0, 4 * 148);
    rows = rows < 256 ? 256 : (rows + 63) / 64 * 64;
    conv0_bias_reg_kernel<10><<<dim3((unsigned)ceil_div(max_L0, rows), (unsigned)n_utts), 256, 0, stream>>>(
        x, samp_off, L0, out_off, w, bias, out, C, stride, rows, w_stride, b_stride);
    CUDA_TRY(cudaGetLastError());
    return SUTA_OK;
  }
  const size_t smem = sizeof(float) * ((size_t)k * C + C);
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    CUDA_TRY(cudaFuncSetAttribute(conv0_bias_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  conv0_bias_kernel<<<dim3((unsigned)ceil_div(max_L0, C0B_ROWS), (unsigned)n_utts), 256, smem, stream>>>(
      x, samp_off, L0, out_off, w, bias, out, C, k, stride, w_stride, b_stride);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

int fill_row_utt(int* row_utt, long long rows, const long long* off, const int* L, int n_utts, int max_L, cudaStream_t stream) {
  CUDA_TRY(cudaMemsetAsync(row_utt, 0xFF, sizeof(int) * (size_t)rows, stream));
  const int gx = max_L / 1024 + 1;
  fill_row_utt_kernel<<<dim3((unsigned)(gx > 64 ? 64 : gx), (unsigned)n_utts), 256, 0, stream>>>(row_utt, off, L);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}
