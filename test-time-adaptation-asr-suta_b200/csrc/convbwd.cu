// Backward pieces of the CNN feature extractor for --train_feature (REF/main.py:88-94 makes every parameter of
// feature_extractor / feature_projection trainable, so autograd reaches the waveform-side layers):
//   cast_params_bf16        refresh the bf16 GEMM-operand copy of per-utterance trainable weights after an update
//   conv_col2im_gelu_grad   gather the dgrad GEMM result Z[t,(j,ci)] back to input rows r = s*t + j and multiply by
//                           GELU'(pre-activation of the layer below), which the forward saved  (backward of HF:269-272)
//   gelu_grad_to_padded     dX * GELU' written into a per-utterance 128-row-aligned slab (zero rows between
//                           utterances) so the weight-gradient GEMM can reduce over time in 64-row steps
//   conv0_groupnorm_backward  backward of Conv1d(1->C,k,s) + GroupNorm(per channel over time)  (HF:319-323)
//   colsum_per_utt          bias gradient of the per-utterance projection
// The contractions themselves (dgrad, wgrad) run on the tcgen05 GEMM with MN-major operands (gemm_tc.cu).
#include "kernels.cuh"

namespace {

__global__ void cast_kernel(const float* __restrict__ P, long long pstride, long long seg_off, long long size, int U,
                            bf16* __restrict__ out) {
  const long long total = size * U;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long u = i / size, e = i - u * size;
    out[i] = __float2bfloat16(P[u * pstride + seg_off + e]);
  }
}

__global__ void gelu_grad_pad_kernel(const float* __restrict__ d, const bf16* __restrict__ pre, bf16* __restrict__ out,
                                     const int* __restrict__ row_utt, const long long* __restrict__ tok_off,
                                     const long long* __restrict__ pad_off, long long M, int C) {
  const int c8 = C >> 3;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * c8) return;
  const long long row = idx / c8;
  const int c = (int)(idx - row * c8) * 8;
  const int u = row_utt[row];
  const long long prow = pad_off[u] + (row - tok_off[u]);
  float4 a = *reinterpret_cast<const float4*>(d + row * C + c), b = *reinterpret_cast<const float4*>(d + row * C + c + 4);
  if (pre) {
    uint4 p = *reinterpret_cast<const uint4*>(pre + row * C + c);
    float2 p0 = unpack_bf16x2(p.x), p1 = unpack_bf16x2(p.y), p2 = unpack_bf16x2(p.z), p3 = unpack_bf16x2(p.w);
    a.x *= p0.x; a.y *= p0.y; a.z *= p1.x; a.w *= p1.y;
    b.x *= p2.x; b.y *= p2.y; b.z *= p3.x; b.w *= p3.y;
  }
  *reinterpret_cast<uint4*>(out + prow * C + c) =
      make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
}

constexpr int C2I_ROWS = 16;      // rows per CTA: 4 per thread, all loads of the 4 rows issued before the first use
constexpr int C2I_MAXTAPS = 2;    // a row receives ceil(k / s) <= 2 taps for the wav2vec2 feature extractor (k <= 2 s)

__global__ void __launch_bounds__(256)
col2im_kernel(Col2imArgs a) {
  const int u = blockIdx.y;
  const int Lo = a.L_out[u], Li = a.L_in[u];
  if (blockIdx.x * C2I_ROWS >= Lo) return;
  const int c8n = a.C >> 3;                       // 16-byte channel groups per row (64 for C = 512)
  const long long oo = a.off_out[u], oi = a.off_in[u];
  const int ldz = a.k * a.C;
  const bf16* __restrict__ Z = a.Z;
  const bf16* __restrict__ pre = a.pre;
  bf16* __restrict__ out = a.out;
  const int rows_per_pass = 256 / c8n;            // 4 for C = 512
  const int c = (threadIdx.x % c8n) * 8;
  const int rsub = threadIdx.x / c8n;
  constexpr int NR = 4;
  uint4 z[NR][C2I_MAXTAPS], pv[NR];
  bool zok[NR][C2I_MAXTAPS], rok[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    const int r = blockIdx.x * C2I_ROWS + rsub + rows_per_pass * i;
    rok[i] = r < Lo && rsub + rows_per_pass * i < C2I_ROWS;
    // taps j = (r mod s), (r mod s) + s, ... < k  read Z row t = (r - j) / s
    const int j0 = r % a.s;
#pragma unroll
    for (int q = 0; q < C2I_MAXTAPS; ++q) {
      const int j = j0 + q * a.s;
      const int t = (r - j) / a.s;
      zok[i][q] = rok[i] && j < a.k && r - j >= 0 && t < Li;
      z[i][q] = make_uint4(0u, 0u, 0u, 0u);
      if (zok[i][q]) z[i][q] = __ldg(reinterpret_cast<const uint4*>(Z + (oi + t) * ldz + j * a.C + c));
    }
    pv[i] = make_uint4(0u, 0u, 0u, 0u);
    if (rok[i]) pv[i] = __ldg(reinterpret_cast<const uint4*>(pre + (oo + r) * a.C + c));
  }
#pragma unroll
  for (int i = 0; i < NR; ++i) {
    if (!rok[i]) continue;
    const int r = blockIdx.x * C2I_ROWS + rsub + rows_per_pass * i;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
    for (int q = 0; q < C2I_MAXTAPS; ++q) {
      const float2 z0 = unpack_bf16x2(z[i][q].x), z1 = unpack_bf16x2(z[i][q].y), z2 = unpack_bf16x2(z[i][q].z), z3 = unpack_bf16x2(z[i][q].w);
      acc[0] += z0.x; acc[1] += z0.y; acc[2] += z1.x; acc[3] += z1.y;
      acc[4] += z2.x; acc[5] += z2.y; acc[6] += z3.x; acc[7] += z3.y;
    }
    const float2 p0 = unpack_bf16x2(pv[i].x), p1 = unpack_bf16x2(pv[i].y), p2 = unpack_bf16x2(pv[i].z), p3 = unpack_bf16x2(pv[i].w);
    acc[0] *= p0.x; acc[1] *= p0.y; acc[2] *= p1.x; acc[3] *= p1.y;
    acc[4] *= p2.x; acc[5] *= p2.y; acc[6] *= p3.x; acc[7] *= p3.y;
    *reinterpret_cast<uint4*>(out + (oo + r) * a.C + c) =
        make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
  }
}

// ---- conv0 + GroupNorm backward -------------------------------------------------------------------------------
// y = gamma * zhat + beta, zhat = (z - mu) * rstd, z[t] = sum_j w[j] x[s t + j]; given dy (already times GELU'):
//   dbeta = S1 = sum_t dy,  dgamma = S2 = sum_t dy zhat,
//   dz[t] = rstd*gamma*(dy[t] - S1/L - zhat[t]*S2/L),  dw[j] = sum_t dz[t] x[s t + j]
//         = rstd*gamma*(A_j - (S1/L) B_j - (S2/L) C_j),  A_j = sum dy x, B_j = sum x, C_j = sum zhat x.
// Only S1 and A_j need a pass over dy (a [C x (k+1)] = dy^T [X | 1] reduction over time, HBM-bound on reading dy);
// the rest follows from the audio moments Sx, R (frontend.cu):  B_j = Sx[j],
//   sum_t z x_j = sum_j' w_j' R[j'][j]  =>  C_j = rstd (sum_j' w_j' R[j'][j] - mu Sx[j]),   S2 = rstd (sum_j w_j A_j - mu S1).
constexpr int C0B_TT = 1024;
constexpr int C0B_MAXK = 16;

__global__ void __launch_bounds__(256)
conv0_bwd_accum_kernel(Conv0BwdArgs a) {
  extern __shared__ float sx[];
  const int u = blockIdx.y;
  const int L0 = a.L0[u];
  const int t0 = blockIdx.x * C0B_TT;
  if (t0 >= L0) return;
  const int nt = min(C0B_TT, L0 - t0);
  const float* x = a.x + a.samp_off[u] + (long long)t0 * a.stride;
  const int nx = (nt - 1) * a.stride + a.k;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) sx[i] = x[i];
  __syncthreads();
  const long long row0 = a.out_off[u] + t0;
  const int k = a.k;
  for (int cp = threadIdx.x; 2 * cp < a.C; cp += blockDim.x) {
    const int c = 2 * cp;
    // the two channels of the pair ride in packed fp32 registers (FFMA2 with the window sample as the broadcast operand):
    // the kernel is bound by FMA-pipe instructions, not by memory
    uint64_t s1 = 0ull, A2[C0B_MAXK];
#pragma unroll
    for (int j = 0; j < C0B_MAXK; ++j) A2[j] = 0ull;
    const bf16* dyp = a.dy + row0 * a.C + c;
    int t = 0;
    const bool vec_ok = a.stride == 5 && k == 10;    // the wav2vec2 front end: 4 frames = 20 samples = five 16-byte words
    // eight independent 4-byte loads in flight per thread: the kernel is latency-bound (IPC ~0.2 per sub-partition with four),
    // and by Little's law 2048 threads x 16 B barely cover the ~35 KB per SM that 6.5 TB/s x 800 ns call for
    for (; vec_ok && t + 8 <= nt; t += 8) {
      unsigned int rr[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) rr[q] = __ldg(reinterpret_cast<const unsigned int*>(dyp + (long long)(t + q) * a.C));
      // the 8 windows cover samples [5 t, 5 t + 45): twelve 16-byte broadcast loads instead of 80 scalar ones
      float xw[48];
#pragma unroll
      for (int v = 0; v < 12; ++v) {
        const float4 f = *reinterpret_cast<const float4*>(sx + 5 * t + 4 * v);
        xw[4 * v] = f.x; xw[4 * v + 1] = f.y; xw[4 * v + 2] = f.z; xw[4 * v + 3] = f.w;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float2 dy = unpack_bf16x2(rr[q]);
        const uint64_t dy2 = pk2(dy.x, dy.y);
        s1 = add2(s1, dy2);
#pragma unroll
        for (int j = 0; j < 10; ++j) A2[j] = fma2(dy2, dup2(xw[5 * q + j]), A2[j]);
      }
    }
    for (; vec_ok && t + 4 <= nt; t += 4) {
      const unsigned int r0 = __ldg(reinterpret_cast<const unsigned int*>(dyp + (long long)(t + 0) * a.C));
      const unsigned int r1 = __ldg(reinterpret_cast<const unsigned int*>(dyp + (long long)(t + 1) * a.C));
      const unsigned int r2 = __ldg(reinterpret_cast<const unsigned int*>(dyp + (long long)(t + 2) * a.C));
      const unsigned int r3 = __ldg(reinterpret_cast<const unsigned int*>(dyp + (long long)(t + 3) * a.C));
      const unsigned int rr[4] = {r0, r1, r2, r3};
      // the 4 windows cover samples [5 t, 5 t + 25): seven 16-byte broadcast loads instead of 40 scalar ones
      float xw[28];
#pragma unroll
      for (int v = 0; v < 7; ++v) {
        const float4 f = *reinterpret_cast<const float4*>(sx + 5 * t + 4 * v);
        xw[4 * v] = f.x; xw[4 * v + 1] = f.y; xw[4 * v + 2] = f.z; xw[4 * v + 3] = f.w;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 dy = unpack_bf16x2(rr[q]);
        const uint64_t dy2 = pk2(dy.x, dy.y);
        s1 = add2(s1, dy2);
#pragma unroll
        for (int j = 0; j < 10; ++j) A2[j] = fma2(dy2, dup2(xw[5 * q + j]), A2[j]);
      }
    }
    for (; t < nt; ++t) {
      const float2 dy = unpack_bf16x2(__ldg(reinterpret_cast<const unsigned int*>(dyp + (long long)t * a.C)));
      const float* xs = sx + t * a.stride;
      const uint64_t dy2 = pk2(dy.x, dy.y);
      s1 = add2(s1, dy2);
#pragma unroll
      for (int j = 0; j < C0B_MAXK; ++j)
        if (j < k) A2[j] = fma2(dy2, dup2(xs[j]), A2[j]);
    }
    // partial sums, layout [U][chunk][k + 1][C]: channel-contiguous, so this store and the finalize kernel's loads coalesce
    float* pa = a.part + ((long long)u * a.n_chunk + blockIdx.x) * (k + 1) * a.C + c;
    float lo, hi;
    upk2(s1, lo, hi);
    *reinterpret_cast<float2*>(pa) = make_float2(lo, hi);
#pragma unroll
    for (int j = 0; j < C0B_MAXK; ++j)           // static indices: a runtime-indexed copy would put the accumulators in local memory
      if (j < k) {
        upk2(A2[j], lo, hi);
        *reinterpret_cast<float2*>(pa + (long long)(1 + j) * a.C) = make_float2(lo, hi);
      }
  }
}

__global__ void conv0_bwd_finalize_kernel(Conv0BwdArgs a) {
  const int u = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C) return;
  const int k = a.k;
  const int L0 = a.L0[u];
  const int nchunk = (L0 + C0B_TT - 1) / C0B_TT;
  double S1 = 0.0, A[C0B_MAXK];
  for (int j = 0; j < C0B_MAXK; ++j) A[j] = 0.0;
  for (int ch = 0; ch < nchunk; ++ch) {
    const float* pp = a.part + ((long long)u * a.n_chunk + ch) * (k + 1) * a.C + c;
    S1 += (double)pp[0];
#pragma unroll
    for (int j = 0; j < C0B_MAXK; ++j)
      if (j < k) A[j] += (double)pp[(long long)(1 + j) * a.C];
  }
  if (a.plain) {                                       // conv + bias only: the accumulated sums ARE the gradients
    float* Gp = a.G + (long long)u * a.pstride;
    Gp[a.b_off + c] = (float)S1;
    for (int j = 0; j < k; ++j) Gp[a.w_off + (long long)c * k + j] = (float)A[j];
    return;
  }
  const double* st = a.stats + ((long long)u * a.C + c) * 2;
  const double mu = st[0] / L0;
  const double rstd = 1.0 / sqrt(st[1] / L0 - mu * mu + 1e-5);
  const int npair = k + k * (k + 1) / 2;
  const double* mom = a.mom + (long long)u * npair;
  const float* w = a.w + (long long)u * a.w_stride + (long long)c * k;
  double wj[C0B_MAXK];
  for (int j = 0; j < C0B_MAXK; ++j) wj[j] = j < k ? (double)w[j] : 0.0;
  double wa = 0.0;
  for (int j = 0; j < k; ++j) wa += wj[j] * A[j];
  const double S2 = rstd * (wa - mu * S1);
  const double gamma = a.P[(long long)u * a.pstride + a.g_off + c];
  float* G = a.G + (long long)u * a.pstride;
  G[a.b_off + c] = (float)S1;
  G[a.g_off + c] = (float)S2;
  for (int j = 0; j < k; ++j) {
    double wr = 0.0;                                   // sum_j' w_j' R[j'][j]
    for (int j2 = 0; j2 < k; ++j2) {
      const int lo = j2 < j ? j2 : j, hi = j2 < j ? j : j2;
      wr += wj[j2] * mom[k + lo * k - lo * (lo - 1) / 2 + (hi - lo)];
    }
    const double Cj = rstd * (wr - mu * mom[j]);
    G[a.w_off + (long long)c * k + j] = (float)(rstd * gamma * (A[j] - S1 / L0 * mom[j] - S2 / L0 * Cj));
  }
}

__global__ void colsum_kernel(const float* __restrict__ x, const long long* __restrict__ tok_off, const int* __restrict__ T,
                              float* __restrict__ G, long long gstride, long long g_off, int C) {
  const int u = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float* p = x + tok_off[u] * C + c;
  float s = 0.f;
  for (int t = 0; t < T[u]; ++t) s += p[(long long)t * C];
  G[(long long)u * gstride + g_off + c] = s;
}

// 32 row lanes x 8 column pairs per CTA; every lane sums its rows in order, the lanes are added in order: same bits every
// run.  (8 lanes x 32 pairs before: with ONE utterance of ~300 rows per launch -- the bias gradients of train_all -- a 12-CTA
// grid walked 33 dependent loads per lane, 8.6 us per launch and 9 % of a train_all step; a warp still reads whole 32-byte sectors)
constexpr int CS_LANES = 32, CS_PAIRS = 8;
__global__ void __launch_bounds__(CS_LANES * CS_PAIRS)
colsum_bf16_kernel(const bf16* __restrict__ x, const long long* __restrict__ row_off, const int* __restrict__ L,
                   float* __restrict__ G, long long gstride, long long g_off, int C) {
  __shared__ float2 red[CS_LANES][CS_PAIRS];
  const int u = blockIdx.y;
  const int cp = threadIdx.x & (CS_PAIRS - 1), lane_r = threadIdx.x / CS_PAIRS;
  const int c = (blockIdx.x * CS_PAIRS + cp) * 2;
  float2 s = make_float2(0.f, 0.f);
  if (c < C) {
    const bf16* p = x + row_off[u] * C + c;
    const int n = L[u];
    for (int t = lane_r; t < n; t += CS_LANES) {
      const float2 v = unpack_bf16x2(__ldg(reinterpret_cast<const unsigned int*>(p + (long long)t * C)));
      s.x += v.x; s.y += v.y;
    }
  }
  red[lane_r][cp] = s;
  __syncthreads();
  if (lane_r == 0 && c < C) {
    float2 a = red[0][cp];
    for (int r = 1; r < CS_LANES; ++r) { a.x += red[r][cp].x; a.y += red[r][cp].y; }
    *reinterpret_cast<float2*>(G + (long long)u * gstride + g_off + c) = a;
  }
}

int grid_for(long long total) {
  long long b = (total + 255) / 256, cap = 148LL * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace

int cast_params_bf16(const float* P, long long pstride, long long seg_off, long long size, int n_utts, bf16* out,
                     cudaStream_t stream) {
  SUTA_CHECK_ARG(P && out && size > 0 && n_utts > 0);
  cast_kernel<<<grid_for(size * n_utts), 256, 0, stream>>>(P, pstride, seg_off, size, n_utts, out);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

int gelu_grad_to_padded(const float* d, const bf16* pre, bf16* out, const int* row_utt, const long long* tok_off,
                        const long long* pad_off, long long M, int C, cudaStream_t stream) {
  SUTA_CHECK_ARG(C % 8 == 0);
  long long n = M * (C >> 3);
  if (n <= 0) return SUTA_OK;
  gelu_grad_pad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d, pre, out, row_utt, tok_off, pad_off, M, C);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

int conv_col2im_gelu_grad(const Col2imArgs& a, cudaStream_t stream) {
  SUTA_CHECK_ARG(a.C % 8 == 0 && a.n_utts > 0 && a.max_L_out > 0);
  SUTA_CHECK_ARG(256 % (a.C >> 3) == 0 && (256 / (a.C >> 3)) * 4 >= C2I_ROWS && a.k <= C2I_MAXTAPS * a.s);
  dim3 grid(ceil_div(a.max_L_out, C2I_ROWS), a.n_utts);
  col2im_kernel<<<grid, 256, 0, stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

int conv0_groupnorm_backward(const Conv0BwdArgs& a, cudaStream_t stream) {
  SUTA_CHECK_ARG(a.k <= C0B_MAXK && a.n_utts > 0 && a.C % 2 == 0 && (a.mom || a.plain) && a.part);
  SUTA_CHECK_ARG(a.n_chunk >= ceil_div(a.max_L0, C0B_TT));
  dim3 grid(ceil_div(a.max_L0, C0B_TT), a.n_utts);
  size_t smem = sizeof(float) * (C0B_TT * a.stride + a.k + 32);      // slack for the 16-byte window loads
  conv0_bwd_accum_kernel<<<grid, 256, smem, stream>>>(a);
  conv0_bwd_finalize_kernel<<<dim3(ceil_div(a.C, 128), a.n_utts), 128, 0, stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

int colsum_per_utt(const float* x, const long long* tok_off, const int* T, float* G, long long gstride, long long g_off,
                   int C, int n_utts, cudaStream_t stream) {
  colsum_kernel<<<dim3(ceil_div(C, 128), n_utts), 128, 0, stream>>>(x, tok_off, T, G, gstride, g_off, C);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

int colsum_per_utt_bf16(const bf16* x, const long long* row_off, const int* L, float* G, long long gstride, long long g_off,
                        int C, int n_utts, cudaStream_t stream) {
  SUTA_CHECK_ARG(C % 2 == 0 && g_off % 2 == 0 && gstride % 2 == 0);
  colsum_bf16_kernel<<<dim3(ceil_div(C, 2 * CS_PAIRS), n_utts), CS_LANES * CS_PAIRS, 0, stream>>>(x, row_off, L, G, gstride, g_off, C);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

long long conv0_bwd_scratch_floats(int n_utts, int C, int k, int max_L0) {
  return (long long)n_utts * ceil_div(max_L0, C0B_TT) * C * (k + 1);
}
int conv0_bwd_chunks(int max_L0) { return ceil_div(max_L0, C0B_TT); }
