// Fused LayerNorm forward / backward with PER-UTTERANCE affine parameters.
// Restates torch native_layer_norm(+backward) as used by HF/modeling_wav2vec2.py:429-434,692,599-602 for a
// token-packed batch in which every utterance carries its own adapted gamma/beta (REF/main.py:81-87).
// One warp per row; row statistics by warp shuffle; dgamma/dbeta accumulated per utterance:
// registers -> shared memory -> one atomicAdd per column per CTA.  HBM-bound.
#include "kernels.cuh"

namespace {

template <typename T>
struct Loader;
template <>
struct Loader<float> {
  static __device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ float ld1(const float* p) { return *p; }
};
template <>
struct Loader<bf16> {
  static __device__ __forceinline__ float4 ld4(const bf16* p) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ float ld1(const bf16* p) { return __bfloat162float(*p); }
};

// Each lane owns columns {4*lane + 128*i + j}: NV = N/128 float4 groups (N % 128 == 0), or for small N
// (64) NV=1 with only the first N/4 lanes active.
template <int N>
struct Cols {
  static constexpr int NV = (N + 127) / 128;
  static __device__ __forceinline__ bool active(int lane, int i) { return 4 * lane + 128 * i < N; }
};

template <int N, typename TIn>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const TIn* __restrict__ x, const int* __restrict__ row_utt, const float* __restrict__ P, long long pstride,
              int g_off, int b_off, float* __restrict__ y32, bf16* __restrict__ y16, float* __restrict__ mean_out,
              float* __restrict__ rstd_out, long long M, float eps) {
  constexpr int NV = Cols<N>::NV;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const TIn* xr = x + row * N;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (Cols<N>::active(lane, i)) {
      v[i] = Loader<TIn>::ld4(xr + 4 * lane + 128 * i);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float mean = warp_sum(s) * (1.0f / N);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    if (Cols<N>::active(lane, i)) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      ss += a * a + b * b + c * c + d * d;
    }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / N) + eps);
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
  const float* gam = P + (long long)row_utt[row] * pstride + g_off;
  const float* bet = P + (long long)row_utt[row] * pstride + b_off;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    if (Cols<N>::active(lane, i)) {
      const int col = 4 * lane + 128 * i;
      float4 g = *reinterpret_cast<const float4*>(gam + col), b = *reinterpret_cast<const float4*>(bet + col);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + b.x;
      o.y = (v[i].y - mean) * rstd * g.y + b.y;
      o.z = (v[i].z - mean) * rstd * g.z + b.z;
      o.w = (v[i].w - mean) * rstd * g.w + b.w;
      if (y32) *reinterpret_cast<float4*>(y32 + row * N + col) = o;
      if (y16) *reinterpret_cast<uint2*>(y16 + row * N + col) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    }
}

constexpr int BWD_ROWS_PER_WARP = 8;
constexpr int BWD_WARPS = 8;

template <int N, typename TIn>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* __restrict__ dy, const TIn* __restrict__ x, const float* __restrict__ mean,
              const float* __restrict__ rstd, const int* __restrict__ row_utt, const float* __restrict__ P,
              long long pstride, int g_off, int b_off, float* __restrict__ G, float* __restrict__ dx32,
              bf16* __restrict__ dx16, long long M) {
  constexpr int NV = Cols<N>::NV;
  __shared__ float s_dg[N];
  __shared__ float s_db[N];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    s_dg[i] = 0.f;
    s_db[i] = 0.f;
  }
  __syncthreads();
  const long long cta_row0 = (long long)blockIdx.x * (BWD_WARPS * BWD_ROWS_PER_WARP);
  const int u_cta = row_utt[cta_row0 < M ? cta_row0 : M - 1];
  const long long r0 = cta_row0 + warp * BWD_ROWS_PER_WARP;

  float4 ag[NV], ab[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  int u_cur = -1;

  auto flush = [&](int u) {
    if (u < 0) return;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (Cols<N>::active(lane, i)) {
        const int col = 4 * lane + 128 * i;
        if (u == u_cta) {
          atomicAdd(&s_dg[col + 0], ag[i].x); atomicAdd(&s_dg[col + 1], ag[i].y);
          atomicAdd(&s_dg[col + 2], ag[i].z); atomicAdd(&s_dg[col + 3], ag[i].w);
          atomicAdd(&s_db[col + 0], ab[i].x); atomicAdd(&s_db[col + 1], ab[i].y);
          atomicAdd(&s_db[col + 2], ab[i].z); atomicAdd(&s_db[col + 3], ab[i].w);
        } else if (G) {
          float* gg = G + (long long)u * pstride + g_off + col;
          float* gb = G + (long long)u * pstride + b_off + col;
          atomicAdd(gg + 0, ag[i].x); atomicAdd(gg + 1, ag[i].y); atomicAdd(gg + 2, ag[i].z); atomicAdd(gg + 3, ag[i].w);
          atomicAdd(gb + 0, ab[i].x); atomicAdd(gb + 1, ab[i].y); atomicAdd(gb + 2, ab[i].z); atomicAdd(gb + 3, ab[i].w);
        }
        ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
  };

  for (int rr = 0; rr < BWD_ROWS_PER_WARP; ++rr) {
    const long long row = r0 + rr;
    if (row >= M) break;
    const int u = row_utt[row];
    if (u != u_cur) {
      flush(u_cur);
      u_cur = u;
    }
    const float mu = mean[row], rs = rstd[row];
    const float* gam = P + (long long)u * pstride + g_off;
    float4 xh[NV], dxh[NV];
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (Cols<N>::active(lane, i)) {
        const int col = 4 * lane + 128 * i;
        float4 xv = Loader<TIn>::ld4(x + row * N + col);
        float4 d = *reinterpret_cast<const float4*>(dy + row * N + col);
        float4 g = *reinterpret_cast<const float4*>(gam + col);
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        ag[i].x += d.x * xh[i].x; ag[i].y += d.y * xh[i].y; ag[i].z += d.z * xh[i].z; ag[i].w += d.w * xh[i].w;
        ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
        dxh[i] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
        c1 += dxh[i].x + dxh[i].y + dxh[i].z + dxh[i].w;
        c2 += dxh[i].x * xh[i].x + dxh[i].y * xh[i].y + dxh[i].z * xh[i].z + dxh[i].w * xh[i].w;
      }
    if (dx32 || dx16) {
      c1 = warp_sum(c1) * (1.0f / N);
      c2 = warp_sum(c2) * (1.0f / N);
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (Cols<N>::active(lane, i)) {
          const int col = 4 * lane + 128 * i;
          float4 o;
          o.x = rs * (dxh[i].x - c1 - xh[i].x * c2);
          o.y = rs * (dxh[i].y - c1 - xh[i].y * c2);
          o.z = rs * (dxh[i].z - c1 - xh[i].z * c2);
          o.w = rs * (dxh[i].w - c1 - xh[i].w * c2);
          if (dx32) *reinterpret_cast<float4*>(dx32 + row * N + col) = o;
          if (dx16) *reinterpret_cast<uint2*>(dx16 + row * N + col) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
        }
    }
  }
  flush(u_cur);
  __syncthreads();
  if (G) {
    float* gg = G + (long long)u_cta * pstride + g_off;
    float* gb = G + (long long)u_cta * pstride + b_off;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      float a = s_dg[i], b = s_db[i];
      if (a != 0.f) atomicAdd(gg + i, a);
      if (b != 0.f) atomicAdd(gb + i, b);
    }
  }
}

template <int N, typename TIn>
int launch_fwd(const TIn* x, const int* row_utt, UttParams prm, int g_off, int b_off, float* y32, bf16* y16, float* mean,
               float* rstd, long long M, float eps, cudaStream_t stream) {
  const int warps = 8;
  ln_fwd_kernel<N, TIn><<<(unsigned)((M + warps - 1) / warps), warps * 32, 0, stream>>>(
      x, row_utt, prm.P, prm.stride, g_off, b_off, y32, y16, mean, rstd, M, eps);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}
template <int N, typename TIn>
int launch_bwd(const float* dy, const TIn* x, const float* mean, const float* rstd, const int* row_utt, UttParams prm,
               int g_off, int b_off, float* G, float* dx32, bf16* dx16, long long M, cudaStream_t stream) {
  const int rows = BWD_WARPS * BWD_ROWS_PER_WARP;
  ln_bwd_kernel<N, TIn><<<(unsigned)((M + rows - 1) / rows), BWD_WARPS * 32, 0, stream>>>(
      dy, x, mean, rstd, row_utt, prm.P, prm.stride, g_off, b_off, G, dx32, dx16, M);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

}  // namespace

#define LN_DISPATCH(N_, CALL)                                                     \
  switch (N_) {                                                                   \
    case 64: { constexpr int NN = 64; CALL; } break;                              \
    case 128: { constexpr int NN = 128; CALL; } break;                            \
    case 256: { constexpr int NN = 256; CALL; } break;                            \
    case 512: { constexpr int NN = 512; CALL; } break;                            \
    case 768: { constexpr int NN = 768; CALL; } break;                            \
    case 1024: { constexpr int NN = 1024; CALL; } break;                          \
    default:                                                                      \
      suta_set_last_error("layernorm: unsupported width %d", N_);                 \
      return SUTA_ERR_ARG;                                                        \
  }

int layernorm_forward(const float* x_f32, const bf16* x_bf16, const int* row_utt, UttParams prm, int g_off, int b_off,
                      float* y_f32, bf16* y_bf16, float* mean, float* rstd, long long M, int N, float eps,
                      cudaStream_t stream) {
  SUTA_CHECK_ARG((x_f32 != nullptr) != (x_bf16 != nullptr));
  SUTA_CHECK_ARG(g_off % 4 == 0 && b_off % 4 == 0 && prm.stride % 4 == 0);
  if (M <= 0) return SUTA_OK;
  if (x_f32) {
    LN_DISPATCH(N, return (launch_fwd<NN, float>(x_f32, row_utt, prm, g_off, b_off, y_f32, y_bf16, mean, rstd, M, eps, stream)));
  } else {
    LN_DISPATCH(N, return (launch_fwd<NN, bf16>(x_bf16, row_utt, prm, g_off, b_off, y_f32, y_bf16, mean, rstd, M, eps, stream)));
  }
  return SUTA_OK;
}

int layernorm_backward(const float* dy, const float* x_f32, const bf16* x_bf16, const float* mean, const float* rstd,
                       const int* row_utt, UttParams prm, int g_off, int b_off, float* G, float* dx_f32, bf16* dx_bf16,
                       long long M, int N, cudaStream_t stream) {
  SUTA_CHECK_ARG((x_f32 != nullptr) != (x_bf16 != nullptr));
  if (M <= 0) return SUTA_OK;
  if (x_f32) {
    LN_DISPATCH(N, return (launch_bwd<NN, float>(dy, x_f32, mean, rstd, row_utt, prm, g_off, b_off, G, dx_f32, dx_bf16, M, stream)));
  } else {
    LN_DISPATCH(N, return (launch_bwd<NN, bf16>(dy, x_bf16, mean, rstd, row_utt, prm, g_off, b_off, G, dx_f32, dx_bf16, M, stream)));
  }
  return SUTA_OK;
}
