// Kernels only --train_all needs (REF/main.py:96-100: every parameter of the model is adapted, SUTA_FLAG_TRAIN_ALL).
//   transpose_cast_bf16           W^T bf16 copy of an fp32 weight inside the trainable vector: the K-major B operand of the
//                                 dgrad GEMM dX = dY W (the forward's operand W itself is a plain bf16 cast of the vector)
//   posconv_weight_norm_forward   HF/modeling_wav2vec2.py:344-352: the positional conv's weight is parametrised as
//                                 w = v * g / ||v||  (torch weight_norm, dim = 2: one norm per tap over [H, H/G]); folds g and
//                                 v of the trainable vector into the two bf16 operand layouts posconv_tc / the GEMM read
//   posconv_weight_norm_backward  d g, d v from the gradient of the folded weight (torch _weight_norm_interface_backward)
// All reductions are two-stage in a fixed order (bit-reproducible).  Tiny next to the GEMMs: element-wise over 4.7 M weights.
#include "kernels.cuh"

namespace {

constexpr int WN_BLOCKS = 256;

// one launch for all the matrices of an update: block -> (job, 32 x 32 tile) through the jobs' running tile counts
__global__ void transpose_cast_kernel(TransposeJobs jobs) {
  __shared__ float tile[32][33];
  int j = 0;
  while (j + 1 < jobs.n && (int)blockIdx.x >= jobs.job[j + 1].tile0) ++j;
  const float* __restrict__ src = jobs.job[j].src;
  bf16* __restrict__ dst = jobs.job[j].dst;
  const int R = jobs.job[j].R, C = jobs.job[j].C;
  const int tx = (C + 31) / 32, t = blockIdx.x - jobs.job[j].tile0;
  const int c0 = (t % tx) * 32, r0 = (t / tx) * 32;
  for (int y = threadIdx.y; y < 32; y += blockDim.y) {
    const int r = r0 + y, c = c0 + threadIdx.x;
    tile[y][threadIdx.x] = (r < R && c < C) ? src[(long long)r * C + c] : 0.f;
  }
  __syncthreads();
  for (int y = threadIdx.y; y < 32; y += blockDim.y) {
    const int c = c0 + y, r = r0 + threadIdx.x;
    if (c < C && r < R) dst[(long long)c * R + r] = __float2bfloat16(tile[threadIdx.x][y]);
  }
}

// stage 1 of both reductions over the rows (o, i) of v [rows][K]: thread = tap k, block b sums its slice of the rows in order
//   MODE 0: part[b][k] = sum v^2        MODE 1: part[b][k] = sum dW[o][k][i] * v[o][i][k]
template <int MODE>
__global__ void wn_partial_kernel(const float* __restrict__ v, const float* __restrict__ dW, float* __restrict__ part, int rows,
                                  int CG, int K) {
  const int k = threadIdx.x;
  const int per = (rows + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float x = v[(long long)r * K + k];
    if (MODE == 0) {
      s = fmaf(x, x, s);
    } else {
      const int o = r / CG, i = r - o * CG;
      s = fmaf(dW[((long long)o * K + k) * CG + i], x, s);
    }
  }
  part[blockIdx.x * K + k] = s;
}

// stage 2, forward: norm[k] = sqrt(sum_b part[b][k]), scale[k] = g[k] / norm[k]
__global__ void wn_norm_kernel(const float* __restrict__ part, const float* __restrict__ g, float* __restrict__ norm,
                               float* __restrict__ scale, int nb, int K) {
  const int k = threadIdx.x;
  float s = 0.f;
  for (int b = 0; b < nb; ++b) s += part[b * K + k];
  const float n = sqrtf(s);
  norm[k] = n;
  scale[k] = g[k] / n;
}

// stage 3, forward: w[o][i][k] = v[o][i][k] * scale[k] into
//   w_fwd [co][(tap, ci)]                       (forward operand)
//   w_bwd [g*CG + ci][(K-1-tap, co_local)]      (dgrad operand: flipped taps, in/out channels swapped inside the group)
__global__ void wn_fold_kernel(const float* __restrict__ v, const float* __restrict__ scale, bf16* __restrict__ w_fwd,
                               bf16* __restrict__ w_bwd, long long total, int CG, int K) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int k = (int)(idx % K);
  const long long r = idx / K;
  const int o = (int)(r / CG), i = (int)(r - (long long)o * CG);
  const int grp = o / CG, ol = o - grp * CG;
  const bf16 w = __float2bfloat16(v[idx] * scale[k]);
  w_fwd[((long long)o * K + k) * CG + i] = w;
  w_bwd[((long long)(grp * CG + i) * K + (K - 1 - k)) * CG + ol] = w;
}

// stage 2, backward: dgn = (sum_b part[b][k]) / norm[k] = d g[k]
__global__ void wn_dg_kernel(const float* __restrict__ part, const float* __restrict__ norm, float* __restrict__ dg, int nb, int K) {
  const int k = threadIdx.x;
  float s = 0.f;
  for (int b = 0; b < nb; ++b) s += part[b * K + k];
  dg[k] = s / norm[k];
}

// stage 3, backward: d v[o][i][k] = scale[k] * (dW[o][k][i] - v[o][i][k] * dg[k] / norm[k])
__global__ void wn_dv_kernel(const float* __restrict__ v, const float* __restrict__ dW, const float* __restrict__ scale,
                             const float* __restrict__ norm, const float* __restrict__ dg, float* __restrict__ dv, long long total,
                             int CG, int K) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int k = (int)(idx % K);
  const long long r = idx / K;
  const int o = (int)(r / CG), i = (int)(r - (long long)o * CG);
  dv[idx] = scale[k] * (dW[((long long)o * K + k) * CG + i] - v[idx] * (dg[k] / norm[k]));
}

}  // namespace

int transpose_cast_bf16(TransposeJobs& jobs, cudaStream_t stream) {
  SUTA_CHECK_ARG(jobs.n > 0 && jobs.n <= TransposeJobs::MAX);
  int tiles = 0;
  for (int j = 0; j < jobs.n; ++j) {
    SUTA_CHECK_ARG(jobs.job[j].src && jobs.job[j].dst && jobs.job[j].R > 0 && jobs.job[j].C > 0);
    jobs.job[j].tile0 = tiles;
    tiles += ceil_div(jobs.job[j].C, 32) * ceil_div(jobs.job[j].R, 32);
  }
  transpose_cast_kernel<<<tiles, dim3(32, 8), 0, stream>>>(jobs);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

long long posconv_weight_norm_scratch_floats(int K) { return (long long)WN_BLOCKS * K + 2LL * K; }

int posconv_weight_norm_forward(const float* g, const float* v, float* scratch, bf16* w_fwd, bf16* w_bwd, int H, int CG, int K,
                                cudaStream_t stream) {
  SUTA_CHECK_ARG(g && v && scratch && w_fwd && w_bwd && K > 0 && K <= 1024 && H % CG == 0);
  const int rows = H * CG;
  const int nb = rows < WN_BLOCKS ? rows : WN_BLOCKS;
  float *part = scratch, *norm = scratch + (long long)WN_BLOCKS * K, *scale = norm + K;
  wn_partial_kernel<0><<<nb, K, 0, stream>>>(v, nullptr, part, rows, CG, K);
  wn_norm_kernel<<<1, K, 0, stream>>>(part, g, norm, scale, nb, K);
  const long long total = (long long)rows * K;
  wn_fold_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(v, scale, w_fwd, w_bwd, total, CG, K);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

// scratch: the one posconv_weight_norm_forward filled for these g, v (norm / scale are read back; the partial sums are reused)
int posconv_weight_norm_backward(const float* v, const float* dW, float* scratch, float* dg, float* dv, int H, int CG, int K,
                                 cudaStream_t stream) {
  SUTA_CHECK_ARG(v && dW && scratch && dg && dv && K > 0 && K <= 1024 && H % CG == 0);
  const int rows = H * CG;
  const int nb = rows < WN_BLOCKS ? rows : WN_BLOCKS;
  float *part = scratch, *norm = scratch + (long long)WN_BLOCKS * K, *scale = norm + K;
  wn_partial_kernel<1><<<nb, K, 0, stream>>>(v, dW, part, rows, CG, K);
  wn_dg_kernel<<<1, K, 0, stream>>>(part, norm, dg, nb, K);
  const long long total = (long long)rows * K;
  wn_dv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(v, dW, scale, norm, dg, dv, total, CG, K);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}
