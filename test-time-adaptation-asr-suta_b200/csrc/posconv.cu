// Layout kernels around the positional convolution (HF/modeling_wav2vec2.py:326-379, :690-691).
// The grouped Conv1d(H->H, k=K, pad=K/2, groups=G) itself runs on the tcgen05 GEMM (gemm_tc.cu): per group, the
// K-tap window of the channels-last group slab [R, H/G] is a contiguous run of K*(H/G) elements, so an
// overlapping-row TMA view makes it a plain K-major A operand.  These kernels
//   * scatter tokens into that slab layout with K/2 zero rows between utterances (the conv's zero padding, and
//     what keeps utterances independent inside one batch),
//   * apply GELU + residual on the way back (and the matching backward pieces).
// Element-wise, HBM-bound.
#include "kernels.cuh"

namespace {

__global__ void pack_kernel(const float* __restrict__ h, const float* __restrict__ conv, const int* __restrict__ row_utt,
                            const long long* __restrict__ tok_off, const long long* __restrict__ pad_off,
                            bf16* __restrict__ xg, long long M, int H, int CG, long long R, int row_shift) {
  const int h8 = H >> 3;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * h8) return;
  const long long row = idx / h8;
  const int c = (int)(idx - row * h8) * 8;
  const int u = row_utt[row];
  const long long prow = pad_off[u] + (row - tok_off[u]);
  float4 a = *reinterpret_cast<const float4*>(h + row * H + c);
  float4 b = *reinterpret_cast<const float4*>(h + row * H + c + 4);
  if (conv) {   // gradient variant: multiply by GELU'(pre-activation)
    const float* cp = conv + (prow + row_shift) * H + c;
    float4 p = *reinterpret_cast<const float4*>(cp), q = *reinterpret_cast<const float4*>(cp + 4);
    a.x *= gelu_erf_grad(p.x); a.y *= gelu_erf_grad(p.y); a.z *= gelu_erf_grad(p.z); a.w *= gelu_erf_grad(p.w);
    b.x *= gelu_erf_grad(q.x); b.y *= gelu_erf_grad(q.y); b.z *= gelu_erf_grad(q.z); b.w *= gelu_erf_grad(q.w);
  }
  const int g = c / CG, cg = c - g * CG;
  uint4 o = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
  *reinterpret_cast<uint4*>(xg + ((long long)g * R + prow) * CG + cg) = o;
}

// GRAD == false: out = h + GELU(conv[prow+shift]);  GRAD == true: out = h + conv[prow+shift]
template <bool GRAD>
__global__ void combine_kernel(const float* __restrict__ h, const float* __restrict__ conv,
                               const int* __restrict__ row_utt, const long long* __restrict__ tok_off,
                               const long long* __restrict__ pad_off, float* __restrict__ out, bf16* __restrict__ out16,
                               long long M, int H, int row_shift) {
  const int h4 = H >> 2;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * h4) return;
  const long long row = idx / h4;
  const int c = (int)(idx - row * h4) * 4;
  const int u = row_utt[row];
  const long long prow = pad_off[u] + (row - tok_off[u]) + row_shift;
  float4 a = *reinterpret_cast<const float4*>(h + row * H + c);
  float4 p = *reinterpret_cast<const float4*>(conv + prow * H + c);
  if (GRAD) { a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w; }
  else { a.x += gelu_erf(p.x); a.y += gelu_erf(p.y); a.z += gelu_erf(p.z); a.w += gelu_erf(p.w); }
  if (out) *reinterpret_cast<float4*>(out + row * H + c) = a;
  if (out16) *reinterpret_cast<uint2*>(out16 + row * H + c) = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
}

}  // namespace

int posconv_pack(const float* h, const int* row_utt, const long long* tok_off, const long long* pad_off, bf16* xg,
                 long long M, int H, int G, int CGP, long long R, cudaStream_t stream) {
  SUTA_CHECK_ARG(H == G * CGP && CGP % 8 == 0);
  long long n = M * (H >> 3);
  if (n <= 0) return SUTA_OK;
  pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(h, nullptr, row_utt, tok_off, pad_off, xg, M, H, CGP, R, 0);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

int posconv_pack_grad(const float* d_out, const float* conv, const int* row_utt, const long long* tok_off,
                      const long long* pad_off, bf16* dg, long long M, int H, int G, int CGP, long long R, int row_shift,
                      cudaStream_t stream) {
  SUTA_CHECK_ARG(H == G * CGP && CGP % 8 == 0 && conv);
  long long n = M * (H >> 3);
  if (n <= 0) return SUTA_OK;
  pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_out, conv, row_utt, tok_off, pad_off, dg, M, H, CGP, R, row_shift);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

int posconv_combine(const float* h, const float* conv, const int* row_utt, const long long* tok_off,
                    const long long* pad_off, float* h_out, long long M, int H, int row_shift, cudaStream_t stream) {
  long long n = M * (H >> 2);
  if (n <= 0) return SUTA_OK;
  combine_kernel<false><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(h, conv, row_utt, tok_off, pad_off, h_out, nullptr, M, H, row_shift);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

int posconv_combine_grad(const float* d_out, const float* dconv, const int* row_utt, const long long* tok_off,
                         const long long* pad_off, float* d_h, bf16* d_h_bf16, long long M, int H, int row_shift,
                         cudaStream_t stream) {
  long long n = M * (H >> 2);
  if (n <= 0) return SUTA_OK;
  combine_kernel<true><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_out, dconv, row_utt, tok_off, pad_off, d_h, d_h_bf16, M, H, row_shift);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}
