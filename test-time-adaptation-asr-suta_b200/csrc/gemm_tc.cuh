// tcgen05/TMEM GEMM:  D[M,N] = A[M,K] * B[N,K]^T  (bf16 operands, fp32 accumulate in tensor memory)
// with a fused epilogue (bias / GELU / GELU' / fp32 residual / fp32+bf16 stores).
// One kernel serves every dense contraction of the path (SURVEY.md 2.3 K2,K3,K4,K6,K8,K10):
//   * encoder linears forward  (A = activations, B = W)           HF/modeling_wav2vec2.py:500-573
//   * encoder linears backward (A = dY,          B = W^T)         autograd dgrad of the same
//   * strided conv layers as implicit GEMM: A is an overlapping-row TMA view of the channels-last input
//     (row t = x[s*t : s*t+k, :] flattened, which is contiguous)   HF/modeling_wav2vec2.py:269-272
//   * positional grouped conv: same trick per group, groups on the z axis  HF/modeling_wav2vec2.py:360-368
#pragma once
#include "common.cuh"

struct GemmOperand {
  const bf16* ptr;         // first element of row 0
  long long rows;          // number of addressable rows (TMA zero-fills beyond)
  long long row_stride;    // elements between consecutive rows (may be < row length: overlapping conv windows)
  // K-major (default): memory is [M or N rows][K contiguous].
  // MN-major (mn_major = 1): memory is [K rows][M or N contiguous] -- the natural layout of both operands of a
  // weight-gradient GEMM (reduction over time) and of W when computing dY * W; `cols` = extent of the M/N axis.
  int mn_major = 0;
  long long cols = 0;
};

struct GemmEpilogue {
  float* out_f32 = nullptr;        // optional fp32 output
  bf16* out_bf16 = nullptr;        // optional bf16 output (same leading dim)
  int out_ld = 0;
  const float* bias = nullptr;     // indexed [b_row_off + out_col], or [(b_row_off / N) * bias_utt_stride + out_col]
  long long bias_utt_stride = 0;   //   when bias_utt_stride != 0 (per-utterance bias inside the trainable vector)
  const float* residual = nullptr; // fp32 [rows, res_ld], added last (row-masked epilogue only)
  int res_ld = 0;
  int accumulate = 0;              // out_f32 += result instead of = (TMA reduce-add at the L2; dense outputs only):
                                   // how the encoder's residual stream is updated in place
  int act = 0;                     // 0 none, 1 GELU(erf), 2 multiply by aux_in
  const bf16* aux_in = nullptr;    // GELU'(pre-activation) saved by the forward (act == 2)
  bf16* aux_out = nullptr;         // where to save GELU'(pre-activation) (act == 1), may be null
  int aux_ld = 0;
};

struct GemmProblem {
  GemmOperand a, b;
  int M = 0, N = 0, K = 0;         // per-z problem; N % 16 == 0
  int nz = 1;                      // batch (z) count: A rows += z*a_z_rows, B rows += z*b_z_rows, out cols += z*c_z_cols
  long long a_z_rows = 0, b_z_rows = 0;
  int c_z_cols = 0;
  const int4* mblk = nullptr;      // optional per-M-block table {a_row0, out_row0, rows_valid, b_row_off}
  int num_mblk = 0;                // = ceil(M/128) when mblk == nullptr
  const int4* mpair = nullptr;     // optional per-256-row table {a_row0, out_row0, rows_valid, b_row_off} for the CTA-pair kernel:
  int num_mpair = 0;               //   both 128-row halves of an entry share b_row_off (256-row-aligned utterances)
  int tiles_own_rows = 0;          // with mblk: every tile may write all 128 rows from its out_row0 (the rows past
                                   // rows_valid are padding nobody else owns) -> eligible for the TMA-store epilogue
  long long out_rows = 0;          // addressable rows of the output buffers (default M)
  const int4* ztab = nullptr;      // optional per-z table {a_k_row0, b_k_row0, k_len, 0} (MN-major operands: the
                                   // reduction runs over rows [k_row0, k_row0 + k_len) of each operand)
  long long out_z_stride = 0;      // elements added to the output pointers per z (per-utterance gradient slabs)
  // MN-major B whose reduction index runs over TAPS of a conv weight [Cout][tap][Cin] (transposed-conv dgrad split by
  // output-row parity): k in [t * b_kwrap, (t+1) * b_kwrap) reads B rows k - t * b_kwrap at columns b_tap_col[t] + n
  int b_kwrap = 0;
  int b_tap_col[2] = {0, 0};
  double flops = 0.0;              // algorithmic FLOPs of this launch when it cannot be derived from M,N,K,nz
  GemmEpilogue epi;
};

int gemm_bf16_tc(const GemmProblem& p, cudaStream_t stream);
// gemm_tc2.cu: the cta_group::2 (CTA pair, 256 x 256 tiles) kernel for dense K-major problems with N % 256 == 0
int gemm_bf16_tc_2cta(const GemmProblem& p, cudaStream_t stream);
int gemm_encode_tmap(CUtensorMap* tm, int dtype, const void* ptr, long long cols, long long rows, long long pitch_bytes,
                     int box_cols, int box_rows, int swizzle);
int gemm_encode_tmap_nd(CUtensorMap* tm, const void* ptr, int rank, const long long* dims, const long long* strides_bytes,
                        const int* box);
int gemm_num_sms();
long long* gemm_trace_buffer(int* cap);   // debug timeline buffer set by suta_debug_set_gemm_trace (null = off)
