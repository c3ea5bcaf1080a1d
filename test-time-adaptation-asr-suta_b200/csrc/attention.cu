// Variable-length (per-utterance) multi-head self-attention, forward and backward, head_dim = 64.
// Restates HF/modeling_wav2vec2.py:500-544 + transformers/integrations/sdpa_attention.py:40-104 for the
// batched engine: non-causal, scale head_dim^-0.5, no mask inside an utterance, and -- because many
// utterances share one packed token axis -- keys/queries of other utterances are never visible.
//
// Round-1 implementation: flash-attention-2 style tiling on the legacy mma.sync bf16 tensor path
// (m16n8k16, fp32 accumulate, online softmax in registers).  The tcgen05/TMEM version is the next step
// (DESIGN.md "what comes next"); attention is 5-10 % of the path's FLOPs at LibriSpeech lengths.
//
// Layout: qkv bf16 [M, 3H] (q | k | v, head h at columns h*64), O bf16 [M, H], LSE fp32 [heads, M] in
// base-2 units, block table int4 {utt_row0, T_u, block_start_in_utt, 0} with one entry per 128 rows (these kernels
// work on 64-row blocks: two CTAs per entry).
#include <stdlib.h>

#include "kernels.cuh"

namespace {

constexpr int HD = 64;      // head dim
constexpr int BLK = 64;     // rows per tile (queries or keys)
constexpr int TILE_BYTES = BLK * HD * 2;

// swizzled [64][64] bf16 tile: 16-byte chunk c of row r lives at chunk (c ^ (r & 7))
__device__ __forceinline__ uint32_t tile_addr(uint32_t base, int r, int chunk) {
  return base + r * 128 + ((chunk ^ (r & 7)) << 4);
}

// cooperative async load of a [64 x 64] bf16 tile (rows row0.. of a row-major matrix with leading dim ld)
__device__ __forceinline__ void load_tile_async(uint32_t sbase, const bf16* g, long long row0, int rows_valid, int ld,
                                                int col0, int tid, int nthreads) {
  for (int i = tid; i < BLK * 8; i += nthreads) {
    int r = i >> 3, c = i & 7;
    bool ok = r < rows_valid;
    const bf16* src = g + (row0 + (ok ? r : 0)) * ld + col0 + c * 8;
    cp_async_16(tile_addr(sbase, r, c), src, ok);
  }
}

// A-operand fragments (16 rows x 64 k) of a warp's 16-row slab starting at row r0 of a swizzled tile
__device__ __forceinline__ void load_a_frags(uint32_t (&f)[4][4], uint32_t sbase, int r0, int lane) {
  int mid = lane >> 3, rr = lane & 7;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk)
    ldmatrix_x4(f[kk], tile_addr(sbase, r0 + rr + (mid & 1) * 8, kk * 2 + (mid >> 1)));
}

// C[16 x 64] += A(16 x 64 k, fragments) * T^T  where tile T is stored [n][k] (rows = n index)
__device__ __forceinline__ void mma_a_tileT(float (&c)[8][4], const uint32_t (&a)[4][4], uint32_t sbase, int lane) {
  int mid = lane >> 3, rr = lane & 7;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      uint32_t b[4];
      ldmatrix_x4(b, tile_addr(sbase, 8 * (2 * jp + (mid >> 1)) + rr, 2 * kk + (mid & 1)));
      mma_bf16_16816(c[2 * jp], a[kk], b[0], b[1]);
      mma_bf16_16816(c[2 * jp + 1], a[kk], b[2], b[3]);
    }
  }
}

// C[16 x 64 n] += P(16 x 64 k, given as C-layout probabilities) * T  where tile T is stored [k][n]
__device__ __forceinline__ void mma_p_tile(float (&c)[8][4], const float (&p)[8][4], uint32_t sbase, int lane) {
  int mid = lane >> 3, rr = lane & 7;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t a[4];
    a[0] = pack_bf16x2(p[2 * kk][0], p[2 * kk][1]);
    a[1] = pack_bf16x2(p[2 * kk][2], p[2 * kk][3]);
    a[2] = pack_bf16x2(p[2 * kk + 1][0], p[2 * kk + 1][1]);
    a[3] = pack_bf16x2(p[2 * kk + 1][2], p[2 * kk + 1][3]);
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      uint32_t b[4];
      ldmatrix_x4_trans(b, tile_addr(sbase, 16 * kk + (mid & 1) * 8 + rr, 2 * jp + (mid >> 1)));
      mma_bf16_16816(c[2 * jp], a, b[0], b[1]);
      mma_bf16_16816(c[2 * jp + 1], a, b[2], b[3]);
    }
  }
}

__device__ __forceinline__ void zero8x4(float (&c)[8][4]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) c[j][e] = 0.f;
}

// -------------------------------------------------------------------------------------------
// forward
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attn_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ O, float* __restrict__ LSE, const int4* __restrict__ tab,
                int H, long long M, float scale_log2) {
  __shared__ __align__(1024) uint8_t smem[5 * TILE_BYTES];
  const uint32_t sQ = smem_u32(smem), sK = sQ + TILE_BYTES, sV = sK + 2 * TILE_BYTES;
  const int4 t = tab[blockIdx.x >> 1];
  const int head = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long urow0 = t.x;
  const int T = t.y, m0 = t.z + 64 * (blockIdx.x & 1);
  if (m0 >= T) return;
  const int ld = 3 * H;
  const int nkv = (T + BLK - 1) / BLK;

  load_tile_async(sQ, qkv, urow0 + m0, T - m0, ld, head * HD, tid, 128);
  load_tile_async(sK, qkv, urow0, T, ld, H + head * HD, tid, 128);
  load_tile_async(sV, qkv, urow0, T, ld, 2 * H + head * HD, tid, 128);
  cp_async_commit();

  uint32_t qf[4][4];
  float o[8][4];
  zero8x4(o);
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  const int g = lane >> 2, c = lane & 3;

  for (int jb = 0; jb < nkv; ++jb) {
    const int buf = jb & 1;
    if (jb + 1 < nkv) {
      load_tile_async(sK + (buf ^ 1) * TILE_BYTES, qkv, urow0 + (jb + 1) * BLK, T - (jb + 1) * BLK, ld, H + head * HD, tid, 128);
      load_tile_async(sV + (buf ^ 1) * TILE_BYTES, qkv, urow0 + (jb + 1) * BLK, T - (jb + 1) * BLK, ld, 2 * H + head * HD, tid, 128);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (jb == 0) load_a_frags(qf, sQ, warp * 16, lane);

    float s[8][4];
    zero8x4(s);
    mma_a_tileT(s, qf, sK + buf * TILE_BYTES, lane);

    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int key = jb * BLK + 8 * j + 2 * c + (e & 1);
        float v = key < T ? s[j][e] * scale_log2 : -INFINITY;
        s[j][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    float alpha[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      float mnew = fmaxf(mrow[r], mx[r]);
      alpha[r] = exp2f(mrow[r] - mnew);
      mrow[r] = mnew;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float pv = exp2f(s[j][e] - mrow[e >> 1]);
        s[j][e] = pv;
        rs[e >> 1] += pv;
      }
#pragma unroll
    for (int r = 0; r < 2; ++r) lrow[r] = lrow[r] * alpha[r] + rs[r];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= alpha[0]; o[j][1] *= alpha[0];
      o[j][2] *= alpha[1]; o[j][3] *= alpha[1];
    }
    mma_p_tile(o, s, sV + buf * TILE_BYTES, lane);
    __syncthreads();
  }

#pragma unroll
  for (int r = 0; r < 2; ++r) {
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    int qi = m0 + warp * 16 + g + r * 8;
    if (qi < T) {
      float inv = 1.f / lrow[r];
      long long row = urow0 + qi;
      bf16* op = O + row * H + head * HD + 2 * c;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint32_t*>(op + 8 * j) = pack_bf16x2(o[j][2 * r] * inv, o[j][2 * r + 1] * inv);
      if (c == 0) LSE[(long long)head * M + row] = mrow[r] + log2f(lrow[r]);
    }
  }
}

// -------------------------------------------------------------------------------------------
// backward: D = rowsum(dO * O) per (row, head)
// -------------------------------------------------------------------------------------------
__global__ void attn_bwd_prep_kernel(const bf16* __restrict__ O, const bf16* __restrict__ dO, float* __restrict__ D,
                                     int H, int heads, long long M) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (row, head)
  if (i >= M * heads) return;
  long long row = i / heads;
  int h = (int)(i - row * heads);
  const uint4* o = reinterpret_cast<const uint4*>(O + row * H + h * HD);
  const uint4* d = reinterpret_cast<const uint4*>(dO + row * H + h * HD);
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint4 a = __ldg(o + k), b = __ldg(d + k);
    float2 x, y;
    x = unpack_bf16x2(a.x); y = unpack_bf16x2(b.x); acc += x.x * y.x + x.y * y.y;
    x = unpack_bf16x2(a.y); y = unpack_bf16x2(b.y); acc += x.x * y.x + x.y * y.y;
    x = unpack_bf16x2(a.z); y = unpack_bf16x2(b.z); acc += x.x * y.x + x.y * y.y;
    x = unpack_bf16x2(a.w); y = unpack_bf16x2(b.w); acc += x.x * y.x + x.y * y.y;
  }
  D[(long long)h * M + row] = acc;
}

// -------------------------------------------------------------------------------------------
// backward: dQ   (CTA = 64 queries of one (utterance, head); loop over key blocks)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dO, const float* __restrict__ LSE,
                   const float* __restrict__ D, bf16* __restrict__ dqkv, const int4* __restrict__ tab, int H, long long M,
                   float scale, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* smem = smem_dyn;
  const uint32_t sQ = smem_u32(smem), sdO = sQ + TILE_BYTES, sK = sdO + TILE_BYTES, sV = sK + 2 * TILE_BYTES;
  const int4 t = tab[blockIdx.x >> 1];
  const int head = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long urow0 = t.x;
  const int T = t.y, m0 = t.z + 64 * (blockIdx.x & 1);
  if (m0 >= T) return;
  const int ld = 3 * H;
  const int nkv = (T + BLK - 1) / BLK;
  const int g = lane >> 2, c = lane & 3;

  load_tile_async(sQ, qkv, urow0 + m0, T - m0, ld, head * HD, tid, 128);
  load_tile_async(sdO, dO, urow0 + m0, T - m0, H, head * HD, tid, 128);
  load_tile_async(sK, qkv, urow0, T, ld, H + head * HD, tid, 128);
  load_tile_async(sV, qkv, urow0, T, ld, 2 * H + head * HD, tid, 128);
  cp_async_commit();

  float lse[2], dd[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    int qi = m0 + warp * 16 + g + r * 8;
    bool ok = qi < T;
    lse[r] = ok ? LSE[(long long)head * M + urow0 + qi] : 0.f;
    dd[r] = ok ? D[(long long)head * M + urow0 + qi] : 0.f;
  }

  uint32_t qf[4][4], dof[4][4];
  float dq[8][4];
  zero8x4(dq);

  for (int jb = 0; jb < nkv; ++jb) {
    const int buf = jb & 1;
    if (jb + 1 < nkv) {
      load_tile_async(sK + (buf ^ 1) * TILE_BYTES, qkv, urow0 + (jb + 1) * BLK, T - (jb + 1) * BLK, ld, H + head * HD, tid, 128);
      load_tile_async(sV + (buf ^ 1) * TILE_BYTES, qkv, urow0 + (jb + 1) * BLK, T - (jb + 1) * BLK, ld, 2 * H + head * HD, tid, 128);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (jb == 0) {
      load_a_frags(qf, sQ, warp * 16, lane);
      load_a_frags(dof, sdO, warp * 16, lane);
    }
    float s[8][4], dp[8][4];
    zero8x4(s);
    zero8x4(dp);
    mma_a_tileT(s, qf, sK + buf * TILE_BYTES, lane);
    mma_a_tileT(dp, dof, sV + buf * TILE_BYTES, lane);
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int key = jb * BLK + 8 * j + 2 * c + (e & 1);
        float pv = key < T ? exp2f(s[j][e] * scale_log2 - lse[e >> 1]) : 0.f;
        s[j][e] = pv * (dp[j][e] - dd[e >> 1]);      // dS (natural units, unscaled)
      }
    mma_p_tile(dq, s, sK + buf * TILE_BYTES, lane);
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    int qi = m0 + warp * 16 + g + r * 8;
    if (qi < T) {
      bf16* op = dqkv + (urow0 + qi) * ld + head * HD + 2 * c;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint32_t*>(op + 8 * j) = pack_bf16x2(dq[j][2 * r] * scale, dq[j][2 * r + 1] * scale);
    }
  }
}

// -------------------------------------------------------------------------------------------
// backward: dK, dV   (CTA = 64 keys of one (utterance, head); loop over query blocks)
// Works on the transposed problem so each warp owns 16 keys and needs no cross-warp reduction:
//   S^T = K Q^T,  dP^T = V dO^T,  dV += P^T dO,  dK += dS^T Q.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attn_bwd_dkv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dO, const float* __restrict__ LSE,
                    const float* __restrict__ D, bf16* __restrict__ dqkv, const int4* __restrict__ tab, int H, long long M,
                    float scale, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* smem = smem_dyn;
  const uint32_t sK = smem_u32(smem), sV = sK + TILE_BYTES, sQ = sV + TILE_BYTES, sdO = sQ + 2 * TILE_BYTES;
  float* sLSE = reinterpret_cast<float*>(smem + 6 * TILE_BYTES);     // [2][64]
  float* sD = sLSE + 2 * BLK;                                        // [2][64]
  const int4 t = tab[blockIdx.x >> 1];
  const int head = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long urow0 = t.x;
  const int T = t.y, n0 = t.z + 64 * (blockIdx.x & 1);            // n0 = first key of this block
  if (n0 >= T) return;
  const int ld = 3 * H;
  const int nq = (T + BLK - 1) / BLK;
  const int g = lane >> 2, c = lane & 3;

  auto load_q_block = [&](int ib, int buf) {
    load_tile_async(sQ + buf * TILE_BYTES, qkv, urow0 + ib * BLK, T - ib * BLK, ld, head * HD, tid, 128);
    load_tile_async(sdO + buf * TILE_BYTES, dO, urow0 + ib * BLK, T - ib * BLK, H, head * HD, tid, 128);
    if (tid < BLK) {
      int qi = ib * BLK + tid;
      bool ok = qi < T;
      sLSE[buf * BLK + tid] = ok ? LSE[(long long)head * M + urow0 + qi] : 0.f;
      sD[buf * BLK + tid] = ok ? D[(long long)head * M + urow0 + qi] : 0.f;
    }
  };

  load_tile_async(sK, qkv, urow0 + n0, T - n0, ld, H + head * HD, tid, 128);
  load_tile_async(sV, qkv, urow0 + n0, T - n0, ld, 2 * H + head * HD, tid, 128);
  load_q_block(0, 0);
  cp_async_commit();

  uint32_t kf[4][4], vf[4][4];
  float dk[8][4], dv[8][4];
  zero8x4(dk);
  zero8x4(dv);
  const int key_lo = n0 + warp * 16 + g;      // this thread's keys: key_lo (e<2) and key_lo+8 (e>=2)

  for (int ib = 0; ib < nq; ++ib) {
    const int buf = ib & 1;
    if (ib + 1 < nq) {
      load_q_block(ib + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (ib == 0) {
      load_a_frags(kf, sK, warp * 16, lane);
      load_a_frags(vf, sV, warp * 16, lane);
    }
    float st[8][4], dpt[8][4];
    zero8x4(st);
    zero8x4(dpt);
    mma_a_tileT(st, kf, sQ + buf * TILE_BYTES, lane);      // S^T  [16 keys x 64 queries]
    mma_a_tileT(dpt, vf, sdO + buf * TILE_BYTES, lane);    // dP^T
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int ql = 8 * j + 2 * c + (e & 1);
        int qi = ib * BLK + ql;
        int key = key_lo + (e >> 1) * 8;
        float pv = (qi < T && key < T) ? exp2f(st[j][e] * scale_log2 - sLSE[buf * BLK + ql]) : 0.f;
        st[j][e] = pv;                                         // P^T
        dpt[j][e] = pv * (dpt[j][e] - sD[buf * BLK + ql]);     // dS^T
      }
    mma_p_tile(dv, st, sdO + buf * TILE_BYTES, lane);
    mma_p_tile(dk, dpt, sQ + buf * TILE_BYTES, lane);
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    int key = key_lo + r * 8;
    if (key < T) {
      bf16* kp = dqkv + (urow0 + key) * ld + H + head * HD + 2 * c;
      bf16* vp = dqkv + (urow0 + key) * ld + 2 * H + head * HD + 2 * c;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        *reinterpret_cast<uint32_t*>(kp + 8 * j) = pack_bf16x2(dk[j][2 * r] * scale, dk[j][2 * r + 1] * scale);
        *reinterpret_cast<uint32_t*>(vp + 8 * j) = pack_bf16x2(dv[j][2 * r], dv[j][2 * r + 1]);
      }
    }
  }
}

constexpr int DQ_SMEM = 6 * TILE_BYTES;
constexpr int DKV_SMEM = 6 * TILE_BYTES + 4 * BLK * 4;

}  // namespace

int attention_forward(const bf16* qkv, bf16* O, float* LSE, const int4* blk_tab, int n_blk, int H, int heads, long long M,
                      cudaStream_t stream) {
  SUTA_CHECK_ARG(H == heads * HD);
  if (n_blk <= 0) return SUTA_OK;
  static const bool legacy = getenv("SUTA_ATTN_LEGACY") != nullptr;    // A/B switch while the tcgen05 kernels are tuned
  if (!legacy) return attention_forward_tc(qkv, O, LSE, blk_tab, n_blk, H, heads, M, stream);
  const float scale = 0.125f;   // 64^-0.5
  attn_fwd_kernel<<<dim3(2 * n_blk, heads), 128, 0, stream>>>(qkv, O, LSE, blk_tab, H, M, scale * 1.4426950408889634f);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

int attention_backward(const bf16* qkv, const bf16* O, const bf16* dO, const float* LSE, float* D, bf16* dqkv,
                       const int4* blk_tab, int n_blk, int H, int heads, long long M, cudaStream_t stream) {
  SUTA_CHECK_ARG(H == heads * HD);
  if (n_blk <= 0) return SUTA_OK;
  static bool attr = false;
  if (!attr) {
    CUDA_TRY(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV_SMEM));
    attr = true;
  }
  const float scale = 0.125f, scale_log2 = scale * 1.4426950408889634f;
  long long n = M * heads;
  attn_bwd_prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(O, dO, D, H, heads, M);
  CUDA_TRY(cudaGetLastError());
  attn_bwd_dq_kernel<<<dim3(2 * n_blk, heads), 128, DQ_SMEM, stream>>>(qkv, dO, LSE, D, dqkv, blk_tab, H, M, scale, scale_log2);
  CUDA_TRY(cudaGetLastError());
  attn_bwd_dkv_kernel<<<dim3(2 * n_blk, heads), 128, DKV_SMEM, stream>>>(qkv, dO, LSE, D, dqkv, blk_tab, H, M, scale, scale_log2);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}
