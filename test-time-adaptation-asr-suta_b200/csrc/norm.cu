// Fused LayerNorm forward / backward with PER-UTTERANCE affine parameters.
// Restates torch native_layer_norm(+backward) as used by HF/modeling_wav2vec2.py:429-434,692,599-602 for a
// token-packed batch in which every utterance carries its own adapted gamma/beta (REF/main.py:81-87).
// Forward: one warp per row, row statistics by warp shuffle, rows streamed through a bulk-async shared-memory ring.
// Backward: see ln_bwd_kernel.  HBM-bound.
#include "kernels.cuh"
#include <algorithm>
#include <atomic>

namespace {

int n_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

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

constexpr int FWD_W = 8;          // warps per CTA = rows per tile
// tiles in flight per CTA: 3 for the wide fp32 rows of the encoder (24-32 KB tiles); narrow bf16 rows (the conv-layer and
// feature-projection LayerNorms: 8 KB tiles at N = 512) need more of them to keep enough bytes in flight per SM -- with 3
// stages those kernels sat at ~0.48 of the HBM rate (48 KB in flight per SM), the wide ones at 0.8
template <int N, typename TIn>
struct FwdSmem {
  static constexpr int TILE_BYTES = FWD_W * N * (int)sizeof(TIn);
  static constexpr int STAGES = TILE_BYTES <= 8192 ? 6 : (TILE_BYTES <= 16384 ? 4 : 3);
  static constexpr int BYTES = STAGES * TILE_BYTES + 64;
};

__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// One warp per row, 8 rows per tile; the tiles of a CTA's row range stream through a 3-deep shared-memory ring filled by
// 1-D bulk copies issued two tiles ahead (the register-only version was latency-bound at 41 % of the HBM rate: its loads
// could not start before the previous row's dependent chain -- statistics, per-utterance gamma/beta, stores -- had drained).
// MODE 0: y = LN(x) (the post-LN encoder, feature projection).  MODE 1: the bf16 output is GELU(LN(x)) -- the conv layers of
// the LayerNorm feature extractor (HF/modeling_wav2vec2.py:291-299).  MODE 2: the fp32 output is x (+ y32_bias), NOT
// LN(x): the pre-LN ("stable") encoder keeps the residual stream beside the normalised branch (HF:638-645), so the
// LayerNorm seeds the next residual buffer with its own INPUT.  MODE 1 only: rows with row_utt < 0 (gaps between utterances
// in the conv layouts) are skipped.
template <int N, typename TIn, int MODE>
__global__ void __launch_bounds__(FWD_W * 32)
ln_fwd_kernel(const TIn* __restrict__ x, const int* __restrict__ row_utt, const float* __restrict__ P, long long pstride,
              int g_off, int b_off, float* __restrict__ y32, bf16* __restrict__ y16, float* __restrict__ mean_out,
              float* __restrict__ rstd_out, long long M, float eps, const float* __restrict__ y32_bias, int rows_per_cta) {
  constexpr int NV = Cols<N>::NV;
  using S = FwdSmem<N, TIn>;
  constexpr int FWD_STAGES = S::STAGES;
  extern __shared__ __align__(128) uint8_t ring[];
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + FWD_STAGES * S::TILE_BYTES);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long cta_row0 = (long long)blockIdx.x * rows_per_cta;
  const int n_tiles = (int)min((long long)(rows_per_cta / FWD_W), (M - cta_row0 + FWD_W - 1) / FWD_W);
  auto issue = [&](int t) {                        // thread 0: fetch tile t into stage t % FWD_STAGES
    const long long r0 = cta_row0 + (long long)t * FWD_W;
    const int nr = (int)min((long long)FWD_W, M - r0);
    uint64_t* bar = &full[t % FWD_STAGES];
    mbar_expect_tx(bar, (uint32_t)(nr * N * sizeof(TIn)));
    bulk_load_1d(smem_u32(ring + (t % FWD_STAGES) * S::TILE_BYTES), x + r0 * N, (uint32_t)(nr * N * sizeof(TIn)), bar);
  };
  if (threadIdx.x == 0) {
    for (int s2 = 0; s2 < FWD_STAGES; ++s2) mbar_init(&full[s2], 1);
    mbar_fence_init();
    for (int t = 0; t < FWD_STAGES - 1 && t < n_tiles; ++t) issue(t);
  }
  __syncthreads();
  int u_next = cta_row0 + warp < M ? __ldg(row_utt + cta_row0 + warp) : 0;
  int u_cur = -1;
  float4 gv[NV], bv[NV], nbv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) gv[i] = bv[i] = nbv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
  for (int it = 0; it < n_tiles; ++it) {
    // every warp passed the __syncthreads of tile it - 1, i.e. finished reading the stage tile it + 2 goes into
    if (threadIdx.x == 0 && it + FWD_STAGES - 1 < n_tiles) issue(it + FWD_STAGES - 1);
    const long long row = cta_row0 + (long long)it * FWD_W + warp;
    const int u = u_next;
    u_next = row + FWD_W < M ? __ldg(row_utt + row + FWD_W) : 0;
    if (u != u_cur && (MODE != 1 || u >= 0) && row < M) {         // per-utterance gamma/beta stay in registers until the utterance changes
      u_cur = u;
      const float* gam = P + (long long)u * pstride + g_off;
      const float* bet = P + (long long)u * pstride + b_off;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (Cols<N>::active(lane, i)) {
          gv[i] = __ldg(reinterpret_cast<const float4*>(gam + 4 * lane + 128 * i));
          bv[i] = __ldg(reinterpret_cast<const float4*>(bet + 4 * lane + 128 * i));
          if (y32_bias) nbv[i] = __ldg(reinterpret_cast<const float4*>(y32_bias + 4 * lane + 128 * i));
        }
    }
    mbar_wait(&full[it % FWD_STAGES], (uint32_t)((it / FWD_STAGES) & 1));
    if (row < M && (MODE != 1 || u >= 0)) {
      const TIn* xr = reinterpret_cast<const TIn*>(ring + (it % FWD_STAGES) * S::TILE_BYTES) + warp * N;
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
      if (MODE == 1) {
        uint64_t acc = 0ull;
        const uint64_t nmean = dup2(-mean);
#pragma unroll
        for (int i = 0; i < NV; ++i)
          if (Cols<N>::active(lane, i)) {
            const uint64_t a = add2(pk2(v[i].x, v[i].y), nmean), c = add2(pk2(v[i].z, v[i].w), nmean);
            acc = fma2(a, a, fma2(c, c, acc));
          }
        float s0, s1;
        upk2(acc, s0, s1);
        ss = s0 + s1;
      }
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (MODE != 1 && Cols<N>::active(lane, i)) {
          float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
          ss += a * a + b * b + c * c + d * d;
        }
      const float rstd = rsqrtf(warp_sum(ss) * (1.0f / N) + eps);
      if (lane == 0) {
        mean_out[row] = mean;
        rstd_out[row] = rstd;
      }
      if (MODE == 1) {
        // GELU(LayerNorm) is bound by instruction issue (ncu: 72 % issue, 26 instructions per element): the affine and
        // the erf polynomial run on packed fp32 pairs (FFMA2), half the FMA-pipe instructions
        const uint64_t rs2 = dup2(rstd), nm2 = dup2(-mean * rstd);
#pragma unroll
        for (int i = 0; i < NV; ++i)
          if (Cols<N>::active(lane, i)) {
            const int col = 4 * lane + 128 * i;
            const float4 g = gv[i], b = bv[i];
            const uint64_t y01 = gelu_erf2(fma2(fma2(pk2(v[i].x, v[i].y), rs2, nm2), pk2(g.x, g.y), pk2(b.x, b.y)));
            const uint64_t y23 = gelu_erf2(fma2(fma2(pk2(v[i].z, v[i].w), rs2, nm2), pk2(g.z, g.w), pk2(b.z, b.w)));
            float4 o;
            upk2(y01, o.x, o.y);
            upk2(y23, o.z, o.w);
            if (y16) *reinterpret_cast<uint2*>(y16 + row * N + col) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
            if (y32) *reinterpret_cast<float4*>(y32 + row * N + col) = o;
          }
      }
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (MODE != 1 && Cols<N>::active(lane, i)) {
          const int col = 4 * lane + 128 * i;
          const float4 g = gv[i], b = bv[i];
          float4 o;
          o.x = (v[i].x - mean) * rstd * g.x + b.x;
          o.y = (v[i].y - mean) * rstd * g.y + b.y;
          o.z = (v[i].z - mean) * rstd * g.z + b.z;
          o.w = (v[i].w - mean) * rstd * g.w + b.w;
          if (y16) *reinterpret_cast<uint2*>(y16 + row * N + col) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
          if (MODE == 2) o = v[i];
          if (y32) {
            if (y32_bias) {   // the fp32 copy seeds the next residual sum: the bias of the GEMM that accumulates into it rides along
              const float4 nb = nbv[i];
              o.x += nb.x; o.y += nb.y; o.z += nb.z; o.w += nb.w;
            }
            *reinterpret_cast<float4*>(y32 + row * N + col) = o;
          }
        }
    }
    __syncthreads();
  }
}

// Backward.  One thread owns 4 fixed columns for all rows of its CTA (N/4 threads per CTA), so dgamma/dbeta live in 8
// registers and are flushed once per (CTA, utterance) into a scratch slot of their own (slot = CTA index + utterance
// index: both are monotone along the packed rows, so the sum is unique); ln_bwd_reduce_kernel then adds the slots of
// every utterance in a fixed order -- the result is bit-reproducible run to run (the first version used fp32
// atomicAdd into G, whose order is not).  The two row reductions are batched
// over R = 4 rows: every thread first issues the 8 independent 16-byte loads of the tile, then 8 warp reductions, one
// exchange through shared memory, one __syncthreads.  (The first version kept whole rows per warp: 150 registers, one
// CTA per SM, 3 TB/s.)
constexpr int BWD_R = 4;          // rows per tile
constexpr int BWD_STAGES = 3;     // tiles in flight per CTA (bulk-async copies into shared memory)

template <int N, typename TIn, typename TDy = float>
struct BwdSmem {
  static constexpr int X_BYTES = BWD_R * N * (int)sizeof(TIn);
  static constexpr int D_BYTES = BWD_R * N * (int)sizeof(TDy);
  static constexpr int STAGE_BYTES = X_BYTES + D_BYTES;
  static constexpr int BYTES = BWD_STAGES * STAGE_BYTES + 64;
};

// The tiles (4 rows of x and of dy: 24 KB at N = 768) stream through a 3-deep shared-memory ring filled by 1-D bulk
// copies that one thread issues two tiles ahead -- with register-only loads the kernel had ~45 KB in flight per SM and
// reached 34 % of the HBM rate (profiles/r01e); the __syncthreads of the row reduction doubles as the "stage is free" signal.
// MODE 0: plain LayerNorm backward.  MODE 1: the forward applied GELU after the LayerNorm (conv layers of the LayerNorm
// feature extractor): dy is the gradient of the GELU OUTPUT; y = xhat gamma + beta is recomputed and dy * GELU'(y) takes
// dy's place (needs beta: b_off).  MODE 2: dx += dx_add (fp32, may alias dx32): the pre-LN encoder's residual path,
// d(residual) = d(residual after the branch) + LayerNorm-backward(d branch input).  MODE 1 only: rows with row_utt < 0 are skipped.
// dx32 / dx16 may alias dy (rows are staged in shared memory before their outputs are written).
template <int N, typename TIn, typename TDy, int MODE>
__global__ void __launch_bounds__((N / 4 + 31) / 32 * 32)
ln_bwd_kernel(const TDy* __restrict__ dy, const TIn* __restrict__ x, const float* __restrict__ mean,
              const float* __restrict__ rstd, const int* __restrict__ row_utt, const float* __restrict__ P,
              long long pstride, int g_off, int b_off, float* __restrict__ part, float* dx32,
              bf16* dx16, const float* dx_add, long long M, int rows_per_cta) {
  constexpr int NT = N / 4;                       // active threads
  constexpr int NW = (NT + 31) / 32;              // warps
  constexpr bool GAPS = MODE == 1;                // only the conv-layer row spaces have gap rows (row_utt < 0)
  using S = BwdSmem<N, TIn, TDy>;
  extern __shared__ __align__(128) uint8_t ring[];
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + BWD_STAGES * S::STAGE_BYTES);
  __shared__ float red[2][2][NW][BWD_R];          // [tile parity][c1 | c2][warp][row]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool act = tid < NT;
  const int col = 4 * tid;
  float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag, g = ag, bb = ag;
  int u_acc = -1;                                 // utterance the accumulators (and g) belong to

  auto flush = [&]() {
    if (u_acc < 0 || !part || !act) return;
    float* slot = part + ((long long)blockIdx.x + u_acc) * (2 * N);
    *reinterpret_cast<float4*>(slot + col) = ag;
    *reinterpret_cast<float4*>(slot + N + col) = ab;
    ag = ab = make_float4(0.f, 0.f, 0.f, 0.f);
  };

  const long long cta_row0 = (long long)blockIdx.x * rows_per_cta;
  const int n_tiles = (int)min((long long)(rows_per_cta / BWD_R), (M - cta_row0 + BWD_R - 1) / BWD_R);
  auto issue = [&](int t) {                        // thread 0: fetch tile t into stage t % BWD_STAGES
    const long long r0 = cta_row0 + (long long)t * BWD_R;
    const int nr = (int)min((long long)BWD_R, M - r0);
    uint64_t* bar = &full[t % BWD_STAGES];
    const uint32_t dst = smem_u32(ring + (t % BWD_STAGES) * S::STAGE_BYTES);
    mbar_expect_tx(bar, (uint32_t)(nr * N * (sizeof(TIn) + sizeof(TDy))));
    bulk_load_1d(dst, x + r0 * N, (uint32_t)(nr * N * sizeof(TIn)), bar);
    bulk_load_1d(dst + S::X_BYTES, dy + r0 * N, (uint32_t)(nr * N * sizeof(TDy)), bar);
  };
  if (tid == 0) {
    for (int s2 = 0; s2 < BWD_STAGES; ++s2) mbar_init(&full[s2], 1);
    mbar_fence_init();
    for (int t = 0; t < BWD_STAGES - 1 && t < n_tiles; ++t) issue(t);
  }
  __syncthreads();
  float mu_n[BWD_R], rs_n[BWD_R];
  int uu_n[BWD_R];
  auto load_scalars = [&](long long r0) {
#pragma unroll
    for (int r = 0; r < BWD_R; ++r) {
      const bool ok = r0 + r < M;
      uu_n[r] = ok ? __ldg(row_utt + r0 + r) : -1;
      mu_n[r] = ok && (!GAPS || uu_n[r] >= 0) ? __ldg(mean + r0 + r) : 0.f;     // gap rows: no statistics were written
      rs_n[r] = ok && (!GAPS || uu_n[r] >= 0) ? __ldg(rstd + r0 + r) : 0.f;
    }
  };
  load_scalars(cta_row0);
#pragma unroll 1
  for (int it = 0; it < n_tiles; ++it) {
    const long long row0 = cta_row0 + it * BWD_R;
    // every thread passed the __syncthreads of tile it - 1, i.e. finished reading the stage tile it + 2 goes into
    if (tid == 0 && it + BWD_STAGES - 1 < n_tiles) issue(it + BWD_STAGES - 1);
    float4 xv[BWD_R], dv[BWD_R];
    float mu[BWD_R], rs[BWD_R];
    int uu[BWD_R];
#pragma unroll
    for (int r = 0; r < BWD_R; ++r) { uu[r] = uu_n[r]; mu[r] = mu_n[r]; rs[r] = rs_n[r]; }
    load_scalars(row0 + BWD_R);                    // next tile's row constants: their L2 latency hides behind this tile
    float4 av[MODE == 2 ? BWD_R : 1];
    if (MODE == 2) {                               // residual-path gradient: fetched now, consumed after the reductions
#pragma unroll
      for (int r = 0; r < BWD_R; ++r) {
        av[MODE == 2 ? r : 0] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + r < M && act) av[MODE == 2 ? r : 0] = *reinterpret_cast<const float4*>(dx_add + (row0 + r) * N + col);
      }
    }
    mbar_wait(&full[it % BWD_STAGES], (uint32_t)((it / BWD_STAGES) & 1));
    const uint8_t* st = ring + (it % BWD_STAGES) * S::STAGE_BYTES;
#pragma unroll
    for (int r = 0; r < BWD_R; ++r) {
      xv[r] = dv[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < M && act && (!GAPS || uu[r] >= 0)) {
        xv[r] = Loader<TIn>::ld4(reinterpret_cast<const TIn*>(st) + r * N + col);
        dv[r] = Loader<TDy>::ld4(reinterpret_cast<const TDy*>(st + S::X_BYTES) + r * N + col);
      }
    }
    float p1[BWD_R], p2[BWD_R];
#pragma unroll
    for (int r = 0; r < BWD_R; ++r) {
      if (uu[r] != u_acc && uu[r] >= 0) {         // utterance boundary (uniform over the CTA): new gamma, new accumulators
        flush();
        u_acc = uu[r];
        if (act) g = __ldg(reinterpret_cast<const float4*>(P + (long long)u_acc * pstride + g_off + col));
        if (act && MODE == 1) bb = __ldg(reinterpret_cast<const float4*>(P + (long long)u_acc * pstride + b_off + col));
      }
      if (MODE == 1) {
        // through the GELU that followed the LayerNorm; packed fp32 pairs (FFMA2): this kernel is bound by instruction
        // issue, not by memory (profiles/r02g_ncu_full_lv60_kernels.md)
        const uint64_t rs2 = dup2(rs[r]), nm2 = dup2(-mu[r] * rs[r]);
        const uint64_t g01 = pk2(g.x, g.y), g23 = pk2(g.z, g.w);
        const uint64_t xh01 = fma2(pk2(xv[r].x, xv[r].y), rs2, nm2), xh23 = fma2(pk2(xv[r].z, xv[r].w), rs2, nm2);
        const uint64_t d01 = mul2(pk2(dv[r].x, dv[r].y), gelu_erf_grad2(fma2(xh01, g01, pk2(bb.x, bb.y))));
        const uint64_t d23 = mul2(pk2(dv[r].z, dv[r].w), gelu_erf_grad2(fma2(xh23, g23, pk2(bb.z, bb.w))));
        upk2(fma2(d01, xh01, pk2(ag.x, ag.y)), ag.x, ag.y);
        upk2(fma2(d23, xh23, pk2(ag.z, ag.w)), ag.z, ag.w);
        upk2(add2(pk2(ab.x, ab.y), d01), ab.x, ab.y);
        upk2(add2(pk2(ab.z, ab.w), d23), ab.z, ab.w);
        const uint64_t dxh01 = mul2(d01, g01), dxh23 = mul2(d23, g23);
        float a0, a1, b0, b1;
        upk2(add2(dxh01, dxh23), a0, a1);
        upk2(fma2(dxh01, xh01, mul2(dxh23, xh23)), b0, b1);
        p1[r] = a0 + a1;
        p2[r] = b0 + b1;
        upk2(xh01, xv[r].x, xv[r].y); upk2(xh23, xv[r].z, xv[r].w);
        upk2(dxh01, dv[r].x, dv[r].y); upk2(dxh23, dv[r].z, dv[r].w);
        continue;
      }
      float4 d = dv[r];
      float4 xh = make_float4((xv[r].x - mu[r]) * rs[r], (xv[r].y - mu[r]) * rs[r], (xv[r].z - mu[r]) * rs[r], (xv[r].w - mu[r]) * rs[r]);
      ag.x = fmaf(d.x, xh.x, ag.x); ag.y = fmaf(d.y, xh.y, ag.y); ag.z = fmaf(d.z, xh.z, ag.z); ag.w = fmaf(d.w, xh.w, ag.w);
      ab.x += d.x; ab.y += d.y; ab.z += d.z; ab.w += d.w;
      const float4 dxh = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
      p1[r] = (dxh.x + dxh.y) + (dxh.z + dxh.w);
      p2[r] = (dxh.x * xh.x + dxh.y * xh.y) + (dxh.z * xh.z + dxh.w * xh.w);
      xv[r] = xh;
      dv[r] = dxh;
    }
    if (dx32 || dx16) {
#pragma unroll
      for (int r = 0; r < BWD_R; ++r) {
        p1[r] = warp_sum(p1[r]);
        p2[r] = warp_sum(p2[r]);
      }
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < BWD_R; ++r) {
          red[it & 1][0][warp][r] = p1[r];
          red[it & 1][1][warp][r] = p2[r];
        }
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < BWD_R; ++r) {
        const long long row = row0 + r;
        if (row >= M || !act || (GAPS && uu[r] < 0)) continue;
        float c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          c1 += red[it & 1][0][w][r];
          c2 += red[it & 1][1][w][r];
        }
        c1 *= 1.0f / N;
        c2 *= 1.0f / N;
        float4 o;
        if (MODE == 1) {
          const uint64_t rs2 = dup2(rs[r]), nc1 = dup2(-c1), nc2 = dup2(-c2);
          upk2(mul2(rs2, fma2(pk2(xv[r].x, xv[r].y), nc2, add2(pk2(dv[r].x, dv[r].y), nc1))), o.x, o.y);
          upk2(mul2(rs2, fma2(pk2(xv[r].z, xv[r].w), nc2, add2(pk2(dv[r].z, dv[r].w), nc1))), o.z, o.w);
        } else {
          o.x = rs[r] * (dv[r].x - c1 - xv[r].x * c2);
          o.y = rs[r] * (dv[r].y - c1 - xv[r].y * c2);
          o.z = rs[r] * (dv[r].z - c1 - xv[r].z * c2);
          o.w = rs[r] * (dv[r].w - c1 - xv[r].w * c2);
        }
        if (MODE == 2) { const float4 a = av[MODE == 2 ? r : 0]; o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
        if (dx32) *reinterpret_cast<float4*>(dx32 + row * N + col) = o;
        if (dx16) *reinterpret_cast<uint2*>(dx16 + row * N + col) = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
      }
    } else {
      __syncthreads();                            // the ring hand-over needs one CTA-wide sync per tile either way
    }
  }
  flush();
}

// G[u][g_off + c] = sum over the CTAs that saw rows of utterance u (ascending) of their dgamma slot; same for dbeta.
__global__ void __launch_bounds__(256)
ln_bwd_reduce_kernel(const float* __restrict__ part, const long long* __restrict__ tok_off, const int* __restrict__ T,
                     int rows_per_cta, int N, float* __restrict__ G, long long pstride, int g_off, int b_off) {
  const int u = blockIdx.x;
  const long long r0 = tok_off[u], r1 = r0 + T[u] - 1;
  const int b0 = (int)(r0 / rows_per_cta), b1 = (int)(r1 / rows_per_cta);
  for (int c = threadIdx.x; c < 2 * N; c += blockDim.x) {
    float s = 0.f;
    for (int b = b0; b <= b1; ++b) s += part[((long long)b + u) * (2 * N) + c];
    G[(long long)u * pstride + (c < N ? g_off + c : b_off + c - N)] = s;
  }
}

// The same reduction for MANY LayerNorms in one launch (the engine defers the reductions of a whole backward pass:
// 2 x layers + 2 small launches become one): blockIdx.y selects the LayerNorm.
__global__ void __launch_bounds__(256)
ln_bwd_reduce_many_kernel(const __grid_constant__ LnReduceBatch b, const long long* __restrict__ tok_off,
                          const int* __restrict__ T, float* __restrict__ G, long long pstride) {
  const LnReduceItem it = b.item[blockIdx.y];
  const int u = blockIdx.x;
  if (it.tok_off) { tok_off = it.tok_off; T = it.T; }      // a LayerNorm over rows of its own (conv layer layouts)
  const long long r0 = tok_off[u], r1 = r0 + T[u] - 1;
  const int b0 = (int)(r0 / it.rows_per_cta), b1 = (int)(r1 / it.rows_per_cta);
  for (int c = threadIdx.x; c < 2 * it.N; c += blockDim.x) {
    float s = 0.f;
    for (int k = b0; k <= b1; ++k) s += it.part[((long long)k + u) * (2 * it.N) + c];
    G[(long long)u * pstride + (c < it.N ? it.g_off + c : it.b_off + c - it.N)] = s;
  }
}

template <int N, typename TIn, int MODE>
int launch_fwd(const TIn* x, const int* row_utt, UttParams prm, int g_off, int b_off, float* y32, bf16* y16, float* mean,
               float* rstd, long long M, float eps, const float* y32_bias, cudaStream_t stream) {
  using S = FwdSmem<N, TIn>;
  // CTAs of this instantiation that fit on one SM.  Published only once it is final (attribute set, value clamped): a second
  // host thread driving another engine must never see the raw occupancy (> 8 would overrun the backward's slot scratch)
  static std::atomic<int> resident_pub{0};
  int resident = resident_pub.load(std::memory_order_acquire);
  if (!resident) {
    CUDA_TRY(cudaFuncSetAttribute(ln_fwd_kernel<N, TIn, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::BYTES));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, ln_fwd_kernel<N, TIn, MODE>, FWD_W * 32, S::BYTES));
    resident = std::max(1, std::min(8, resident));
    resident_pub.store(resident, std::memory_order_release);
  }
  // one wave: at most (resident CTAs per SM) x (SMs) CTAs, rows per CTA a multiple of the tile height
  long long rows = (M + (long long)resident * n_sms() - 1) / ((long long)resident * n_sms());
  rows = std::max<long long>(FWD_W, (rows + FWD_W - 1) / FWD_W * FWD_W);
  ln_fwd_kernel<N, TIn, MODE><<<(unsigned)((M + rows - 1) / rows), FWD_W * 32, S::BYTES, stream>>>(
      x, row_utt, prm.P, prm.stride, g_off, b_off, y32, y16, mean, rstd, M, eps, y32_bias, (int)rows);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}
// rows per CTA of the backward grid: one wave of at most (resident CTAs per SM) x (SMs) CTAs; 8 is the cap on `resident`,
// so 8 x SMs + n_utts slots always suffice for the dgamma/dbeta scratch
long long bwd_rows_per_cta(long long M, int resident) {
  long long rows = (M + (long long)resident * n_sms() - 1) / ((long long)resident * n_sms());
  return std::max<long long>(16, (rows + BWD_R - 1) / BWD_R * BWD_R);
}

template <int N, typename TIn, typename TDy, int MODE>
int launch_bwd(const TDy* dy, const TIn* x, const float* mean, const float* rstd, const int* row_utt, UttParams prm,
               int g_off, int b_off, float* G, float* dx32, bf16* dx16, const float* dx_add, long long M,
               const long long* tok_off, const int* T, int n_utts, float* scratch, cudaStream_t stream, LnReduceItem* defer) {
  constexpr int threads = (N / 4 + 31) / 32 * 32;
  using S = BwdSmem<N, TIn, TDy>;
  static std::atomic<int> resident_pub{0};         // see launch_fwd: published only once final
  int resident = resident_pub.load(std::memory_order_acquire);
  if (!resident) {
    CUDA_TRY(cudaFuncSetAttribute(ln_bwd_kernel<N, TIn, TDy, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::BYTES));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, ln_bwd_kernel<N, TIn, TDy, MODE>, threads, S::BYTES));
    resident = std::max(1, std::min(8, resident));
    resident_pub.store(resident, std::memory_order_release);
  }
  const long long rows = bwd_rows_per_cta(M, resident);
  ln_bwd_kernel<N, TIn, TDy, MODE><<<(unsigned)((M + rows - 1) / rows), threads, S::BYTES, stream>>>(
      dy, x, mean, rstd, row_utt, prm.P, prm.stride, g_off, b_off, G ? scratch : nullptr, dx32, dx16, dx_add, M, (int)rows);
  CUDA_TRY(cudaGetLastError());
  if (G && defer) {
    *defer = LnReduceItem{scratch, g_off, b_off, N, (int)rows, nullptr, nullptr};
  } else if (G) {
    ln_bwd_reduce_kernel<<<n_utts, 256, 0, stream>>>(scratch, tok_off, T, (int)rows, N, G, prm.stride, g_off, b_off);
    CUDA_TRY(cudaGetLastError());
  }
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
                      cudaStream_t stream, const float* y32_bias, int mode) {
  SUTA_CHECK_ARG((x_f32 != nullptr) != (x_bf16 != nullptr));
  SUTA_CHECK_ARG(g_off % 4 == 0 && b_off % 4 == 0 && prm.stride % 4 == 0);
  SUTA_CHECK_ARG(mode == LN_PLAIN || (mode == LN_GELU && x_bf16) || (mode == LN_KEEP_INPUT && x_f32));
  if (M <= 0) return SUTA_OK;
  if (mode == LN_GELU) {
    LN_DISPATCH(N, return (launch_fwd<NN, bf16, 1>(x_bf16, row_utt, prm, g_off, b_off, y_f32, y_bf16, mean, rstd, M, eps, y32_bias, stream)));
  } else if (mode == LN_KEEP_INPUT) {
    LN_DISPATCH(N, return (launch_fwd<NN, float, 2>(x_f32, row_utt, prm, g_off, b_off, y_f32, y_bf16, mean, rstd, M, eps, y32_bias, stream)));
  } else if (x_f32) {
    LN_DISPATCH(N, return (launch_fwd<NN, float, 0>(x_f32, row_utt, prm, g_off, b_off, y_f32, y_bf16, mean, rstd, M, eps, y32_bias, stream)));
  } else {
    LN_DISPATCH(N, return (launch_fwd<NN, bf16, 0>(x_bf16, row_utt, prm, g_off, b_off, y_f32, y_bf16, mean, rstd, M, eps, y32_bias, stream)));
  }
  return SUTA_OK;
}

long long layernorm_backward_scratch_floats(int N, int n_utts) { return ((long long)8 * n_sms() + n_utts + 1) * 2 * N; }

int layernorm_backward(const float* dy, const float* x_f32, const bf16* x_bf16, const float* mean, const float* rstd,
                       const int* row_utt, UttParams prm, int g_off, int b_off, float* G, float* dx_f32, bf16* dx_bf16,
                       long long M, int N, const long long* tok_off, const int* T, int n_utts, float* scratch,
                       cudaStream_t stream, LnReduceItem* defer, const float* dx_add) {
  SUTA_CHECK_ARG((x_f32 != nullptr) != (x_bf16 != nullptr));
  SUTA_CHECK_ARG(!G || (tok_off && T && n_utts > 0 && scratch));
  SUTA_CHECK_ARG(!dx_add || x_f32);
  if (M <= 0) return SUTA_OK;
  if (dx_add) {
    LN_DISPATCH(N, return (launch_bwd<NN, float, float, 2>(dy, x_f32, mean, rstd, row_utt, prm, g_off, b_off, G, dx_f32, dx_bf16, dx_add, M, tok_off, T, n_utts, scratch, stream, defer)));
  } else if (x_f32) {
    LN_DISPATCH(N, return (launch_bwd<NN, float, float, 0>(dy, x_f32, mean, rstd, row_utt, prm, g_off, b_off, G, dx_f32, dx_bf16, nullptr, M, tok_off, T, n_utts, scratch, stream, defer)));
  } else {
    LN_DISPATCH(N, return (launch_bwd<NN, bf16, float, 0>(dy, x_bf16, mean, rstd, row_utt, prm, g_off, b_off, G, dx_f32, dx_bf16, nullptr, M, tok_off, T, n_utts, scratch, stream, defer)));
  }
  return SUTA_OK;
}

// backward of GELU(LayerNorm(x)) with a bf16 input x: dy (gradient of the GELU output) is fp32 or bf16, exactly one non-null
int layernorm_gelu_backward(const float* dy_f32, const bf16* dy_bf16, const bf16* x_bf16, const float* mean, const float* rstd,
                            const int* row_utt, UttParams prm, int g_off, int b_off, float* G, float* dx_f32, bf16* dx_bf16,
                            long long M, int N, const long long* tok_off, const int* T, int n_utts, float* scratch,
                            cudaStream_t stream, LnReduceItem* defer) {
  SUTA_CHECK_ARG((dy_f32 != nullptr) != (dy_bf16 != nullptr) && x_bf16);
  SUTA_CHECK_ARG(!G || (tok_off && T && n_utts > 0 && scratch));
  if (M <= 0) return SUTA_OK;
  if (dy_f32) {
    LN_DISPATCH(N, return (launch_bwd<NN, bf16, float, 1>(dy_f32, x_bf16, mean, rstd, row_utt, prm, g_off, b_off, G, dx_f32, dx_bf16, nullptr, M, tok_off, T, n_utts, scratch, stream, defer)));
  } else {
    LN_DISPATCH(N, return (launch_bwd<NN, bf16, bf16, 1>(dy_bf16, x_bf16, mean, rstd, row_utt, prm, g_off, b_off, G, dx_f32, dx_bf16, nullptr, M, tok_off, T, n_utts, scratch, stream, defer)));
  }
  return SUTA_OK;
}

int layernorm_backward_reduce(const LnReduceBatch& b, const long long* tok_off, const int* T, int n_utts, float* G,
                              long long pstride, cudaStream_t stream) {
  SUTA_CHECK_ARG(b.n >= 0 && b.n <= LN_REDUCE_MAX && tok_off && T && n_utts > 0 && G);
  if (b.n == 0) return SUTA_OK;
  ln_bwd_reduce_many_kernel<<<dim3((unsigned)n_utts, (unsigned)b.n), 256, 0, stream>>>(b, tok_off, T, G, pstride);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}
