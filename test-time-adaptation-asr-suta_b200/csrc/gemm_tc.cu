// tcgen05 + TMEM + TMA GEMM for sm_100a.  See gemm_tc.cuh for what it computes.
//
// Structure (one persistent CTA per SM, 10 warps, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor 128x64 (A) and BNx64 (B) bf16 boxes, 128B swizzle,
//               into a STAGES-deep shared-memory ring guarded by full/empty mbarriers
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M=128, N=BN, K=16) x4 per stage into a
//               double-buffered fp32 accumulator in tensor memory; tcgen05.commit releases the stage
//   warps 2..9  epilogue: two warps per TMEM lane quarter, each owning half of the tile's columns;
//               overlaps the next tile's mainloop through the double-buffered accumulator.  Two variants:
//     EPI_TMA   (dense outputs) tcgen05.ld a 32x32 chunk (one row per thread), bias / GELU (+GELU' saved) / x GELU'
//               in registers, write the chunk into a swizzled shared-memory patch and hand it to the TMA:
//               cp.async.bulk.tensor store, or cp.reduce.async.bulk.tensor .add for "out += acc" (the residual
//               stream is accumulated in place at the L2, the SM never loads it).  Nothing in the warp's dependent
//               chain waits on a shared- or global-memory LOAD: measured on B200, any such round trip costs
//               500-3000 cycles while the mainloop saturates the shared-memory and L2 request paths
//               (profiles/r01c_gemm_timeline.md).
//     manual    (row-masked outputs: per-utterance M-block tables, z-batched slabs, BN = 48) the chunk is
//               transposed through a private swizzled patch so lanes run along columns and every global access is
//               a full 128-byte row segment; all loads of a chunk are issued before the first use.
#include "gemm_tc.cuh"
#include <cstdlib>

#include <stdlib.h>
#include <string.h>

#include <mutex>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int GEMM_THREADS = 320;   // TMA warp + MMA warp + 8 epilogue warps
constexpr int SMEM_MAX = 232448;    // 227 KB opt-in limit per CTA

struct GemmKernelParams {
  int M, N, K;
  int out_rows;         // addressable rows of the output (TMA epilogue)
  int num_mblk, num_nblk, nz;
  long long a_z_rows, b_z_rows;
  int c_z_cols;
  const int4* mblk;
  const int4* ztab;
  long long out_z_stride;
  long long* trace;     // debug: per-tile clock64 timeline of CTA 0 (8 slots per tile iteration), or null
  int trace_cap;        // tile iterations that fit
  int b_kwrap, b_tap_col0, b_tap_col1;   // tap-split MN-major B (GemmProblem::b_kwrap)
  int aux_tma;          // TMA epilogue, act == 2, bf16 output: the saved GELU' tile of a chunk arrives by TMA in the patch's idle half
  GemmEpilogue epi;
};

long long* g_trace = nullptr;
int g_trace_cap = 0;

#define TRACE(iter, slot)                                                                              \
  do {                                                                                                 \
    if (p.trace && blockIdx.x == 0 && (iter) < p.trace_cap) p.trace[(iter) * 8 + (slot)] = clock64(); \
  } while (0)

// LEAN (TMA epilogue only): one 4 KB region per epilogue warp -- an fp32 patch (then no bias) or a 2 KB bf16 patch plus the
// warp's bias values -- which leaves room for a 4th 48 KB stage at BN = 256 (the kernel is mainloop-bound with 3)
template <int BN, bool EPI_TMA, bool LEAN = false>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int ACC_STRIDE = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int CH = (BN % 32 == 0) ? 32 : 16;       // epilogue column chunk
  static constexpr int EPI_SPLIT = (BN >= 128) ? 2 : 1;      // epilogue warps per TMEM lane quarter
  // per epilogue warp: TMA variant = two 4 KB patches (ping-pong) + 512 B of bias; manual = one 4 KB transpose patch
  static constexpr int PATCH_BYTES = 4096;
  static constexpr int WARP_EPI_BYTES = (EPI_TMA && !LEAN) ? 2 * PATCH_BYTES : PATCH_BYTES;
  static constexpr int BIAS_BYTES = (EPI_TMA && !LEAN) ? 8 * 512 : 0;
  static constexpr int EPI_BYTES = 8 * WARP_EPI_BYTES + BIAS_BYTES;
  static constexpr int BAR_BYTES = 384;   // up to 8 stages x 2 + 4 accumulator barriers + 16 aux-tile barriers + TMEM slot
  static constexpr int SMEM_BUDGET = SMEM_MAX - EPI_BYTES - BAR_BYTES - 1024 /*align slack*/;
  static constexpr int STAGES = (SMEM_BUDGET / STAGE_BYTES) > 8 ? 8 : (SMEM_BUDGET / STAGE_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
  static_assert(!EPI_TMA || CH == 32, "the TMA epilogue works on 32-column chunks");
  static_assert(!LEAN || EPI_TMA, "LEAN is a variant of the TMA epilogue");
  static_assert(STAGES >= 2, "shared memory budget");
};

// manual path: one (row, 4 consecutive columns) piece of the output; b4 = bias of these columns (zeros when absent),
// r4 = residual (zeros when absent), d2 = 4 saved GELU' values (act == 2)
__device__ __forceinline__ void epilogue4(const GemmEpilogue& e, float4 v, const float4 b4, const float4 r4, const uint2 d2,
                                          long long row, int oc, long long out_off) {
  v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
  if (e.act == 1) {
    float4 d;
    gelu_erf_both2(v.x, v.y, v.x, v.y, d.x, d.y);
    gelu_erf_both2(v.z, v.w, v.z, v.w, d.z, d.w);
    if (e.aux_out)
      *reinterpret_cast<uint2*>(e.aux_out + row * e.aux_ld + oc) = make_uint2(pack_bf16x2(d.x, d.y), pack_bf16x2(d.z, d.w));
  } else if (e.act == 2) {
    const float2 d0 = unpack_bf16x2(d2.x), d1 = unpack_bf16x2(d2.y);
    v.x *= d0.x; v.y *= d0.y; v.z *= d1.x; v.w *= d1.y;
  }
  v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
  if (e.out_f32) *reinterpret_cast<float4*>(e.out_f32 + out_off + row * e.out_ld + oc) = v;
  if (e.out_bf16)
    *reinterpret_cast<uint2*>(e.out_bf16 + out_off + row * e.out_ld + oc) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

// A_MN / B_MN: the operand is MN-major in memory ([K rows][M|N contiguous]); its tile is fetched as 64x64 TMA boxes
// (64 M|N elements = one 128-byte swizzle row, 64 k-rows) laid 8 KB apart, described to the tensor core with
// leading-dimension byte offset 8192 (next 64 M|N elements) and stride byte offset 1024 (next 8 k-rows).
template <int BN, bool A_MN, bool B_MN, bool EPI_TMA, bool LEAN = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_aux,
                    const GemmKernelParams p) {
  using C = GemmCfg<BN, EPI_TMA, LEAN>;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles must start on 1024-byte boundaries
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* epi_smem = tiles + C::STAGES * C::STAGE_BYTES;             // 1024-aligned: STAGE_BYTES is a multiple of 1024
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + C::EPI_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + C::STAGES;          // [STAGES]
  uint64_t* acc_full = bars + 2 * C::STAGES;       // [2]
  uint64_t* acc_empty = bars + 2 * C::STAGES + 2;  // [2]
  uint64_t* aux_bar = bars + 2 * C::STAGES + 4;    // [8 epilogue warps][2]: the chunk's saved GELU' tile has landed (act == 2, TMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4 + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (EPI_TMA) {
      tma_prefetch_desc(&tma_out);
      if (p.epi.aux_out || p.aux_tma) tma_prefetch_desc(&tma_aux);
    }
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 4 * C::EPI_SPLIT);
    }
    for (int i = 0; i < 16; ++i) mbar_init(&aux_bar[i], 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kblk_all = (p.K + BK - 1) / BK;
  const int tiles_per_z = p.num_mblk * p.num_nblk;
  const int total_tiles = tiles_per_z * p.nz;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
        TRACE(iter, 7);
        const int z = tile / tiles_per_z;
        const int r = tile - z * tiles_per_z;
        const int m_blk = r / p.num_nblk;
        const int n_blk = r - m_blk * p.num_nblk;
        int a_row0 = m_blk * BM, b_off = 0;
        if (p.mblk) {
          int4 mi = __ldg(&p.mblk[m_blk]);
          a_row0 = mi.x;
          b_off = mi.w;
        }
        const int a_row = a_row0 + (int)(z * p.a_z_rows);
        const int b_row = b_off + (int)(z * p.b_z_rows) + n_blk * BN;
        int a_k0 = 0, b_k0 = 0, num_kblk = num_kblk_all;
        if (p.ztab) {
          int4 zi = __ldg(&p.ztab[z]);
          a_k0 = zi.x; b_k0 = zi.y; num_kblk = (zi.z + BK - 1) / BK;
        }
        for (int kb = 0; kb < num_kblk; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = tiles + stage * C::STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          if constexpr (A_MN) {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              tma_load_2d(sa + i * 8192, &tma_a, &full_bar[stage], a_row + i * 64, a_k0 + kb * BK);
          } else {
            tma_load_2d(sa, &tma_a, &full_bar[stage], a_k0 + kb * BK, a_row);
          }
          if constexpr (B_MN) {
            int bk = b_k0 + kb * BK, bc = n_blk * BN + (int)(z * p.b_z_rows);
            if (p.b_kwrap) {
              const int t = bk / p.b_kwrap;
              bk -= t * p.b_kwrap;
              bc += t ? p.b_tap_col1 : p.b_tap_col0;
            }
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_2d(sa + C::A_BYTES + i * 8192, &tma_b, &full_bar[stage], bc + i * 64, b_off + bk);
          } else {
            tma_load_2d(sa + C::A_BYTES, &tma_b, &full_bar[stage], b_k0 + kb * BK, b_row);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the whole warp runs the loop, one elected lane issues =====================
    {
      constexpr uint32_t idesc = umma_idesc_bf16(BN) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int iter = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
        int num_kblk = num_kblk_all;
        if (p.ztab) num_kblk = (__ldg(&p.ztab[tile / tiles_per_z]).z + BK - 1) / BK;
        if (lane == 0) TRACE(iter, 0);
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        if (lane == 0) TRACE(iter, 1);
        const uint32_t d_tmem = tmem_base + acc * C::ACC_STRIDE;
        for (int kb = 0; kb < num_kblk; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(tiles + stage * C::STAGE_BYTES);
          // descriptors advance through the stage by adds in the 16-byte-unit address field: 32 B (K-major) or 2048 B (MN-major) per k16
          const uint64_t da = A_MN ? umma_desc_sw128_mn(a_addr) : umma_desc_sw128(a_addr);
          const uint64_t db = B_MN ? umma_desc_sw128_mn(a_addr + C::A_BYTES) : umma_desc_sw128(a_addr + C::A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_ss(d_tmem, da + (A_MN ? 128 : 2) * k, db + (B_MN ? 128 : 2) * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(&empty_bar[stage]);        // frees the smem stage once these MMAs retire
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(&acc_full[acc]);   // accumulator complete -> epilogue
        __syncwarp();
        if (lane == 0) TRACE(iter, 2);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;              // which column half of the tile this warp owns
    constexpr int COLS_PER_WARP = BN / C::EPI_SPLIT;
    constexpr int NCH = COLS_PER_WARP / C::CH;     // chunks per warp and tile
    const uint32_t patch = smem_u32(epi_smem + (warp - 2) * C::WARP_EPI_BYTES);
    int acc = 0;
    uint32_t acc_phase = 0;
    int iter = 0;
    const int tslot = (lane == 0 && (warp == 2 || warp == 6)) ? (warp == 2 ? 3 : 5) : -1;

    if constexpr (EPI_TMA) {
      // ---------------------------------------------------------------- TMA-store epilogue
      const uint32_t bias_s = LEAN ? patch + 2048 : smem_u32(epi_smem + 8 * C::WARP_EPI_BYTES + (warp - 2) * 512);
      const bool out_is_f32 = p.epi.out_f32 != nullptr;
      uint32_t pc = 0;                             // running chunk counter -> patch ping-pong
      for (int tile = blockIdx.x; half < C::EPI_SPLIT && tile < total_tiles; tile += gridDim.x, ++iter) {
        const int m_blk = tile / p.num_nblk;       // nz == 1; with an M-block table every tile OWNS its 128 output rows
        const int n_blk = tile - m_blk * p.num_nblk;
        const int row0 = (p.mblk ? __ldg(&p.mblk[m_blk]).y : m_blk * BM) + q * 32;      // first output row of this warp
        const int col0 = n_blk * BN + half * COLS_PER_WARP;
        if (p.epi.bias) {                          // this warp's COLS_PER_WARP bias values -> shared memory
          __syncwarp();
          for (int i = lane * 4; i < COLS_PER_WARP; i += 128) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.epi.bias + col0 + i));
            st_shared_v4(bias_s + i * 4, __float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w));
          }
          __syncwarp();
        }
        // act == 2 (multiply by the saved GELU').  bf16 output without a LEAN bias: the 32 x 32 tile of a chunk comes by TMA
        // into the idle second half of the chunk's patch (row-per-thread global loads, 16 x 16 B per thread at a multi-KB
        // row pitch, kept the LSU busy for thousands of cycles per tile); otherwise: register prefetch.
        uint64_t* my_aux = aux_bar + (warp - 2) * 2;
        uint8_t* my_patch = epi_smem + (warp - 2) * C::WARP_EPI_BYTES;
        auto aux_issue = [&](uint32_t chunk_no, int kc) {      // lane 0: fetch chunk kc's tile for running chunk number chunk_no
          const int slot = LEAN ? 0 : (int)(chunk_no & 1);
          mbar_expect_tx(&my_aux[slot], 2048);
          tma_load_2d(my_patch + slot * C::PATCH_BYTES + 2048, &tma_aux, &my_aux[slot], col0 + kc * 32, row0);
        };
        if (p.aux_tma && lane == 0) aux_issue(pc, 0);
        uint4 auxr[NCH][4];                        // act == 2 fallback: this thread's row of saved GELU' (32 bf16 per chunk)
        if (p.epi.act == 2 && !p.aux_tma) {
          const bool ok = row0 + lane < p.out_rows;
          const uint4* ap = reinterpret_cast<const uint4*>(p.epi.aux_in + (long long)(row0 + (ok ? lane : 0)) * p.epi.aux_ld + col0);
#pragma unroll
          for (int kc = 0; kc < NCH; ++kc)
#pragma unroll
            for (int j = 0; j < 4; ++j) auxr[kc][j] = ok ? __ldg(ap + kc * 4 + j) : make_uint4(0u, 0u, 0u, 0u);
        }

        mbar_wait(&acc_full[acc], acc_phase);
        tc_fence_after();
        if (tslot >= 0) TRACE(iter, tslot);
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * C::ACC_STRIDE + half * COLS_PER_WARP;
#pragma unroll
        for (int kc = 0; kc < NCH; ++kc) {
          uint32_t raw[32];
          tmem_ld_32x32(t_addr + kc * 32, raw);
          float v[32];
          if (p.epi.bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = ld_shared_v4(bias_s + (kc * 32 + 4 * j) * 4);
              v[4 * j] = b.x; v[4 * j + 1] = b.y; v[4 * j + 2] = b.z; v[4 * j + 3] = b.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
          }
          tmem_ld_wait();
          if (kc == NCH - 1) {                     // accumulator drained: hand the TMEM buffer back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(raw[i]);

          if (p.aux_tma) {
            const int slot = LEAN ? 0 : (int)(pc & 1);
            if (!LEAN) {                           // two patches: the next chunk's tile goes into the other one
              __syncwarp();
              if (lane == 0 && kc + 1 < NCH) aux_issue(pc + 1, kc + 1);
            }
            mbar_wait(&my_aux[slot], LEAN ? (pc & 1) : ((pc >> 1) & 1));
            const uint32_t arow = patch + slot * C::PATCH_BYTES + 2048 + lane * 64;
            const int aswz = (lane >> 1) & 3;
            float4 uf[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) uf[j] = ld_shared_v4(arow + ((j ^ aswz) << 4));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 u = make_uint4(__float_as_uint(uf[j].x), __float_as_uint(uf[j].y), __float_as_uint(uf[j].z), __float_as_uint(uf[j].w));
              const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
              v[8 * j] *= f0.x; v[8 * j + 1] *= f0.y; v[8 * j + 2] *= f1.x; v[8 * j + 3] *= f1.y;
              v[8 * j + 4] *= f2.x; v[8 * j + 5] *= f2.y; v[8 * j + 6] *= f3.x; v[8 * j + 7] *= f3.y;
            }
            if (LEAN) {
              // one patch: refill it only once every lane has CONSUMED this chunk's tile (the multiplies above depend on the
              // loads).  Issuing the refill right after the ld.shared instructions raced with them: under the mainloop's
              // shared-memory traffic a queued ld.shared can take longer than the 2 KB TMA refill (L2 hit), and a few lanes
              // then read the NEXT chunk's GELU' -- dozens of wrong rows per 10^6, different ones in every run
              // (found by tests/test_gpu_e2e.py::test_full_size_batch_is_reproducible_and_utterances_are_independent)
              asm volatile("" ::"f"(v[0]), "f"(v[8]), "f"(v[16]), "f"(v[24]) : "memory");
              __syncwarp();
              if (lane == 0 && kc + 1 < NCH) aux_issue(pc + 1, kc + 1);
            }
          } else if (p.epi.act == 2) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 u = auxr[kc][j];
              const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
              v[8 * j] *= f0.x; v[8 * j + 1] *= f0.y; v[8 * j + 2] *= f1.x; v[8 * j + 3] *= f1.y;
              v[8 * j + 4] *= f2.x; v[8 * j + 5] *= f2.y; v[8 * j + 6] *= f3.x; v[8 * j + 7] *= f3.y;
            }
          }
          const uint32_t pp = LEAN ? patch : patch + (pc & 1) * C::PATCH_BYTES;
          ++pc;
          if (lane == 0) {                         // the store that last read this patch has finished reading it
            if (LEAN) bulk_wait_read<0>(); else bulk_wait_read<1>();
          }
          __syncwarp();
          if (out_is_f32) {
            const uint32_t wrow = pp + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              st_shared_v4(wrow + ((j ^ (lane & 7)) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                           __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
          } else {
            const uint32_t wrow = pp + lane * 64;
            const int swz = (lane >> 1) & 3;
            if (p.epi.act == 1) {
              float d[32];
#pragma unroll
              for (int i = 0; i < 32; i += 2) gelu_erf_both2(v[i], v[i + 1], v[i], v[i + 1], d[i], d[i + 1]);
              if (p.epi.aux_out) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  st_shared_v4(wrow + 2048 + ((j ^ swz) << 4), pack_bf16x2(d[8 * j], d[8 * j + 1]), pack_bf16x2(d[8 * j + 2], d[8 * j + 3]),
                               pack_bf16x2(d[8 * j + 4], d[8 * j + 5]), pack_bf16x2(d[8 * j + 6], d[8 * j + 7]));
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              st_shared_v4(wrow + ((j ^ swz) << 4), pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                           pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (row0 < p.out_rows) {
              const int cc = col0 + kc * 32;
              if (p.epi.accumulate) tma_reduce_add_2d(&tma_out, pp, cc, row0);
              else tma_store_2d(&tma_out, pp, cc, row0);
              if (p.epi.act == 1 && p.epi.aux_out) tma_store_2d(&tma_aux, pp + 2048, cc, row0);
            }
            bulk_commit();
          }
        }
        if (tslot >= 0) TRACE(iter, tslot + 1);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (lane == 0) bulk_wait<0>();               // shared memory must outlive the last store's reads
    } else {
      // ---------------------------------------------------------------- manual (row-masked) epilogue
      constexpr int CGN = C::CH / 4;               // 4-column groups per chunk row (8 or 4)
      constexpr int RPI = 32 / CGN;                // rows covered by one warp-wide instruction in the column phase
      const int cg = lane % CGN, r0 = lane / CGN;
      for (int tile = blockIdx.x; half < C::EPI_SPLIT && tile < total_tiles; tile += gridDim.x, ++iter) {
        const int z = tile / tiles_per_z;
        const int r = tile - z * tiles_per_z;
        const int m_blk = r / p.num_nblk;
        const int n_blk = r - m_blk * p.num_nblk;
        int out_row0 = m_blk * BM, rows_valid = min(BM, p.M - m_blk * BM), b_off = 0;
        if (p.mblk) {
          int4 mi = __ldg(&p.mblk[m_blk]);
          out_row0 = mi.y;
          rows_valid = mi.z;
          b_off = mi.w;
        }
        const int rows_here = rows_valid - q * 32;   // valid rows of this warp's 32-row slab
        const long long row_base = (long long)out_row0 + q * 32;
        const long long bias_off = p.epi.bias_utt_stride ? (long long)(b_off / p.N) * p.epi.bias_utt_stride : (long long)b_off;
        const long long out_off = (long long)z * p.out_z_stride;
        const int oc0 = z * p.c_z_cols + n_blk * BN + half * COLS_PER_WARP + cg * 4;
        float4 bias4[NCH];                           // loaded before the accumulator is awaited
#pragma unroll
        for (int kc = 0; kc < NCH; ++kc) {
          bias4[kc] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.epi.bias) bias4[kc] = __ldg(reinterpret_cast<const float4*>(p.epi.bias + bias_off + oc0 + kc * C::CH));
        }

        mbar_wait(&acc_full[acc], acc_phase);
        tc_fence_after();
        if (tslot >= 0) TRACE(iter, tslot);
        const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * C::ACC_STRIDE + half * COLS_PER_WARP;
#pragma unroll
        for (int kc = 0; kc < NCH; ++kc) {
          uint32_t raw[C::CH];
          if constexpr (C::CH == 32) tmem_ld_32x32(t_addr + kc * C::CH, raw); else tmem_ld_32x16(t_addr + kc * C::CH, raw);
          const int oc = oc0 + kc * C::CH;
          // side inputs of this chunk, issued before anything waits
          float4 res4[CGN];
          uint2 aux2[CGN];
#pragma unroll
          for (int i = 0; i < CGN; ++i) {
            const int rt = r0 + RPI * i;
            const bool ok = rt < rows_here;
            res4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            aux2[i] = make_uint2(0u, 0u);
            if (p.epi.residual && ok) res4[i] = __ldg(reinterpret_cast<const float4*>(p.epi.residual + (row_base + rt) * p.epi.res_ld + oc));
            if (p.epi.act == 2 && ok) aux2[i] = __ldg(reinterpret_cast<const uint2*>(p.epi.aux_in + (row_base + rt) * p.epi.aux_ld + oc));
          }
          tmem_ld_wait();
          if (kc == NCH - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
          }
          if (rows_here > 0) {
            // row-per-thread -> patch (16-byte chunk j of row `lane` at chunk position j ^ swz(lane): conflict-free)
            const uint32_t wrow = patch + lane * (C::CH * 4);
            const int wswz = (lane * CGN / 8) & (CGN - 1);
#pragma unroll
            for (int j = 0; j < CGN; ++j)
              st_shared_v4(wrow + ((j ^ wswz) << 4), raw[4 * j], raw[4 * j + 1], raw[4 * j + 2], raw[4 * j + 3]);
            __syncwarp();
            float4 vv[CGN];
#pragma unroll
            for (int i = 0; i < CGN; ++i) {
              const int rt = r0 + RPI * i;
              const int rswz = (rt * CGN / 8) & (CGN - 1);
              vv[i] = ld_shared_v4(patch + rt * (C::CH * 4) + ((cg ^ rswz) << 4));
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < CGN; ++i) {
              const int rt = r0 + RPI * i;
              if (rt < rows_here) epilogue4(p.epi, vv[i], bias4[kc], res4[i], aux2[i], row_base + rt, oc, out_off);
            }
          }
        }
        if (tslot >= 0) TRACE(iter, tslot + 1);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// generic 2-D map: dims {cols (contiguous), rows}, row pitch in bytes, box {box_cols, box_rows}
int encode_tmap(CUtensorMap* tm, CUtensorMapDataType dt, const void* ptr, long long cols, long long rows,
                long long pitch_bytes, int box_cols, int box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    suta_set_last_error("cuTensorMapEncodeTiled entry point not available");
    return SUTA_ERR_DRIVER;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (strides[0] & 15) || rows <= 0 || cols <= 0) {
    suta_set_last_error("gemm operand not TMA-compatible: ptr=%p pitch=%lld rows=%lld cols=%lld", ptr, pitch_bytes, rows, cols);
    return SUTA_ERR_ARG;
  }
  CUresult r = enc(tm, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    suta_set_last_error("cuTensorMapEncodeTiled failed (%d): cols=%lld rows=%lld pitch=%lld box=%dx%d", (int)r, cols, rows,
                        pitch_bytes, box_cols, box_rows);
    return SUTA_ERR_DRIVER;
  }
  return SUTA_OK;
}

// bf16 GEMM operand: K-major (dims {K, rows}) or MN-major (dims {cols, rows = K}), 64-element (128-byte) box rows
int make_tmap(CUtensorMap* tm, const GemmOperand& op, int K, int box_rows) {
  return encode_tmap(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, op.ptr, K, op.rows, op.row_stride * (long long)sizeof(bf16), BK,
                     box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}
int make_tmap_mn(CUtensorMap* tm, const GemmOperand& op) { return make_tmap(tm, op, (int)op.cols, 64); }

template <int BN, bool A_MN, bool B_MN, bool EPI_TMA, bool LEAN = false>
int launch(const GemmProblem& p, cudaStream_t stream) {
  using C = GemmCfg<BN, EPI_TMA, LEAN>;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(gemm_bf16_tc_kernel<BN, A_MN, B_MN, EPI_TMA, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  C::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap ta, tb, to, tx;
  memset(&to, 0, sizeof(to));
  memset(&tx, 0, sizeof(tx));
  if (A_MN) SUTA_TRY(make_tmap_mn(&ta, p.a)); else SUTA_TRY(make_tmap(&ta, p.a, p.K, BM));
  if (B_MN) SUTA_TRY(make_tmap_mn(&tb, p.b)); else SUTA_TRY(make_tmap(&tb, p.b, p.K, BN));
  if (EPI_TMA) {
    // 32 x 32 output boxes: fp32 rows are 128 B (128B swizzle), bf16 rows 64 B (64B swizzle); rows >= out_rows are clipped
    const long long orows = p.out_rows > 0 ? p.out_rows : p.M;
    if (p.epi.out_f32)
      SUTA_TRY(encode_tmap(&to, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, p.epi.out_f32, p.N, orows, (long long)p.epi.out_ld * 4, 32, 32,
                           CU_TENSOR_MAP_SWIZZLE_128B));
    else
      SUTA_TRY(encode_tmap(&to, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, p.epi.out_bf16, p.N, orows, (long long)p.epi.out_ld * 2, 32, 32,
                           CU_TENSOR_MAP_SWIZZLE_64B));
    if (p.epi.act == 2 && p.epi.out_bf16 && !(LEAN && p.epi.bias) && p.epi.aux_ld % 8 == 0)
      SUTA_TRY(encode_tmap(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, p.epi.aux_in, p.N, orows, (long long)p.epi.aux_ld * 2, 32, 32,
                           CU_TENSOR_MAP_SWIZZLE_64B));
    if (p.epi.act == 1 && p.epi.aux_out)
      SUTA_TRY(encode_tmap(&tx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, p.epi.aux_out, p.N, orows, (long long)p.epi.aux_ld * 2, 32, 32,
                           CU_TENSOR_MAP_SWIZZLE_64B));
  }
  GemmKernelParams kp;
  kp.M = p.M; kp.N = p.N; kp.K = p.K;
  kp.out_rows = p.out_rows > 0 ? (int)p.out_rows : p.M;
  kp.num_mblk = p.mblk ? p.num_mblk : ceil_div(p.M, BM);
  kp.num_nblk = ceil_div(p.N, BN);
  kp.nz = p.nz;
  kp.a_z_rows = p.a_z_rows; kp.b_z_rows = p.b_z_rows; kp.c_z_cols = p.c_z_cols;
  kp.mblk = p.mblk;
  kp.ztab = p.ztab;
  kp.out_z_stride = p.out_z_stride;
  kp.b_kwrap = p.b_kwrap; kp.b_tap_col0 = p.b_tap_col[0]; kp.b_tap_col1 = p.b_tap_col[1];
  kp.aux_tma = (EPI_TMA && p.epi.act == 2 && p.epi.out_bf16 && !(LEAN && p.epi.bias) && p.epi.aux_ld % 8 == 0) ? 1 : 0;
  kp.trace = g_trace;
  kp.trace_cap = g_trace_cap;
  kp.epi = p.epi;
  long long total = (long long)kp.num_mblk * kp.num_nblk * kp.nz;
  if (total <= 0) return SUTA_OK;
  int grid = (int)(total < gemm_num_sms() ? total : gemm_num_sms());
  gemm_bf16_tc_kernel<BN, A_MN, B_MN, EPI_TMA, LEAN><<<grid, GEMM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, to, tx, kp);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

}  // namespace

// tensor-map encoder shared with gemm_tc2.cu: dtype 0 = fp32, 1 = bf16; swizzle 128 / 64 bytes
int gemm_encode_tmap(CUtensorMap* tm, int dtype, const void* ptr, long long cols, long long rows, long long pitch_bytes,
                     int box_cols, int box_rows, int swizzle) {
  return encode_tmap(tm, dtype ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, ptr, cols, rows, pitch_bytes,
                     box_cols, box_rows, swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

// plain (un-swizzled) bf16 tensor map of rank 2 or 3 for kernels that lay operands out themselves (posconv_tc.cu):
// dims / box innermost first, strides in bytes for dims 1..rank-1
int gemm_encode_tmap_nd(CUtensorMap* tm, const void* ptr, int rank, const long long* dims, const long long* strides_bytes,
                        const int* box) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    suta_set_last_error("cuTensorMapEncodeTiled entry point not available");
    return SUTA_ERR_DRIVER;
  }
  SUTA_CHECK_ARG(rank >= 2 && rank <= 3 && !(reinterpret_cast<uintptr_t>(ptr) & 15));
  cuuint64_t d[3], s[2];
  cuuint32_t b[3], es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) { d[i] = (cuuint64_t)dims[i]; b[i] = (cuuint32_t)box[i]; }
  for (int i = 0; i + 1 < rank; ++i) { s[i] = (cuuint64_t)strides_bytes[i]; SUTA_CHECK_ARG((s[i] & 15) == 0); }
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    suta_set_last_error("cuTensorMapEncodeTiled (rank %d, plain) failed (%d)", rank, (int)r);
    return SUTA_ERR_DRIVER;
  }
  return SUTA_OK;
}

// debug hook: CTA 0 of every following GEMM launch writes its per-tile clock64 timeline into dev_buf[cap][8]
// (0 MMA warp waits for a free accumulator, 1 starts issuing, 2 has issued the tile; 3/4 and 5/6 first epilogue
// warp of each column half starts/finishes; 7 producer starts the tile).  Pass null to switch off.
extern "C" void suta_debug_set_gemm_trace(long long* dev_buf, int cap) {
  g_trace = dev_buf;
  g_trace_cap = dev_buf ? cap : 0;
}
long long* gemm_trace_buffer(int* cap) {
  *cap = g_trace_cap;
  return g_trace;
}

int gemm_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int gemm_bf16_tc(const GemmProblem& p, cudaStream_t stream) {
  SUTA_CHECK_ARG(p.M > 0 && p.N > 0 && (p.K > 0 || p.ztab) && p.N % 16 == 0 && p.K % 8 == 0);
  SUTA_CHECK_ARG(p.epi.out_f32 || p.epi.out_bf16);
  SUTA_CHECK_ARG(p.epi.out_ld % 8 == 0 && p.epi.res_ld % 4 == 0 && p.epi.aux_ld % 8 == 0);
  SUTA_CHECK_ARG(!(p.epi.accumulate && (p.epi.residual || !p.epi.out_f32)));
  // dense outputs: every tile owns all 128 of its output rows (plain row blocks, or an M-block table whose owner says so
  // with tiles_own_rows), one z slice, shared bias -> TMA-store epilogue
  const bool dense = (!p.mblk || p.tiles_own_rows) && !p.ztab && p.nz == 1 && !p.epi.residual && !p.epi.bias_utt_stride &&
                     p.N % 32 == 0 && ((p.epi.out_f32 != nullptr) != (p.epi.out_bf16 != nullptr)) &&
                     (p.epi.act != 1 || p.epi.out_bf16);
  if (p.a.mn_major || p.b.mn_major) {
    SUTA_CHECK_ARG(p.N % 64 == 0 && (!p.a.mn_major || p.b.mn_major) && !p.epi.accumulate);
    if (p.a.mn_major) {
      // Weight gradients.  With few tiles (ONE utterance's gradient under train_all: K = its ~300 frames, 4-5 k-blocks) the
      // row-masked fp32 epilogue of a 128 x 256 tile outlasts the mainloop and most SMs idle: take the widest tile that still
      // gives every SM one.  The sum over k of an output element does not depend on the tile width (same bits).
      const long long mb = p.mblk ? p.num_mblk : ceil_div(p.M, BM);
      auto fills = [&](int bn) { return p.N % bn == 0 && mb * (p.N / bn) * p.nz >= gemm_num_sms(); };
      if (fills(256)) return launch<256, true, true, false>(p, stream);
      if (fills(128)) return launch<128, true, true, false>(p, stream);
      return launch<64, true, true, false>(p, stream);
    }
    // SUTA_NO_LEAN_BMN=1: the 3-stage variant (two epilogue patches) instead of the 4-stage LEAN one (debug switch)
    static const bool no_lean_bmn = getenv("SUTA_NO_LEAN_BMN") != nullptr;
    const bool lean = dense && !(p.epi.act == 1 && p.epi.aux_out) && !(p.epi.out_f32 && p.epi.bias) && !no_lean_bmn;
    if (dense) {
      if (p.N % 256 == 0 && lean) return launch<256, false, true, true, true>(p, stream);
      if (p.N % 256 == 0) return launch<256, false, true, true>(p, stream);
      if (p.N % 128 == 0) return launch<128, false, true, true>(p, stream);
      return launch<64, false, true, true>(p, stream);
    }
    if (p.N % 256 == 0) return launch<256, false, true, false>(p, stream);
    if (p.N % 128 == 0) return launch<128, false, true, false>(p, stream);
    return launch<64, false, true, false>(p, stream);
  }
  if (dense) {
    // 4-stage variant when the epilogue needs neither the second patch (GELU' side output) nor an fp32 patch AND a bias
    const bool lean = !(p.epi.act == 1 && p.epi.aux_out) && !(p.epi.out_f32 && p.epi.bias);
    // plain row blocks with N % 256 == 0: CTA pairs on 256 x 256 tiles (gemm_tc2.cu) -- one SM cannot ingest 48 KB per k-block
    static const bool use_pairs = getenv("SUTA_NO_GEMM2") == nullptr;
    if (use_pairs && (!p.mblk || p.mpair) && p.N % 256 == 0 && p.M > 128 && !(p.epi.out_f32 && p.epi.bias)) return gemm_bf16_tc_2cta(p, stream);
    if (p.N % 256 == 0 && lean) return launch<256, false, false, true, true>(p, stream);
    if (p.N % 256 == 0) return launch<256, false, false, true>(p, stream);
    if (p.N % 128 == 0) return launch<128, false, false, true>(p, stream);
    if (p.N % 64 == 0) return launch<64, false, false, true>(p, stream);
    return launch<32, false, false, true>(p, stream);
  }
  SUTA_CHECK_ARG(!p.epi.accumulate);
  if (p.N % 256 == 0) return launch<256, false, false, false>(p, stream);
  if (p.N % 128 == 0) return launch<128, false, false, false>(p, stream);
  if (p.N % 64 == 0) return launch<64, false, false, false>(p, stream);
  if (p.N % 48 == 0) return launch<48, false, false, false>(p, stream);
  if (p.N % 32 == 0) return launch<32, false, false, false>(p, stream);
  suta_set_last_error("gemm: N=%d must be a multiple of 32 or 48", p.N);
  return SUTA_ERR_ARG;
}
