// Two-SM (cta_group::2) tcgen05 GEMM for the dense K-major contractions of the encoder:
//   D[M,N] (+)= A[M,K] B[N,K]^T (+ bias)(GELU, GELU' saved | x saved GELU'),  bf16 operands, fp32 accumulation in TMEM.
//
// Why: one SM ingests at most ~70 B/clk through TMA (tools/ubench/mma2.cu mode 6: 32 KB per ~470 cycles per SM, the same
// rate the one-SM kernel shows at 48 KB per ~690 cycles), so a 128x256 tile fed by one SM (48 KB per 512-cycle k-block =
// 96 B/clk) is feed-bound at ~74 % of the tensor pipe no matter how deep the ring is.  A CTA PAIR computing a 256 x 256
// tile needs 16 KB of A + 16 KB of B per SM per k-block (64 B/clk): each SM loads its own 128 rows of A and HALF of the
// B tile, and the leader's tcgen05.mma.cta_group::2 (M = 256) reads both halves.
//
// Protocol (CUTLASS's 2-SM scheme, measured in tools/ubench/mma2.cu at ~510 cycles per k-block):
//   producer (one thread per CTA)  waits its OWN empty barrier, issues its two TMA loads with .cta_group::2 so the bytes
//                                  complete on the LEADER's full barrier; the leader's producer posts expect_tx for both CTAs
//   MMA warp (leader CTA only)     waits full, one elected lane issues 4 MMAs, tcgen05.commit ... multicast releases the
//                                  stage on BOTH CTAs' empty barriers; accumulator-complete commit multicast to both acc_full
//   epilogue warps (both CTAs)     drain their CTA's 128 x 256 accumulator from its own TMEM exactly like gemm_tc.cu's
//                                  lean TMA-store / reduce-add epilogue; they hand the accumulator back with a (remote)
//                                  arrive on the LEADER's acc_empty barrier (16 arrivals)
// Waiting: try_wait everywhere except acc_empty (arrivals are remote mbarrier.arrive: a sleeping waiter is woken late by
// those, tools/ubench/pingpong.cu).  A relayed variant (peer-local full barrier + forwarded arrive) measured 1300-1500
// cycles per k-block and is gone.
#include <stdlib.h>
#include <string.h>

#include "gemm_tc.cuh"

namespace {

constexpr int BM = 128;            // rows per CTA (256 per pair)
constexpr int BN = 256;
constexpr int BK = 64;
constexpr int THREADS = 320;
constexpr int SMEM_MAX = 232448;

struct Params2 {
  int M, N, K;
  int num_mpair, num_nblk;
  int out_rows;
  int aux_tma;          // act == 2 with a bf16 output: the saved GELU' tile of a chunk arrives by TMA in the patch's idle half
  int tail_tiles, tail_split;   // reduce-add epilogue only: the last (partial) wave's tiles are split tail_split ways along K
  const int4* mpair;    // optional per-pair-tile table {a_row0, out_row0, rows_valid, b_row_off}
  long long* trace;     // debug: clock64 timeline of the leader CTA of pair 0 (same 8 slots per tile as gemm_tc.cu)
  int trace_cap;
  GemmEpilogue epi;
};

#define TRACE2(iter, slot)                                                                                          \
  do {                                                                                                              \
    if (p.trace && blockIdx.x == 0 && (iter) < p.trace_cap) p.trace[(iter) * 8 + (slot)] = clock64();               \
  } while (0)

// Work of one CTA pair: whole 256 x 256 tiles strided over the pairs; with the reduce-add epilogue (partial sums need no
// fix-up pass) the tiles of the last, partial wave are split along K so that wave costs 1/tail_split of a tile time
// instead of a whole one (255 tiles on 74 pairs: 3.5 tile times instead of 4).  Splitting EVERY tile (stream-K) was
// slower: pairs sharing an A tile stop walking k in step and lose each other's L2 hits.
struct WorkIter2 {
  int tile, kb0, kb1;
  int next_tile, stride, full_tiles, KB, tail_split, tail_units, cid;
  bool tail_done;
  __device__ __forceinline__ WorkIter2(const int total_tiles, int KB_, int cid_, int ncl, int tail_tiles, int tail_split_)
      : tile(0), kb0(0), kb1(KB_), next_tile(cid_), stride(ncl), full_tiles(total_tiles - tail_tiles), KB(KB_),
        tail_split(tail_split_), tail_units(tail_tiles * tail_split_), cid(cid_), tail_done(false) {}
  __device__ __forceinline__ bool next() {
    if (next_tile < full_tiles) {
      tile = next_tile;
      next_tile += stride;
      kb0 = 0;
      kb1 = KB;
      return true;
    }
    if (tail_done || cid >= tail_units) return false;
    tail_done = true;
    tile = full_tiles + cid / tail_split;
    const int part = cid - (cid / tail_split) * tail_split;
    kb0 = (int)((long long)KB * part / tail_split);
    kb1 = (int)((long long)KB * (part + 1) / tail_split);
    return kb1 > kb0;
  }
};

// Shared memory: 5 stages of 32 KB (tools/ubench/mma2.cu: 5 and 6 stages run the mainloop equally fast, 4 do not), two
// 4 KB epilogue patches per epilogue warp (ping-pong: chunk c+1 is converted while the TMA store of chunk c still reads its
// patch -- with K = 768 the epilogue, not the mainloop, bounds the tile rate), 2 x 2 x 512 B of bias (per column half, double
// buffered by tile parity).  AUX (GELU' side output): a patch holds the 2 KB bf16 output chunk and the 2 KB GELU' chunk.
struct Cfg2 {
  static constexpr int A_BYTES = BM * BK * 2;            // 16 KB
  static constexpr int B_BYTES = (BN / 2) * BK * 2;      // this CTA's half of the B tile, 16 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;               // double-buffered accumulator
  static constexpr int COLS_PER_WARP = BN / 2;
  static constexpr int NCH = COLS_PER_WARP / 32;
  static constexpr int PATCH_BYTES = 4096;
  static constexpr int WARP_EPI_BYTES = 2 * PATCH_BYTES;
  static constexpr int BIAS_BYTES = 2 * 2 * 512;
  static constexpr int EPI_BYTES = 8 * WARP_EPI_BYTES + BIAS_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int STAGES = 5;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;   // dynamic smem base must be 1024-aligned
  static_assert(SMEM_BYTES <= SMEM_MAX, "configuration");
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank0(uint32_t smem_addr) {   // same offset in the leader CTA's shared memory
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(0));
  return r;
}
__device__ __forceinline__ void mbar_remote_arrive(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t smem_dst, const void* tmap, uint32_t cluster_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(cluster_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {     // arrive on this barrier in both CTAs of the pair
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__host__ __device__ constexpr uint32_t idesc2_bf16(int n) {             // M = 256 (cta_group::2), K-major A and B
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <bool AUX>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
             const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_aux, const Params2 p) {
  using C = Cfg2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if (smem_u32(smem_raw) & 1023u) __trap();        // 128B-swizzled tiles need 1024-byte alignment; no slack is budgeted
  uint8_t* tiles = smem_raw;
  uint8_t* epi_smem = tiles + C::STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + C::EPI_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES] the LEADER's collect both CTAs' bytes
  uint64_t* empty_bar = bars + C::STAGES;          // [STAGES] per CTA, released by the multicast commit
  uint64_t* acc_full = bars + 2 * C::STAGES;       // [2] per CTA (multicast commit)
  uint64_t* acc_empty = bars + 2 * C::STAGES + 2;  // [2] used in the leader CTA (16 arrivals)
  uint64_t* aux_bar = bars + 2 * C::STAGES + 4;    // [8 epilogue warps][2] per CTA: the chunk's saved GELU' tile has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4 + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_out);
    if (AUX || p.aux_tma) tma_prefetch_desc(&tma_aux);
    for (int i = 0; i < 16; ++i) mbar_init(&aux_bar[i], 1);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 16);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc2<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // both CTAs' barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kblk = (p.K + BK - 1) / BK;
  const int total_tiles = p.num_mpair * p.num_nblk;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t leader_full0 = mapa_rank0(smem_u32(&full_bar[0]));
      int iter = 0;
      for (WorkIter2 w(total_tiles, num_kblk, cid, ncl, p.tail_tiles, p.tail_split); w.next(); ++iter) {
        TRACE2(iter, 7);
        const int tile = w.tile;
        const int m_pair = tile / p.num_nblk;
        const int n_blk = tile - m_pair * p.num_nblk;
        int a_row = (2 * m_pair + (int)rank) * BM;
        int b_row = n_blk * BN + (int)rank * (BN / 2);
        if (p.mpair) {
          const int4 mi = __ldg(&p.mpair[m_pair]);
          a_row = mi.x + (int)rank * BM;
          b_row += mi.w;
        }
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * C::STAGE_BYTES);
          const uint32_t sa = smem_u32(tiles + stage * C::STAGE_BYTES);
          const uint32_t lb = leader_full0 + stage * 8;
          tma_load_2d_2cta(sa, &tma_a, lb, kb * BK, a_row);
          tma_load_2d_2cta(sa + C::A_BYTES, &tma_b, lb, kb * BK, b_row);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA; converged warp, one elected lane issues) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = idesc2_bf16(BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int iter = 0;
      for (WorkIter2 w(total_tiles, num_kblk, cid, ncl, p.tail_tiles, p.tail_split); w.next(); ++iter) {
        if (lane == 0) TRACE2(iter, 0);
        mbar_wait_cluster(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        if (lane == 0) TRACE2(iter, 1);
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(tiles + stage * C::STAGE_BYTES);
          const uint64_t da = umma_desc_sw128(a_addr), db = umma_desc_sw128(a_addr + C::A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma2_bf16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb != w.kb0 || k != 0) ? 1u : 0u);
            umma2_commit_both(&empty_bar[stage]);  // the stage is free in both CTAs once these MMAs retire
          }
          __syncwarp();
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma2_commit_both(&acc_full[acc]);   // accumulator complete in both CTAs
        __syncwarp();
        if (lane == 0) TRACE2(iter, 2);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9, both CTAs): TMA store / reduce-add from the CTA's own TMEM ==========
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t patch = smem_u32(epi_smem + (warp - 2) * C::WARP_EPI_BYTES);
    const uint32_t bias_base = smem_u32(epi_smem + 8 * C::WARP_EPI_BYTES) + half * 512;
    uint32_t pc = 0;                               // running chunk counter -> patch ping-pong
    const bool out_is_f32 = p.epi.out_f32 != nullptr;
    const uint32_t leader_acc_empty0 = mapa_rank0(smem_u32(&acc_empty[0]));
    int acc = 0;
    uint32_t acc_phase = 0;
    int iter = 0;
    const int tslot = (lane == 0 && (warp == 2 || warp == 6)) ? (warp == 2 ? 3 : 5) : -1;
    for (WorkIter2 w(total_tiles, num_kblk, cid, ncl, p.tail_tiles, p.tail_split); w.next(); ++iter) {
      const int tile = w.tile;
      const int m_pair = tile / p.num_nblk;
      const int n_blk = tile - m_pair * p.num_nblk;
      const int row0 = (p.mpair ? __ldg(&p.mpair[m_pair]).y : 2 * m_pair * BM) + (int)rank * BM + q * 32;
      const int col0 = n_blk * BN + half * C::COLS_PER_WARP;
      float4 breg = make_float4(0.f, 0.f, 0.f, 0.f);   // this lane's 4 of the half's 128 bias values (bf16 outputs only)
      if (p.epi.bias) breg = __ldg(reinterpret_cast<const float4*>(p.epi.bias + col0 + lane * 4));
      // act == 2 (multiply by the saved GELU').  bf16 output: the 32 x 32 tile of a chunk is fetched by TMA into the idle
      // half of the chunk's patch, one chunk ahead (row-per-thread global loads -- 16 x 16 B per thread at a 6 KB row
      // pitch -- kept the LSU busy for ~4 k cycles per tile); fp32 output (patch fully used): register prefetch as before.
      uint64_t* my_aux = aux_bar + (warp - 2) * 2;
      uint8_t* my_patch = epi_smem + (warp - 2) * C::WARP_EPI_BYTES;
      const bool aux_tma = !AUX && p.aux_tma;        // compile-time off in the GELU'-writing variant (it spilled otherwise)
      if (aux_tma) {
        if (lane == 0) {
          mbar_expect_tx(&my_aux[pc & 1], 2048);
          tma_load_2d(my_patch + (pc & 1) * C::PATCH_BYTES + 2048, &tma_aux, &my_aux[pc & 1], col0, row0);
        }
      }
      uint4 auxr[C::NCH][4];
      if (!AUX && p.epi.act == 2 && !aux_tma) {
        const bool ok = row0 + lane < p.out_rows;
        const uint4* ap = reinterpret_cast<const uint4*>(p.epi.aux_in + (long long)(row0 + (ok ? lane : 0)) * p.epi.aux_ld + col0);
#pragma unroll
        for (int kc = 0; kc < C::NCH; ++kc)
#pragma unroll
          for (int j = 0; j < 4; ++j) auxr[kc][j] = ok ? __ldg(ap + kc * 4 + j) : make_uint4(0u, 0u, 0u, 0u);
      }
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      if (tslot >= 0) TRACE2(iter, tslot);
      // the four warps of a column half share one bias buffer per tile parity and all write the same values; tile i + 2
      // reuses tile i's buffer, and acc_full[i + 2] implies every warp released accumulator i, i.e. finished reading it
      const uint32_t bias_s = bias_base + (iter & 1) * 1024;
      if (p.epi.bias) {
        st_shared_v4(bias_s + lane * 16, __float_as_uint(breg.x), __float_as_uint(breg.y), __float_as_uint(breg.z), __float_as_uint(breg.w));
        __syncwarp();
      }
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + half * C::COLS_PER_WARP;
#pragma unroll
      for (int kc = 0; kc < C::NCH; ++kc) {
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
        if (kc == C::NCH - 1) {                    // accumulator drained: release it on the leader's barrier
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_remote_arrive(leader_acc_empty0 + acc * 8);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(raw[i]);
        if (aux_tma) {
          __syncwarp();                            // everyone has read the other patch's tile (the previous chunk's)
          if (lane == 0 && kc + 1 < C::NCH) {
            mbar_expect_tx(&my_aux[(pc + 1) & 1], 2048);
            tma_load_2d(my_patch + ((pc + 1) & 1) * C::PATCH_BYTES + 2048, &tma_aux, &my_aux[(pc + 1) & 1], col0 + (kc + 1) * 32, row0);
          }
          mbar_wait(&my_aux[pc & 1], (pc >> 1) & 1);
          const uint32_t arow = patch + (pc & 1) * C::PATCH_BYTES + 2048 + lane * 64;
          const int aswz = (lane >> 1) & 3;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 uf = ld_shared_v4(arow + ((j ^ aswz) << 4));
            const uint4 u = make_uint4(__float_as_uint(uf.x), __float_as_uint(uf.y), __float_as_uint(uf.z), __float_as_uint(uf.w));
            const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
            v[8 * j] *= f0.x; v[8 * j + 1] *= f0.y; v[8 * j + 2] *= f1.x; v[8 * j + 3] *= f1.y;
            v[8 * j + 4] *= f2.x; v[8 * j + 5] *= f2.y; v[8 * j + 6] *= f3.x; v[8 * j + 7] *= f3.y;
          }
        } else if (!AUX && p.epi.act == 2) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 u = auxr[kc][j];
            const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
            v[8 * j] *= f0.x; v[8 * j + 1] *= f0.y; v[8 * j + 2] *= f1.x; v[8 * j + 3] *= f1.y;
            v[8 * j + 4] *= f2.x; v[8 * j + 5] *= f2.y; v[8 * j + 6] *= f3.x; v[8 * j + 7] *= f3.y;
          }
        }
        const uint32_t pp = patch + (pc & 1) * C::PATCH_BYTES;
        ++pc;
        if (lane == 0) bulk_wait_read<1>();        // the store that last read THIS patch (two chunks ago) has finished reading
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
            if (AUX) {
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
            if (AUX) tma_store_2d(&tma_aux, pp + 2048, cc, row0);
          }
          bulk_commit();
        }
      }
      if (tslot >= 0) TRACE2(iter, tslot + 1);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait<0>();                 // shared memory must outlive the last store's reads
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // nobody leaves while the peer may still touch this CTA's memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2<C::TMEM_COLS>(tmem_base);
  }
}

template <bool AUX>
int launch2(const GemmProblem& p, cudaStream_t stream) {
  using C = Cfg2;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(gemm2_kernel<AUX>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap ta, tb, to, tx;
  memset(&tx, 0, sizeof(tx));
  SUTA_TRY(gemm_encode_tmap(&ta, 1, p.a.ptr, p.K, p.a.rows, p.a.row_stride * 2, BK, BM, 128));
  SUTA_TRY(gemm_encode_tmap(&tb, 1, p.b.ptr, p.K, p.b.rows, p.b.row_stride * 2, BK, BN / 2, 128));
  const long long orows = p.out_rows > 0 ? p.out_rows : p.M;
  if (p.epi.out_f32)
    SUTA_TRY(gemm_encode_tmap(&to, 0, p.epi.out_f32, p.N, orows, (long long)p.epi.out_ld * 4, 32, 32, 128));
  else
    SUTA_TRY(gemm_encode_tmap(&to, 1, p.epi.out_bf16, p.N, orows, (long long)p.epi.out_ld * 2, 32, 32, 64));
  if (AUX) SUTA_TRY(gemm_encode_tmap(&tx, 1, p.epi.aux_out, p.N, orows, (long long)p.epi.aux_ld * 2, 32, 32, 64));
  const bool aux_tma = !AUX && p.epi.act == 2 && p.epi.out_bf16 && p.epi.aux_ld % 8 == 0;
  if (aux_tma) SUTA_TRY(gemm_encode_tmap(&tx, 1, p.epi.aux_in, p.N, orows, (long long)p.epi.aux_ld * 2, 32, 32, 64));
  Params2 kp;
  kp.M = p.M; kp.N = p.N; kp.K = p.K;
  kp.num_mpair = p.mpair ? p.num_mpair : ceil_div(ceil_div(p.M, BM), 2);
  kp.mpair = p.mpair;
  kp.aux_tma = aux_tma ? 1 : 0;
  kp.num_nblk = p.N / BN;
  kp.out_rows = (int)orows;
  kp.trace = gemm_trace_buffer(&kp.trace_cap);
  kp.epi = p.epi;
  const long long total = (long long)kp.num_mpair * kp.num_nblk;
  const int pairs = gemm_num_sms() / 2;
  const int grid = 2 * (int)(total < pairs ? total : pairs);
  kp.tail_tiles = 0;
  kp.tail_split = 1;
  static const bool tail_on = getenv("SUTA_NO_TAIL_SPLIT") == nullptr;
  const int tail = (int)(total % pairs), kblocks = ceil_div(p.K, BK);
  if (tail_on && p.epi.accumulate && p.epi.out_f32 && !p.epi.bias && total > pairs && tail > 0 && 2 * tail <= pairs) {
    int split = pairs / tail;
    while (split > 1 && kblocks / split < 4) --split;     // keep at least 4 k-blocks per part
    if (split > 1) { kp.tail_tiles = tail; kp.tail_split = split; }
  }
  gemm2_kernel<AUX><<<grid, THREADS, C::SMEM_BYTES, stream>>>(ta, tb, to, tx, kp);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

}  // namespace

// Eligibility (checked by gemm_bf16_tc): dense, K-major, un-batched, plain row blocks or an M-PAIR table, N % 256 == 0, exactly one output,
// no residual operand, a bias only with a bf16 output.
int gemm_bf16_tc_2cta(const GemmProblem& p, cudaStream_t stream) {
  if (p.epi.act == 1 && p.epi.aux_out) return launch2<true>(p, stream);
  return launch2<false>(p, stream);
}
