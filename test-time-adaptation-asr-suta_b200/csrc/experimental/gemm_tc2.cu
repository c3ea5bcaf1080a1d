// Two-SM (cta_group::2) tcgen05 GEMM for the dense contractions of the encoder:
//   D[M,N] (+)= A[M,K] B[N,K]^T (+ bias)(GELU, GELU' saved | x saved GELU'),  bf16 operands, fp32 accumulation in TMEM.
//
// A CTA pair (cluster of 2, same TPC) computes one 256 x BN tile: CTA r loads rows [256 p + 128 r, +128) of A and rows
// [BN n + (BN/2) r, +BN/2) of B; the leader CTA's single MMA thread issues tcgen05.mma.cta_group::2 (M = 256), which
// reads both CTAs' shared memory and leaves each CTA the 128 x BN accumulator of ITS rows in its own tensor memory.
// Against the one-SM kernel (gemm_tc.cu) a stage is 32 KB instead of 48 KB (5 stages instead of 3 next to the epilogue
// buffers) and every byte of B is fetched from L2 and read by the tensor core once per pair instead of once per SM --
// the one-SM kernel measured mainloop-bound at 3 stages (profiles/r01c_gemm_timeline.md).
//
// Protocol (per CTA the same code; r = %cluster_ctarank, leader = rank 0):
//   producer warp   plain TMA loads of its A and B halves completing on its OWN full barrier; waits on its OWN empty barrier
//   relay thread    peer CTA only (its otherwise idle MMA warp): waits on the peer's full barrier and forwards ONE remote
//                   arrive per stage to the leader's peer_ready barrier.  (Letting the peer's TMA complete its bytes
//                   directly on the leader's barrier -- UTMALDG.2CTA with a remote mbarrier -- measured ~6000 cycles per
//                   stage round trip: every transaction becomes a cross-SM barrier update.)
//   MMA thread      leader only: waits its full barrier and peer_ready, issues the MMAs, tcgen05.commit ... multicast to
//                   the empty barrier of BOTH CTAs; accumulator-complete commit multicast to both CTAs' acc_full
//   epilogue warps  as in gemm_tc.cu (TMA-store / reduce-add epilogue from the CTA's own TMEM); they release the
//                   accumulator by arriving on the LEADER's acc_empty barrier (count 16)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../gemm_tc.cuh"

namespace {

constexpr int BM = 128;            // rows per CTA (256 per pair)
constexpr int BK = 64;
constexpr int THREADS = 320;
constexpr int SMEM_MAX = 232448;

struct Params2 {
  int M, N, K;
  int num_mpair, num_nblk;
  int out_rows;
  int debug;
  GemmEpilogue epi;
};

template <int BN>
struct Cfg2 {
  static constexpr int A_BYTES = BM * BK * 2;            // 16 KB
  static constexpr int B_BYTES = (BN / 2) * BK * 2;      // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;               // double-buffered accumulator
  static constexpr int COLS_PER_WARP = BN / 2;
  static constexpr int NCH = COLS_PER_WARP / 32;
  static constexpr int PATCH_BYTES = 4096;
  static constexpr int EPI_BYTES = 8 * PATCH_BYTES + 8 * 512;
  static constexpr int BAR_BYTES = 512;   // up to 8 stages x 4 barriers + 4 + slot
  static constexpr int STAGES = ((SMEM_MAX - EPI_BYTES - BAR_BYTES - 1024) / STAGE_BYTES) > 8 ? 8 : ((SMEM_MAX - EPI_BYTES - BAR_BYTES - 1024) / STAGE_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
  static_assert(TMEM_COLS <= 512 && STAGES >= 3, "configuration");
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
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void umma2_commit_local(uint64_t* bar) {    // arrive on the executing (leader) CTA's barrier only
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_remote_expect_tx(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
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
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
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
__host__ __device__ constexpr uint32_t idesc2_bf16(int n) {             // M = 256 (cta_group::2), K-major A and B
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
             const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_aux, const Params2 p) {
  using C = Cfg2<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // same offset in both CTAs
  uint8_t* epi_smem = tiles + C::STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + C::EPI_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES] per CTA: its own TMA bytes
  uint64_t* empty_bar = bars + C::STAGES;          // [STAGES] per CTA
  uint64_t* peer_ready = bars + 2 * C::STAGES;     // [STAGES] used in the leader CTA: the peer's stage has landed
  uint64_t* peer_empty = bars + 3 * C::STAGES;     // [STAGES] used in the peer CTA: the leader saw the stage retire
  uint64_t* acc_full = bars + 4 * C::STAGES;       // [2] per CTA
  uint64_t* acc_empty = bars + 4 * C::STAGES + 2;  // [2] used in the leader CTA (16 arrivals)
  uint64_t* dbg_bar = bars + 4 * C::STAGES + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * C::STAGES + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    tma_prefetch_desc(&tma_out);
    if (p.epi.aux_out) tma_prefetch_desc(&tma_aux);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&peer_ready[s], 1);
      mbar_init(&peer_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 16);
    }
    mbar_init(dbg_bar, 1);
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
      long long nk = 0;                              // k-blocks issued so far
      for (int tile = cid; tile < total_tiles; tile += ncl) {
        const int m_pair = tile / p.num_nblk;
        const int n_blk = tile - m_pair * p.num_nblk;
        const int a_row = (2 * m_pair + (int)rank) * BM;
        const int b_row = n_blk * BN + (int)rank * (BN / 2);
        long long ew = 0, lat[4] = {0, 0, 0, 0};
        const long long pt0 = clock64();
        for (int kb = 0; kb < num_kblk; ++kb) {
          const long long e0 = clock64();
          if (rank == 0) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            // forward "stage retired" to the peer (one remote arrive; not for the initial fill of the ring)
            if (nk >= C::STAGES) mbar_remote_arrive(mapa_rank(smem_u32(&peer_empty[stage]), 1));
          } else {
            mbar_wait_cluster(&peer_empty[stage], phase ^ 1);
          }
          ++nk;
          ew += clock64() - e0;
          uint8_t* sa = tiles + stage * C::STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          tma_load_2d(sa, &tma_a, &full_bar[stage], kb * BK, a_row);
          tma_load_2d(sa + C::A_BYTES, &tma_b, &full_bar[stage], kb * BK, b_row);
          if (p.debug && cid == 0 && tile == ncl && kb >= 8 && kb < 12) {      // debug: load latency seen by the issuing thread
            const long long l0 = clock64();
            mbar_wait(&full_bar[stage], phase);
            lat[kb - 8] = clock64() - l0;
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (p.debug && cid == 0 && tile < 3 * ncl)
          printf("gemm2 producer rank %u tile %d: %lld cycles, waiting for free stages %lld; load latencies %lld %lld %lld %lld\n", rank, tile,
                 clock64() - pt0, ew, lat[0], lat[1], lat[2], lat[3]);
      }
    }
  } else if (warp == 1) {
    // ===================== relay (peer CTA): forward "my stage has landed" to the leader =====================
    if (lane == 0 && rank != 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cid; tile < total_tiles; tile += ncl)
        for (int kb = 0; kb < num_kblk; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          mbar_remote_arrive(mapa_rank0(smem_u32(&peer_ready[stage])));
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
    }
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = idesc2_bf16(BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int titer = 0;
      for (int tile = cid; tile < total_tiles; tile += ncl, ++titer) {
        const long long d0 = clock64();
        mbar_wait_cluster(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const long long d1 = clock64();
        long long dwait = 0;
        const uint32_t d_tmem = tmem_base + acc * BN;
        if (p.debug && cid == 0 && titer == 0) {      // debug: sustained MMA rate on stages that are already resident
          for (int s2 = 0; s2 < C::STAGES; ++s2) { mbar_wait(&full_bar[s2], 0); mbar_wait_cluster(&peer_ready[s2], 0); }
          tc_fence_after();
          const long long m0 = clock64();
          for (int rep = 0; rep < 4; ++rep)
            for (int s2 = 0; s2 < C::STAGES; ++s2) {
              const uint32_t a2 = smem_u32(tiles + s2 * C::STAGE_BYTES), b2 = a2 + C::A_BYTES;
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma2_bf16_ss(tmem_base + (acc ^ 1) * BN, umma_desc_sw128(a2 + k * 32), umma_desc_sw128(b2 + k * 32), idesc, 1u);
            }
          const long long m1 = clock64();
          umma2_commit_local(dbg_bar);
          mbar_wait(dbg_bar, 0);
          printf("gemm2 debug: %d MMAs issued in %lld cycles, retired after %lld cycles\n", 4 * C::STAGES * 4, m1 - m0, clock64() - m0);
        }
        for (int kb = 0; kb < num_kblk; ++kb) {
          const long long w0 = clock64();
          mbar_wait(&full_bar[stage], phase);
          mbar_wait_cluster(&peer_ready[stage], phase);
          dwait += clock64() - w0;
          tc_fence_after();
          const uint32_t a_addr = smem_u32(tiles + stage * C::STAGE_BYTES);
          const uint32_t b_addr = a_addr + C::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma2_bf16_ss(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
          umma2_commit_local(&empty_bar[stage]);     // stage retired; the leader's producer forwards it to the peer
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma2_commit_both(&acc_full[acc]);           // accumulator complete in both CTAs
        if (p.debug && cid == 0 && titer < 4)
          printf("gemm2 tile %d: acc_empty wait %lld, mainloop %lld cycles for %d k-blocks, of which waiting for data %lld\n", titer,
                 d1 - d0, clock64() - d1, num_kblk, dwait);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9, both CTAs): TMA store / reduce-add from the CTA's own TMEM ==========
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t patch = smem_u32(epi_smem + (warp - 2) * C::PATCH_BYTES);
    const uint32_t bias_s = smem_u32(epi_smem + 8 * C::PATCH_BYTES + (warp - 2) * 512);
    const bool out_is_f32 = p.epi.out_f32 != nullptr;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cid; tile < total_tiles; tile += ncl) {
      const int m_pair = tile / p.num_nblk;
      const int n_blk = tile - m_pair * p.num_nblk;
      const int row0 = (2 * m_pair + (int)rank) * BM + q * 32;
      const int col0 = n_blk * BN + half * C::COLS_PER_WARP;
      if (p.epi.bias) {
        __syncwarp();
        for (int i = lane * 4; i < C::COLS_PER_WARP; i += 128) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.epi.bias + col0 + i));
          st_shared_v4(bias_s + i * 4, __float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w));
        }
        __syncwarp();
      }
      uint4 auxr[C::NCH][4];
      if (p.epi.act == 2) {
        const bool ok = row0 + lane < p.out_rows;
        const uint4* ap = reinterpret_cast<const uint4*>(p.epi.aux_in + (long long)(row0 + (ok ? lane : 0)) * p.epi.aux_ld + col0);
#pragma unroll
        for (int kc = 0; kc < C::NCH; ++kc)
#pragma unroll
          for (int j = 0; j < 4; ++j) auxr[kc][j] = ok ? __ldg(ap + kc * 4 + j) : make_uint4(0u, 0u, 0u, 0u);
      }
      mbar_wait_cluster(&acc_full[acc], acc_phase);
      tc_fence_after();
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
          if (lane == 0) mbar_remote_arrive(mapa_rank0(smem_u32(&acc_empty[acc])));
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(raw[i]);
        if (p.epi.act == 2) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 u = auxr[kc][j];
            const float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
            v[8 * j] *= f0.x; v[8 * j + 1] *= f0.y; v[8 * j + 2] *= f1.x; v[8 * j + 3] *= f1.y;
            v[8 * j + 4] *= f2.x; v[8 * j + 5] *= f2.y; v[8 * j + 6] *= f3.x; v[8 * j + 7] *= f3.y;
          }
        }
        if (lane == 0) bulk_wait_read<0>();        // the previous chunk's store has finished reading the patch
        __syncwarp();
        if (out_is_f32) {
          const uint32_t wrow = patch + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(wrow + ((j ^ (lane & 7)) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                         __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
        } else {
          const uint32_t wrow = patch + lane * 64;
          const int swz = (lane >> 1) & 3;
          if (p.epi.act == 1) {
            float d[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) gelu_erf_both(v[i], v[i], d[i]);
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
            if (p.epi.accumulate) tma_reduce_add_2d(&tma_out, patch, cc, row0);
            else tma_store_2d(&tma_out, patch, cc, row0);
            if (p.epi.act == 1 && p.epi.aux_out) tma_store_2d(&tma_aux, patch + 2048, cc, row0);
          }
          bulk_commit();
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // nobody leaves while the peer may still touch this CTA's memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2<C::TMEM_COLS>(tmem_base);
  }
}

template <int BN>
int launch2(const GemmProblem& p, cudaStream_t stream) {
  using C = Cfg2<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(gemm2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
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
  if (p.epi.act == 1 && p.epi.aux_out)
    SUTA_TRY(gemm_encode_tmap(&tx, 1, p.epi.aux_out, p.N, orows, (long long)p.epi.aux_ld * 2, 32, 32, 64));
  Params2 kp;
  kp.M = p.M; kp.N = p.N; kp.K = p.K;
  kp.num_mpair = ceil_div(ceil_div(p.M, BM), 2);
  kp.num_nblk = p.N / BN;
  kp.out_rows = (int)orows;
  static const bool dbg = getenv("SUTA_GEMM2_DEBUG") != nullptr;
  static int dbg_left = 3;
  kp.debug = dbg && dbg_left-- > 0;
  kp.epi = p.epi;
  const long long total = (long long)kp.num_mpair * kp.num_nblk;
  const int pairs = gemm_num_sms() / 2;
  const int grid = 2 * (int)(total < pairs ? total : pairs);
  gemm2_kernel<BN><<<grid, THREADS, C::SMEM_BYTES, stream>>>(ta, tb, to, tx, kp);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

}  // namespace

// dense, K-major, un-batched problems whose N is a multiple of 128 (eligibility is decided by gemm_bf16_tc)
int gemm_bf16_tc_2cta(const GemmProblem& p, cudaStream_t stream) {
  if (p.N % 256 == 0) return launch2<256>(p, stream);
  return launch2<128>(p, stream);
}
