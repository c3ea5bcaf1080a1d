// Mainloop-only micro-benchmark of the two-SM (cta_group::2) tcgen05 GEMM pipeline: TMA producer in both CTAs of a pair,
// leader-issued 256 x 256 x 16 MMAs, canonical barrier protocol (both CTAs' TMA bytes complete on the LEADER's full barrier,
// tcgen05.commit multicast releases the stage in both CTAs).  No epilogue: the accumulators are never read.
//   mode bit 0: no TMA traffic after the ring is filled once      bit 1: no MMAs (commit only)
//   mode bit 4: the MMA warp waits with try_wait too
//   mode bit 2: producers wait with try_wait instead of polling   bit 3: peer's bytes complete on its own barrier + relay arrive
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o mma2 mma2.cu -lcuda
#include "../../test-time-adaptation-asr-suta_b200/csrc/common.cuh"
#include <cstdlib>
#include <cudaTypedefs.h>

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = (BN / 2) * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
#ifndef STAGES
#define STAGES 6
#endif

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t a, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }
__device__ __forceinline__ void mbar_remote_arrive(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t smem_dst, const void* tmap, uint32_t cluster_bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(cluster_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma2_bf16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void umma2_commit_local(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int NCOLS> __device__ __forceinline__ void tmem_alloc2(uint32_t* dst) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS> __device__ __forceinline__ void tmem_dealloc2(uint32_t t) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(t), "n"(NCOLS) : "memory");
}
__host__ __device__ constexpr uint32_t idesc2_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
__device__ __forceinline__ void wait_any(uint64_t* bar, uint32_t parity, bool sleepy) {
  if (sleepy) mbar_wait(bar, parity); else mbar_wait_cluster(bar, parity);
}

struct P { int K, num_nblk, total_tiles, mode; long long* out; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k2(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const P p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES);
  uint64_t* full_bar = bars;                 // leader's are used (both CTAs' bytes); with mode bit 3 each CTA uses its own
  uint64_t* empty_bar = bars + STAGES;       // per CTA, released by the multicast commit
  uint64_t* peer_ready = bars + 2 * STAGES;  // mode bit 3: leader side, relay arrive
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const bool no_tma = p.mode & 1, no_mma = p.mode & 2, sleepy = p.mode & 4, relay = p.mode & 8;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a); tma_prefetch_desc(&tma_b);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); mbar_init(&peer_ready[s], 1); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc2<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_kblk = p.K / BK;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0; long long nk = 0;
      const uint32_t leader_full0 = mapa_rank(smem_u32(&full_bar[0]), 0);
      for (int tile = cid; tile < p.total_tiles; tile += ncl) {
        const int m_pair = tile / p.num_nblk, n_blk = tile - m_pair * p.num_nblk;
        const int a_row = (2 * m_pair + (int)rank) * BM, b_row = n_blk * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < num_kblk; ++kb, ++nk) {
          wait_any(&empty_bar[stage], phase ^ 1, sleepy);
          const uint32_t sa = smem_u32(tiles + stage * STAGE_BYTES);
          if (no_tma && nk >= STAGES) {
            if (relay) mbar_arrive(&full_bar[stage]); else if (rank == 0) mbar_arrive(&full_bar[stage]);
          } else if (relay) {
            mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
            tma_load_2d(tiles + stage * STAGE_BYTES, &tma_a, &full_bar[stage], kb * BK, a_row);
            tma_load_2d(tiles + stage * STAGE_BYTES + A_BYTES, &tma_b, &full_bar[stage], kb * BK, b_row);
          } else {
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
            const uint32_t lb = leader_full0 + stage * 8;
            tma_load_2d_2cta(sa, &tma_a, lb, kb * BK, a_row);
            tma_load_2d_2cta(sa + A_BYTES, &tma_b, lb, kb * BK, b_row);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (relay && rank != 0) {
      if (lane == 0) {
        int stage = 0; uint32_t phase = 0;
        for (int tile = cid; tile < p.total_tiles; tile += ncl)
          for (int kb = 0; kb < num_kblk; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            mbar_remote_arrive(mapa_rank(smem_u32(&peer_ready[stage]), 0));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
      }
    }
    if (rank == 0) {
      constexpr uint32_t idesc = idesc2_bf16(BN);
      int stage = 0; uint32_t phase = 0; int iter = 0;
      for (int tile = cid; tile < p.total_tiles; tile += ncl, ++iter) {
        const long long t0 = clock64();
        long long waited = 0;
        for (int kb = 0; kb < num_kblk; ++kb) {
          const long long w0 = clock64();
          wait_any(&full_bar[stage], phase, (sleepy && relay) || (p.mode & 16));
          if (relay) mbar_wait_cluster(&peer_ready[stage], phase);
          waited += clock64() - w0;
          tc_fence_after();
          const uint32_t a_addr = smem_u32(tiles + stage * STAGE_BYTES);
          const uint64_t da = umma_desc_sw128(a_addr), db = umma_desc_sw128(a_addr + A_BYTES);
          if (elect_one()) {
            if (!no_mma) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma2_bf16_ss(tmem_base + (iter & 1) * BN, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma2_commit_both(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (cid == 0 && lane == 0 && iter < 4) { p.out[2 * iter] = clock64() - t0; p.out[2 * iter + 1] = waited; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) { tc_fence_after(); tmem_dealloc2<512>(tmem_base); }
}

static int encode(CUtensorMap* tm, void* ptr, long long cols, long long rows, int box_cols, int box_rows) {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return 1;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows}, es[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}

int main(int argc, char** argv) {
  const int M = 20224, N = argc > 1 ? atoi(argv[1]) : 768, K = argc > 2 ? atoi(argv[2]) : 3072;
  const int grid = argc > 3 ? atoi(argv[3]) : 148;
  void *A, *B; long long* out;
  cudaMalloc(&A, (size_t)M * K * 2); cudaMalloc(&B, (size_t)N * K * 2); cudaMalloc(&out, 64);
  cudaMemset(A, 0, (size_t)M * K * 2); cudaMemset(B, 0, (size_t)N * K * 2);
  CUtensorMap ta, tb;
  if (encode(&ta, A, K, M, BK, BM) || encode(&tb, B, K, N, BK, BN / 2)) { printf("encode failed\n"); return 1; }
  const int smem = STAGES * STAGE_BYTES + 1024 + 512;
  cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("M=%d N=%d K=%d grid=%d stages=%d (ideal 512 cycles per 256x256x64 k-block)\n", M, N, K, grid, STAGES);
  for (int mode : {0, 4, 20, 6, 22}) {
    P p{K, N / BN, (M / 256) * (N / BN), mode, out};
    cudaMemset(out, 0, 64);
    k2<<<grid, 128, smem>>>(ta, tb, p);
    cudaEventRecord(e0);
    k2<<<grid, 128, smem>>>(ta, tb, p);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[8]; cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
    const double fl = 2.0 * (M / 256 * 256) * N * K;
    printf("mode %2d: %8.1f us %7.1f TFLOP/s-equiv | cycles per k-block (waiting share): ", mode, ms * 1e3, fl / (ms * 1e-3) / 1e12);
    for (int i = 0; i < 4; ++i) printf("%lld (%.0f%%) ", h[2 * i] / (K / BK), h[2 * i] ? 100.0 * h[2 * i + 1] / h[2 * i] : 0.0);
    printf(" [%s]\n", cudaGetErrorString(err != cudaSuccess ? err : cudaGetLastError()));
    if (err != cudaSuccess) return 1;
  }
  return 0;
}
