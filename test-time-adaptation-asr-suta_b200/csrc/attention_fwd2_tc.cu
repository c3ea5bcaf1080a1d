// Variable-length multi-head self-attention forward on tcgen05 / TMEM, head_dim = 64: TWO INDEPENDENT PIPELINES per CTA.
// Contract (HF/modeling_wav2vec2.py:500-544, non-causal, scale 64^-0.5, utterances never see each other): qkv bf16 [M, 3H], O bf16 [M, H], LSE fp32 [heads, M] in base-2 units, block table int4 {utt_row0, T_u,
// block_start_in_utt, 0} per 128 query rows.
//
// The first version (removed in round 2) let two softmax groups share one work item (even / odd key blocks, split-KV merge at
// the end): with the short utterances of this workload (median 5 key blocks per item) the per-item drain -- last P V, three CTA-wide barriers
// for the merge, normalise, store -- cost ~40 % of the item, and both groups paid it at the same time.  Here every softmax
// group owns a complete pipeline and its own item stream (a kernel that allocates TMEM gets one CTA per SM, so the two
// "virtual CTAs" live in one): its own TMA warp, MMA warp, Q double buffer, K/V ring, S / P double buffers and accumulator.
// While one group drains an item the other keeps the tensor core and the MUFU busy; nothing is merged, no CTA-wide barrier
// exists after start-up.  Consecutive items (heads h, h+1 of one query tile) go to the two groups, so they stay balanced.
//   warps 0 / 10   TMA producer of group 0 / 1
//   warps 1 / 11   MMA issuer of group 0 / 1: S_{j+1} = Q K_{j+1}^T is issued before O += P_j V_j (software pipelining)
//   warps 2..5 / 6..9   softmax group 0 / 1: one query row per thread (TMEM lane = row), online softmax with lazy rescaling
#include <stdlib.h>
#include <string.h>

#include "attention_tc.cuh"

namespace {

using namespace attn_tc;

constexpr int HD = 64;
constexpr int BQ = 128;                      // queries per item (= TMEM lanes)
constexpr int BKV = 64;                      // keys per block
constexpr int QTILE = BQ * HD * 2;           // 16 KB: 128 rows x 128 B
constexpr int KTILE = BKV * HD * 2;          // 8 KB
constexpr int NSG = 3;                       // K/V ring depth per group
constexpr int THREADS = 384;

// shared-memory map (offsets from a 1024-byte aligned base); per group: Q[2] | {K,V}[NSG] | P[2]
constexpr int G_Q = 0;
constexpr int G_KV = G_Q + 2 * QTILE;
constexpr int G_P = G_KV + NSG * 2 * KTILE;
constexpr int G_BYTES = G_P + 2 * QTILE;     // 112 KB
constexpr int F_BAR = 2 * G_BYTES;
constexpr int F_SMEM = F_BAR + 512;
constexpr int NBAR = 2 + 2 + NSG + NSG + 2 + 2 + 2 + 1 + 1;   // barriers per group

__global__ void __launch_bounds__(THREADS, 1)
attn_fwd2_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                    const __grid_constant__ CUtensorMap tm_o, bf16* __restrict__ O,
                    float* __restrict__ LSE, const int4* __restrict__ tab, int n_blk, int heads, int H, long long M,
                    float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = n_blk * heads;
  // role and group of this warp
  const int g = (warp == 0 || warp == 1 || (warp >= 2 && warp <= 5)) ? 0 : 1;
  const bool is_tma = warp == 0 || warp == 10, is_mma = warp == 1 || warp == 11;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F_BAR) + g * NBAR;
  uint64_t* q_full = bars;                 // [2]   TMA: Q of item parity landed
  uint64_t* q_empty = bars + 2;            // [2]   MMA: all S = Q K^T of that item retired
  uint64_t* kv_full = bars + 4;            // [NSG] TMA: K_j and V_j landed
  uint64_t* kv_empty = kv_full + NSG;      // [NSG] MMA: P_j V_j retired, stage reusable
  uint64_t* s_full = kv_empty + NSG;       // [2]   MMA: S[b] complete in TMEM
  uint64_t* p_full = s_full + 2;           // [2]   softmax (128 arrivals): P[b] in shared memory, S[b] consumed
  uint64_t* p_empty = p_full + 2;          // [2]   MMA: P[b] V retired (O includes it; the buffer may be rewritten)
  uint64_t* o_full = p_empty + 2;          //       MMA: the item's accumulator is final
  uint64_t* o_empty = o_full + 1;          //       softmax (128 arrivals): accumulator read, the next item may overwrite it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + F_BAR + 2 * NBAR * 8);

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();      // the swizzled tiles need a 1024-byte aligned base
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_o);
    for (int gg = 0; gg < 2; ++gg) {
      uint64_t* b = reinterpret_cast<uint64_t*>(smem + F_BAR) + gg * NBAR;
      for (int i = 0; i < 4; ++i) mbar_init(&b[i], 1);                        // q_full, q_empty
      for (int i = 0; i < 2 * NSG; ++i) mbar_init(&b[4 + i], 1);              // kv_full, kv_empty
      uint64_t* sf = b + 4 + 2 * NSG;
      for (int i = 0; i < 2; ++i) { mbar_init(&sf[i], 1); mbar_init(&sf[2 + i], 128); mbar_init(&sf[4 + i], 1); }
      mbar_init(&sf[6], 1);                                                   // o_full
      mbar_init(&sf[7], 128);                                                 // o_empty
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_g = *tmem_slot + g * 256;    // this group's columns: S[b] at b * 64, O at 128
  uint8_t* sm_g = smem + g * G_BYTES;
  const uint32_t sQ = smem_u32(sm_g + G_Q), sKV = smem_u32(sm_g + G_KV), sP = smem_u32(sm_g + G_P);
  const int w0 = 2 * blockIdx.x + g, wstep = 2 * gridDim.x;   // this group's item stream (longest items first)

  if (is_tma) {
    if (lane == 0) {
      // ===================== TMA producer of group g =====================
      uint32_t kvc = 0;                        // running K/V block count -> ring stage / phase
      int it = 0;
      for (int w = w0; w < n_items; w += wstep, ++it) {
        const int item = n_items - 1 - w;      // the table is sorted by length: longest utterances first
        const int4 t = __ldg(&tab[item / heads]);
        const int head = item - (item / heads) * heads;
        const int urow0 = t.x, T = t.y, m0 = t.z;
        const int nkv = (T + BKV - 1) / BKV;
        const int qb = it & 1;
        mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[qb], QTILE);
        tma_load_2d(sm_g + G_Q + qb * QTILE, &tm_q, &q_full[qb], head * HD, urow0 + m0);
        for (int j = 0; j < nkv; ++j, ++kvc) {
          const int s = kvc % NSG;
          mbar_wait(&kv_empty[s], ((kvc / NSG) & 1) ^ 1);
          mbar_expect_tx(&kv_full[s], 2 * KTILE);
          tma_load_2d(sm_g + G_KV + s * 2 * KTILE, &tm_kv, &kv_full[s], H + head * HD, urow0 + j * BKV);
          tma_load_2d(sm_g + G_KV + s * 2 * KTILE + KTILE, &tm_kv, &kv_full[s], 2 * H + head * HD, urow0 + j * BKV);
        }
      }
    }
  } else if (is_mma) {
    // ===================== MMA issuer of group g (converged warp, one elected lane issues) =====================
    constexpr uint32_t idesc_qk = idesc_bf16(128, BKV, false, false);
    constexpr uint32_t idesc_pv = idesc_bf16(128, HD, false, true);
    uint32_t kvc = 0;                          // ring position of block 0 of the current item
    uint32_t sc = 0, pc = 0;                   // S tiles issued / P tiles consumed so far (buffer = count & 1)
    int it = 0;
    for (int w = w0; w < n_items; w += wstep, ++it) {
      const int item = n_items - 1 - w;
      const int T = __ldg(&tab[item / heads]).y;
      const int nkv = (T + BKV - 1) / BKV;
      const int qb = it & 1;
      const uint64_t q_desc = umma_desc_sw128(sQ + qb * QTILE);
      mbar_wait(&q_full[qb], (it >> 1) & 1);
      // S[b] = Q K_j^T.  The buffer was last used two S tiles ago; the P V of that tile has been issued by this warp
      // already, and that waited for p_full, which the softmax posts after reading S -- so the buffer is free.
      auto issue_s = [&](int j) {
        const uint32_t c = kvc + j;
        const int s = c % NSG, sb = sc & 1;
        ++sc;
        mbar_wait(&kv_full[s], (c / NSG) & 1);
        tc_fence_after();
        const uint64_t k_desc = umma_desc_sw128(sKV + s * 2 * KTILE);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) umma_bf16_ss(tmem_g + sb * BKV, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, k != 0);
          umma_commit(&s_full[sb]);
        }
        __syncwarp();
      };
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        if (j + 1 < nkv) issue_s(j + 1);       // the next block's scores are produced while this block's softmax runs
        const uint32_t c = kvc + j;
        const int s = c % NSG, pb = pc & 1;
        const int nvalid = min(BKV, T - j * BKV);
        const int ksteps = (nvalid + 15) >> 4;
        mbar_wait(&p_full[pb], (pc >> 1) & 1);
        ++pc;
        if (j == 0 && it > 0) mbar_wait(o_empty, (it - 1) & 1);     // the previous item's accumulator has been read
        tc_fence_after();
        const uint64_t p_desc = umma_desc_sw128(sP + pb * QTILE);
        const uint64_t v_desc = umma_desc_sw128_mn(sKV + s * 2 * KTILE + KTILE);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            if (ks < ksteps) umma_bf16_ss(tmem_g + 128, p_desc + 2 * ks, v_desc + 128 * ks, idesc_pv, (j > 0 || ks > 0) ? 1u : 0u);
          umma_commit(&kv_empty[s]);           // K_j was consumed by S_j long ago, V_j by these MMAs
          umma_commit(&p_empty[pb]);
        }
        __syncwarp();
      }
      if (elect_one()) {
        umma_commit(o_full);
        umma_commit(&q_empty[qb]);
      }
      __syncwarp();
      kvc += nkv;
    }
  } else {
    // ===================== softmax group g: one query row per thread =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;                       // row inside the 128-query block = TMEM lane
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tO = tmem_g + lane_off + 128;
    uint32_t cnt = 0;                                  // blocks processed so far (s_full / p_empty phase)
    float sc = scale_log2;
    asm volatile("" : "+f"(sc));                       // keep the scale in a register
    int it = 0;
#ifdef ATTN_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};      // s_full wait | S load | max | exp | p_empty wait + P store | fence + arrive | o_full wait | epilogue
    long long tprev = clock64();
    const long long tstart = tprev;
    int nblocks = 0;
#define LAP(k) do { const long long tn = clock64(); tacc[k] += tn - tprev; tprev = tn; } while (0)
#else
#define LAP(k) do {} while (0)
#endif
    int4 t_next = w0 < n_items ? __ldg(&tab[(n_items - 1 - w0) / heads]) : make_int4(0, 0, 0, 0);
    for (int w = w0; w < n_items; w += wstep, ++it) {
      const int item = n_items - 1 - w;
      const int4 t = t_next;                           // fetched one item ahead: a dependent global load per item start is
      if (w + wstep < n_items) t_next = __ldg(&tab[(n_items - 1 - w - wstep) / heads]);   // ~700 exposed cycles otherwise
      const int head = item - (item / heads) * heads;
      const int urow0 = t.x, T = t.y, m0 = t.z;
      const int nkv = (T + BKV - 1) / BKV;
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < nkv; ++j, ++cnt) {
        const int nvalid = min(BKV, T - j * BKV);
        const int sb = cnt & 1;
        const uint32_t tS = tmem_g + lane_off + sb * BKV;
        const uint32_t prow = sP + sb * QTILE + r * 128;
        LAP(7);
        mbar_wait(&s_full[sb], (cnt >> 1) & 1);
        tc_fence_after();
        LAP(0);
        // One step of the online softmax over the block's 64 keys.  P = exp2(S * scale - m), row sum, bf16 A operand of
        // the second MMA (K-major, 128B swizzle).  Scores of masked keys (another utterance's rows) are exponentiated
        // too and then discarded by a select, never multiplied.
        uint32_t sa[32], sb2[32];
        tmem_ld_32x32(tS, sa);
        tmem_ld_32x32(tS + 32, sb2);
        tmem_ld_wait();
        LAP(1);
        const bool full = nvalid == BKV;               // only an utterance's last block is partial
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], fmaxf(__uint_as_float(sa[i]), __uint_as_float(sb2[i])));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            m4[i & 3] = fmaxf(m4[i & 3], fmaxf(i < nvalid ? __uint_as_float(sa[i]) : -INFINITY, 32 + i < nvalid ? __uint_as_float(sb2[i]) : -INFINITY));
        }
        const float m_new = fmaxf(m_run, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * sc);
        if (j == 0) {
          m_run = m_new;                               // first block: nothing accumulated yet
        } else {
          const bool need = m_new > m_run + 8.0f;      // lazy rescale: stale maxima up to 2^8 below are harmless in fp32/bf16
          if (__any_sync(0xffffffffu, need)) {
            const float alpha = need ? ex2_approx(m_run - m_new) : 1.0f;
            if (need) m_run = m_new;
            l_run *= alpha;
            // O holds the earlier blocks; wait for the last P V into it, then rescale in place
            mbar_wait(&p_empty[(cnt - 1) & 1], ((cnt - 1) >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t o[16];
              tmem_ld_32x16(tO + c * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_32x16(tO + c * 16, o);
            }
            tmem_st_wait();
          }
        }
        const float nm = -m_run;
        LAP(2);
        float lsum;
        uint32_t pk[32];
        if (full) {
          uint64_t lacc[2] = {0ull, 0ull};             // packed fp32 pairs (FFMA2 / FADD2)
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float x0, x1, x2, x3;
            upk2(fma2(pk2(__uint_as_float(sa[2 * i]), __uint_as_float(sa[2 * i + 1])), dup2(sc), dup2(nm)), x0, x1);
            upk2(fma2(pk2(__uint_as_float(sb2[2 * i]), __uint_as_float(sb2[2 * i + 1])), dup2(sc), dup2(nm)), x2, x3);
            const float e0 = ex2_approx(x0), e1 = ex2_approx(x1), e2 = ex2_approx(x2), e3 = ex2_approx(x3);
            lacc[i & 1] = add2(lacc[i & 1], add2(pk2(e0, e1), pk2(e2, e3)));
            pk[i] = pack_bf16x2(e0, e1);
            pk[16 + i] = pack_bf16x2(e2, e3);
          }
          float a0, a1, a2, a3;
          upk2(lacc[0], a0, a1);
          upk2(lacc[1], a2, a3);
          lsum = (a0 + a1) + (a2 + a3);
        } else {
          float l4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float e0 = ex2_approx(fmaf(__uint_as_float(sa[2 * i]), sc, nm)), e1 = ex2_approx(fmaf(__uint_as_float(sa[2 * i + 1]), sc, nm));
            float e2 = ex2_approx(fmaf(__uint_as_float(sb2[2 * i]), sc, nm)), e3 = ex2_approx(fmaf(__uint_as_float(sb2[2 * i + 1]), sc, nm));
            e0 = 2 * i < nvalid ? e0 : 0.f;
            e1 = 2 * i + 1 < nvalid ? e1 : 0.f;
            e2 = 32 + 2 * i < nvalid ? e2 : 0.f;
            e3 = 33 + 2 * i < nvalid ? e3 : 0.f;
            l4[i & 3] += (e0 + e1) + (e2 + e3);
            pk[i] = pack_bf16x2(e0, e1);
            pk[16 + i] = pack_bf16x2(e2, e3);
          }
          lsum = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        }
        asm volatile("" : "+f"(lsum));
        LAP(3);
        mbar_wait(&p_empty[sb], ((cnt >> 1) & 1) ^ 1);   // the P V that read this buffer two blocks ago has retired
        if (j < 2) {                                   // ... and so has the previous item's output store staged in it
          if (lane == 0) bulk_wait_read<0>();
          __syncwarp();
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
          st_shared_v4(prow + ((i ^ (r & 7)) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        l_run += lsum;
        LAP(4);
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&p_full[sb]);
        LAP(5);
#ifdef ATTN_TIMING
        ++nblocks;
#endif
      }
      // ---- end of item: normalise this group's accumulator and store (nothing to merge, nobody else to wait for) ----
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      LAP(6);
      uint32_t o0[32], o1[32];
      tmem_ld_32x32(tO, o0);
      tmem_ld_32x32(tO + 32, o1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(o_empty);
      const bool row_ok = m0 + r < T;
      const long long row = (long long)urow0 + m0 + r;
      const float inv = 1.0f / l_run;
      uint4 ov[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        ov[i] = make_uint4(pack_bf16x2(__uint_as_float(o0[8 * i]) * inv, __uint_as_float(o0[8 * i + 1]) * inv),
                           pack_bf16x2(__uint_as_float(o0[8 * i + 2]) * inv, __uint_as_float(o0[8 * i + 3]) * inv),
                           pack_bf16x2(__uint_as_float(o0[8 * i + 4]) * inv, __uint_as_float(o0[8 * i + 5]) * inv),
                           pack_bf16x2(__uint_as_float(o0[8 * i + 6]) * inv, __uint_as_float(o0[8 * i + 7]) * inv));
        ov[4 + i] = make_uint4(pack_bf16x2(__uint_as_float(o1[8 * i]) * inv, __uint_as_float(o1[8 * i + 1]) * inv),
                               pack_bf16x2(__uint_as_float(o1[8 * i + 2]) * inv, __uint_as_float(o1[8 * i + 3]) * inv),
                               pack_bf16x2(__uint_as_float(o1[8 * i + 4]) * inv, __uint_as_float(o1[8 * i + 5]) * inv),
                               pack_bf16x2(__uint_as_float(o1[8 * i + 6]) * inv, __uint_as_float(o1[8 * i + 7]) * inv));
      }
      // A thread's 128 output bytes lie 2 H bytes from its neighbour's: 8 scattered 16-byte stores per thread cost ~2 k
      // cycles per item.  When the warp's 32 rows all exist they go through the warp's own 32 rows of the P buffer that is
      // written LAST (free: every P V of the item has retired) and out with one TMA store.
      if (m0 + q * 32 + 32 <= T) {
        const uint32_t stage = sP + ((cnt & 1) ^ 1) * QTILE + r * 128;
#pragma unroll
        for (int i = 0; i < 8; ++i) st_shared_v4(stage + ((i ^ (r & 7)) << 4), ov[i].x, ov[i].y, ov[i].z, ov[i].w);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tm_o, sP + ((cnt & 1) ^ 1) * QTILE + q * 32 * 128, head * HD, (int)(urow0 + m0 + q * 32));
          bulk_commit();
        }
      } else if (row_ok) {
        uint4* op = reinterpret_cast<uint4*>(O + row * H + head * HD);
#pragma unroll
        for (int i = 0; i < 8; ++i) op[i] = ov[i];
      }
      if (row_ok) LSE[(long long)head * M + row] = m_run + log2f(l_run);
    }
    if (lane == 0) bulk_wait<0>();                     // shared memory must outlive the last store's reads
#ifdef ATTN_TIMING
    LAP(7);
    if ((threadIdx.x == 64 || threadIdx.x == 192) && (blockIdx.x == 3 || blockIdx.x == 100))
      printf("attn fwd2 cta %d thread %d: items %d blocks %d total %lld | s_wait %lld ld %lld max %lld exp %lld pstore %lld fence %lld o_wait %lld epi %lld\n",
             blockIdx.x, threadIdx.x, it, nblocks, clock64() - tstart, tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5], tacc[6], tacc[7]);
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(*tmem_slot);
  }
}

}  // namespace

int attention_forward(const bf16* qkv, bf16* O, float* LSE, const int4* blk_tab, int n_blk, int H, int heads, long long M,
                         cudaStream_t stream) {
  SUTA_CHECK_ARG(H == heads * HD);
  if (n_blk <= 0) return SUTA_OK;
  static bool attr = false;
  static int n_sm = 148;
  if (!attr) {
    CUDA_TRY(cudaFuncSetAttribute(attn_fwd2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    attr = true;
  }
  CUtensorMap tmq, tmkv, tmo;
  SUTA_TRY(make_map(&tmq, qkv, M, 3LL * H, 3LL * H, BQ));
  SUTA_TRY(make_map(&tmkv, qkv, M, 3LL * H, 3LL * H, BKV));
  SUTA_TRY(make_map(&tmo, O, M, H, H, 32));                  // output: 32-row boxes, one per softmax warp
  const float scale_log2 = 0.125f * 1.4426950408889634f;   // 64^-0.5 * log2(e)
  const long long n_items = (long long)n_blk * heads;
  const long long want = (n_items + 1) / 2;                 // two item streams per CTA
  const int grid = (int)(want < n_sm ? want : n_sm);
  attn_fwd2_tc_kernel<<<grid, THREADS, F_SMEM, stream>>>(tmq, tmkv, tmo, O, LSE, blk_tab, n_blk, heads, H, M, scale_log2);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}
