// Variable-length (per-utterance) multi-head self-attention on tcgen05 / TMEM, head_dim = 64.
// Restates HF/modeling_wav2vec2.py:500-544 + transformers/integrations/sdpa_attention.py:40-104 for the batched
// engine: non-causal, scale head_dim^-0.5, no mask inside an utterance, and -- because many utterances share one
// packed token axis -- keys/queries of other utterances are never visible.
//
// Layout: qkv bf16 [M, 3H] (q | k | v, head h at columns h*64), O bf16 [M, H], LSE fp32 [heads, M] in base-2 units,
// block table int4 {utt_row0, T_u, block_start_in_utt, 0} with one entry per 128 query rows.
//
// Forward.  A kernel that allocates tensor memory runs ONE CTA per SM on this driver (the occupancy API reports 1 even
// with no shared memory), so the kernel is persistent and fills the SM by itself: work item = (128 queries, head),
// items are dealt round-robin to the CTAs and the next item's tiles are prefetched while the current one finishes.
//   warp 0      TMA producer: Q (double buffered across items), then the K_j / V_j tiles (64 keys x 64) through a
//               4-stage ring that runs straight across item boundaries
//   warp 1      score issuer:  S[g][b] = Q K_j^T  (M=128, N=64, K=64) into TMEM, g = j & 1, four key blocks ahead
//   warp 10     output issuer: O_g += P[g][b] V_j (M=128, N=64, K=64; A = P from shared memory, B = V_j MN-major)
//   warps 2..5  softmax group 0 (even key blocks), warps 6..9 softmax group 1 (odd key blocks): one query row per
//               thread (TMEM lane = row).  The groups ping-pong: while one exponentiates its block the tensor core
//               produces the other's scores, and every SM sub-partition always has two softmax warps to interleave.
//               Each group keeps its own running max / sum and its own accumulator O_g (lazy rescaling: O_g is only
//               rescaled when the row maximum grows by more than 2^8); the two partial results are merged at the end
//               of the item (split-KV merge through shared memory), normalised, and stored with LSE.
// The exp2 throughput of the SM (16/clk) is the bound: 128 x 128 scores take 1024 cycles against 512 cycles of MMA.
#include <stdlib.h>
#include <string.h>

#include "attention_tc.cuh"

namespace {

constexpr int TOKEN_EARLY = 11;      // exp2 iteration (of 16) after which a softmax warp releases the MUFU token


using namespace attn_tc;

constexpr int HD = 64;
constexpr int BQ = 128;                      // queries per item (= TMEM lanes)
constexpr int BKV = 64;                      // keys per block
constexpr int QTILE = BQ * HD * 2;           // 16 KB: 128 rows x 128 B
constexpr int KTILE = BKV * HD * 2;          // 8 KB
constexpr int NS = 6;                        // K/V ring depth (S runs four key blocks ahead of P V)
constexpr int FWD_THREADS = 352;            // TMA, score MMA, 8 softmax warps, output MMA

// shared-memory map of the forward kernel (offsets from a 1024-byte aligned base)
constexpr int F_Q = 0;                       // 2 buffers (item parity)
constexpr int F_KV = F_Q + 2 * QTILE;        // NS stages of {K_j, V_j}
constexpr int F_P = F_KV + NS * 2 * KTILE;   // P[g][b]: 4 x [128 rows][64 keys] bf16, K-major, 128B swizzle (b = 0 also merge buffers)
constexpr int F_X = F_P + 4 * QTILE;         // merge scalars: m[2][128], l[2][128] fp32
constexpr int F_BAR = F_X + 4 * 128 * 4;
constexpr int F_SMEM = F_BAR + 512;

__global__ void __launch_bounds__(FWD_THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, bf16* __restrict__ O,
                   float* __restrict__ LSE, const int4* __restrict__ tab, int n_blk, int heads, int H, long long M,
                   float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = n_blk * heads;
#ifdef ATTN_TIMING
  const long long tk0 = clock64();
  long long tk[24];
  int ntk = 0;
  int it = 0;
#define STAMP() do { if (ntk < 24 && it == 1) tk[ntk++] = clock64() - tk0; } while (0)
#else
#define STAMP() do {} while (0)
#endif

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F_BAR);
  uint64_t* q_full = bars;             // [2]  TMA: Q of item parity landed
  uint64_t* q_empty = bars + 2;        // [2]  MMA: all S = Q K^T of that item retired
  uint64_t* kv_full = bars + 4;        // [NS] TMA: K_j and V_j landed
  uint64_t* kv_empty = bars + 4 + NS;  // [NS] MMA: P_j V_j retired, stage reusable
  uint64_t* s_full = bars + 4 + 2 * NS;   // [g][b] MMA: S[g][b] complete in TMEM
  uint64_t* p_full = s_full + 4;       // [g][b] softmax group g (128 arrivals): P[g][b] in shared memory, S[g][b] consumed
  uint64_t* p_empty = p_full + 4;      // [g][b] MMA: P[g][b] V retired (O_g includes it; the buffer may be rewritten)
  uint64_t* o_full = p_empty + 4;      //      MMA: both accumulators of the item are final
  uint64_t* o_empty = o_full + 1;      //      softmax (256 arrivals): accumulators read, next item may overwrite them
  uint64_t* tok = o_empty + 1;         // [2][4] MUFU token of each sub-partition's warp pair: [0][q] "group 1 done", [1][q] "group 0 done"
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tok + 8);

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();      // the swizzled tiles need a 1024-byte aligned base
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); mbar_init(&p_empty[i], 1); }
    for (int i = 0; i < NS; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(o_full, 1);
    mbar_init(o_empty, 256);
    for (int i = 0; i < 8; ++i) mbar_init(&tok[i], 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;       // columns: S[g][b] at (2g + b) * 64, O_g at 256 + 64 g
  const uint32_t sQ = smem_u32(smem + F_Q), sKV = smem_u32(smem + F_KV), sP = smem_u32(smem + F_P);

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      uint32_t kvc = 0;                        // running K/V block count -> ring stage / phase
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int item = n_items - 1 - w;      // the table is sorted by length: longest utterances first
        const int4 t = __ldg(&tab[item / heads]);
        const int head = item - (item / heads) * heads;
        const int urow0 = t.x, T = t.y, m0 = t.z;
        const int nkv = (T + BKV - 1) / BKV;
        const int qb = it & 1;
        mbar_wait(&q_empty[qb], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[qb], QTILE);
        tma_load_2d(smem + F_Q + qb * QTILE, &tm_q, &q_full[qb], head * HD, urow0 + m0);
        for (int j = 0; j < nkv; ++j, ++kvc) {
          const int s = kvc % NS;
          mbar_wait(&kv_empty[s], ((kvc / NS) & 1) ^ 1);
          mbar_expect_tx(&kv_full[s], 2 * KTILE);
          tma_load_2d(smem + F_KV + s * 2 * KTILE, &tm_kv, &kv_full[s], H + head * HD, urow0 + j * BKV);
          tma_load_2d(smem + F_KV + s * 2 * KTILE + KTILE, &tm_kv, &kv_full[s], 2 * H + head * HD, urow0 + j * BKV);
        }
      }
    }
  } else if (warp == 1) {
    {
      // ===================== score issuer: S[g][b] = Q K_j^T, four key blocks ahead of the softmax =====================
      // (The whole warp runs the loop and one elected lane issues, see elect_one(); the two kinds of MMA have an issuing
      // warp each; descriptors advance by plain adds: +2 per 32 bytes in the address field.)
      constexpr uint32_t idesc_qk = idesc_bf16(128, BKV, false, false);
      uint32_t kvc = 0;                        // ring position of block 0 of the current item
      uint32_t sc0 = 0, sc1 = 0;               // S tiles issued for group 0 / 1 so far  (buffer = count & 1)
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int item = n_items - 1 - w;
        const int T = __ldg(&tab[item / heads]).y;
        const int nkv = (T + BKV - 1) / BKV;
        const int qb = it & 1;
        const uint64_t q_desc = umma_desc_sw128(sQ + qb * QTILE);
        mbar_wait(&q_full[qb], (it >> 1) & 1);
        for (int j = 0; j < nkv; ++j) {
          const uint32_t c = kvc + j;
          const int s = c % NS;
          const int g = j & 1;
          const uint32_t n = g ? sc1 : sc0;
          const int sb = g * 2 + (n & 1);
          if (g) ++sc1; else ++sc0;
          // S[g][b] is free once the softmax of this group's block two (local) blocks earlier has consumed it
          if (n >= 2) mbar_wait(&p_full[sb], ((n >> 1) & 1) ^ 1);
          mbar_wait(&kv_full[s], (c / NS) & 1);
          tc_fence_after();
          const uint64_t k_desc = umma_desc_sw128(sKV + s * 2 * KTILE);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) umma_bf16_ss(tmem_base + sb * BKV, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, k != 0);
            umma_commit(&s_full[sb]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&q_empty[qb]);
        __syncwarp();
        kvc += nkv;
      }
    }
  } else if (warp == 10) {
    {
      // ===================== output issuer: O_g += P[g][b] V_j =====================
      constexpr uint32_t idesc_pv = idesc_bf16(128, HD, false, true);
      uint32_t kvc = 0;
      uint32_t pc0 = 0, pc1 = 0;               // P tiles consumed from group 0 / 1 so far
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int item = n_items - 1 - w;
        const int T = __ldg(&tab[item / heads]).y;
        const int nkv = (T + BKV - 1) / BKV;
        for (int j = 0; j < nkv; ++j) {
          const int g = j & 1;
          const uint32_t c = kvc + j;
          const int s = c % NS;
          const int nvalid = min(BKV, T - j * BKV);
          const int ksteps = (nvalid + 15) >> 4;
          const uint32_t n = g ? pc1 : pc0;
          const int pb = g * 2 + (n & 1);
          if (g) ++pc1; else ++pc0;
          mbar_wait(&p_full[pb], (n >> 1) & 1);
          if (j < 2 && it > 0) mbar_wait(o_empty, (it - 1) & 1);       // previous item's accumulators have been read
          tc_fence_after();
          const uint64_t p_desc = umma_desc_sw128(sP + pb * QTILE);
          const uint64_t v_desc = umma_desc_sw128_mn(sKV + s * 2 * KTILE + KTILE);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              if (ks < ksteps) umma_bf16_ss(tmem_base + 256 + g * HD, p_desc + 2 * ks, v_desc + 128 * ks, idesc_pv, (j >= 2 || ks > 0) ? 1u : 0u);
            umma_commit(&kv_empty[s]);         // K_j was consumed by S_j long ago (its softmax has finished), V_j by these MMAs
            umma_commit(&p_empty[pb]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(o_full);
        __syncwarp();
        kvc += nkv;
      }
    }
  } else {
    // ===================== softmax groups: one query row per thread =====================
    const int g = (warp - 2) >> 2;                   // group 0: even key blocks, group 1: odd key blocks (warps 2..9)
    const int q = warp & 3;
    const int r = q * 32 + lane;                       // row inside the 128-query block = TMEM lane
    const uint32_t tO = tmem_base + ((uint32_t)(q * 32) << 16) + 256 + g * HD;
    float* xm = reinterpret_cast<float*>(smem + F_X);   // [2][128] running max of each group
    float* xl = xm + 256;                               // [2][128] running sum
    uint32_t cnt = 0;                                   // blocks this group has processed so far (s_full / p_empty phase)
    uint32_t rnd = 0;                                   // rounds so far (token phase)
    float sc = scale_log2;
    asm volatile("" : "+f"(sc));                        // keep the scale in a register (not re-read from the constant bank per element)
    int it = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int item = n_items - 1 - w;
      const int4 t = __ldg(&tab[item / heads]);
      const int head = item - (item / heads) * heads;
      const int urow0 = t.x, T = t.y, m0 = t.z;
      const int nkv = (T + BKV - 1) / BKV;
      float m_run = -INFINITY, l_run = 0.f;
      // Rounds: in round rr group g owns key block 2 rr + g.  The two warps that share an SM sub-partition (same q, one per
      // group) pass a token around the exp2 section so that they never compete for the MUFU (8 cycles per warp
      // instruction): while one exponentiates, the other loads scores, reduces the row maximum, packs and stores P.
      // Both groups run every round (an odd block count gives group 1 an empty last round) so the token keeps alternating.
      const int rounds = (nkv + 1) >> 1;
      for (int rr = 0; rr < rounds; ++rr, ++rnd) {
        const int j = 2 * rr + g;
        if (j >= nkv) {                                // empty round: pass the token on
          mbar_wait(&tok[4 + q], rnd & 1);
          __syncwarp();
          if (lane == 0) mbar_arrive(&tok[q]);
          continue;
        }
        const int nvalid = min(BKV, T - j * BKV);
        const int sb2 = g * 2 + (cnt & 1);               // S / P buffer of this block
        const uint32_t tS = tmem_base + ((uint32_t)(q * 32) << 16) + sb2 * BKV;
        const uint32_t prow = sP + sb2 * QTILE + r * 128;
        STAMP();
        mbar_wait(&s_full[sb2], (cnt >> 1) & 1);
        tc_fence_after();
        // One step of the online softmax over the block's 64 keys.  P = exp2(S * scale - m), row sum, bf16 A operand of
        // the second MMA (K-major, 128B swizzle).  Scores of masked keys (another utterance's rows) are exponentiated
        // too and then discarded by a select, never multiplied.
        uint32_t sa[32], sb[32];
        tmem_ld_32x32(tS, sa);
        tmem_ld_32x32(tS + 32, sb);
        tmem_ld_wait();
        STAMP();
        const bool full = nvalid == BKV;               // only an utterance's last block is partial
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], fmaxf(__uint_as_float(sa[i]), __uint_as_float(sb[i])));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            m4[i & 3] = fmaxf(m4[i & 3], fmaxf(i < nvalid ? __uint_as_float(sa[i]) : -INFINITY, 32 + i < nvalid ? __uint_as_float(sb[i]) : -INFINITY));
        }
        const float m_new = fmaxf(m_run, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * sc);
        if (j < 2) {
          m_run = m_new;                               // first block of this group: nothing accumulated yet
        } else {
          const bool need = m_new > m_run + 8.0f;      // lazy rescale: stale maxima up to 2^8 below are harmless in fp32/bf16
          if (__any_sync(0xffffffffu, need)) {
            const float alpha = need ? ex2_approx(m_run - m_new) : 1.0f;
            if (need) m_run = m_new;
            l_run *= alpha;
            // O_g holds this group's earlier blocks; wait for the last P V into it, then rescale in place
            mbar_wait(&p_empty[g * 2 + ((cnt - 1) & 1)], ((cnt - 1) >> 1) & 1);
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
        float nm = -m_run;
        STAMP();
        // ---- MUFU token: group 0 goes first in every round, then group 1 ----
        mbar_wait(&tok[(g ? 4 : 0) + q], g ? (rnd & 1) : ((rnd & 1) ^ 1));
        asm volatile("" : "+f"(nm));                   // nothing of the exp2 section may be scheduled above the wait
        float l4[4] = {0.f, 0.f, 0.f, 0.f};
        uint64_t lacc[2] = {0ull, 0ull};               // two packed pairs of partial row sums
        const uint64_t sc2 = dup2(sc), nm2 = dup2(nm);
        uint32_t pk[32];
        bool token_passed = false;
        if (full) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            // packed fp32 (FFMA2 / FADD2): the section shares its scheduler with the partner warp's score loads, row
            // maximum and P stores, and 64 scalar FFMAs + 63 FADDs on top of the 64 MUFUs made it issue-bound
            float x0, x1, x2, x3;
            upk2(fma2(pk2(__uint_as_float(sa[2 * i]), __uint_as_float(sa[2 * i + 1])), sc2, nm2), x0, x1);
            upk2(fma2(pk2(__uint_as_float(sb[2 * i]), __uint_as_float(sb[2 * i + 1])), sc2, nm2), x2, x3);
            const float e0 = ex2_approx(x0), e1 = ex2_approx(x1), e2 = ex2_approx(x2), e3 = ex2_approx(x3);
            lacc[i & 1] = add2(lacc[i & 1], add2(pk2(e0, e1), pk2(e2, e3)));
            pk[i] = pack_bf16x2(e0, e1);
            pk[16 + i] = pack_bf16x2(e2, e3);
            if (i == TOKEN_EARLY) {                      // hand the MUFU token on early: the partner needs ~200 cycles to wake
              asm volatile("" : "+l"(lacc[i & 1]));
              __syncwarp();
              if (lane == 0) mbar_arrive(&tok[(g ? 0 : 4) + q]);
            }
          }
          token_passed = true;
          float a0, a1, a2, a3;
          upk2(lacc[0], a0, a1);
          upk2(lacc[1], a2, a3);
          l4[0] = a0; l4[1] = a1; l4[2] = a2; l4[3] = a3;
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float e0 = ex2_approx(fmaf(__uint_as_float(sa[2 * i]), sc, nm)), e1 = ex2_approx(fmaf(__uint_as_float(sa[2 * i + 1]), sc, nm));
            float e2 = ex2_approx(fmaf(__uint_as_float(sb[2 * i]), sc, nm)), e3 = ex2_approx(fmaf(__uint_as_float(sb[2 * i + 1]), sc, nm));
            e0 = 2 * i < nvalid ? e0 : 0.f;
            e1 = 2 * i + 1 < nvalid ? e1 : 0.f;
            e2 = 32 + 2 * i < nvalid ? e2 : 0.f;
            e3 = 33 + 2 * i < nvalid ? e3 : 0.f;
            l4[i & 3] += (e0 + e1) + (e2 + e3);
            pk[i] = pack_bf16x2(e0, e1);
            pk[16 + i] = pack_bf16x2(e2, e3);
          }
        }
        float lsum = (l4[0] + l4[1]) + (l4[2] + l4[3]);
        asm volatile("" : "+f"(lsum));                 // every exp2 has been issued and consumed before the token moves on
        __syncwarp();
        if (!token_passed && lane == 0) mbar_arrive(&tok[(g ? 0 : 4) + q]);
        STAMP();
        mbar_wait(&p_empty[sb2], ((cnt >> 1) & 1) ^ 1);  // the P V that read this buffer two blocks ago has retired
#pragma unroll
        for (int i = 0; i < 8; ++i)
          st_shared_v4(prow + ((i ^ (r & 7)) << 4), pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        l_run += lsum;
        STAMP();
        fence_proxy_async_smem();
        STAMP();
        tc_fence_before();
        mbar_arrive(&p_full[sb2]);
        STAMP();
        ++cnt;
      }
      STAMP();
      // ---- end of item: merge the two groups' partial results (split-KV merge), normalise, store ----
      xm[g * 128 + r] = m_run;
      xl[g * 128 + r] = l_run;
      mbar_wait(o_full, it & 1);
      tc_fence_after();
      named_bar_sync(1, 256);
      const float m_all = fmaxf(xm[r], xm[128 + r]);
      const float w0 = ex2_approx(xm[r] - m_all), w1 = ex2_approx(xm[128 + r] - m_all);   // an idle group has m = -inf -> 0
      const float l_all = xl[r] * w0 + xl[128 + r] * w1;
      const float wg = (g ? w1 : w0) / l_all;
      const bool has = g < nkv;                        // this group accumulated at least one block
      // my accumulator, scaled; the half I do not store goes to the peer group through shared memory
      uint32_t keep[32], give[32];
      if (has) {
        tmem_ld_32x32(tO + g * 32, keep);              // group 0 stores columns 0..31, group 1 columns 32..63
        tmem_ld_32x32(tO + (g ^ 1) * 32, give);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(o_empty);
      const uint32_t xrow = sP + g * 2 * QTILE + r * 128;   // the P buffers are idle now: [128 rows][32 fp32], 16-byte chunks swizzled
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has) v = make_float4(__uint_as_float(give[4 * i]) * wg, __uint_as_float(give[4 * i + 1]) * wg,
                                 __uint_as_float(give[4 * i + 2]) * wg, __uint_as_float(give[4 * i + 3]) * wg);
        st_shared_v4(xrow + ((i ^ (r & 7)) << 4), __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
      }
      named_bar_sync(1, 256);
      const bool ok = m0 + r < T;
      const long long row = (long long)urow0 + m0 + r;
      const uint32_t yrow = sP + (g ^ 1) * 2 * QTILE + r * 128;
      float o[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 v = ld_shared_v4(yrow + ((i ^ (r & 7)) << 4));
        o[4 * i] = v.x; o[4 * i + 1] = v.y; o[4 * i + 2] = v.z; o[4 * i + 3] = v.w;
      }
      if (has) {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = fmaf(__uint_as_float(keep[i]), wg, o[i]);
      }
      if (ok) {
        uint4* op = reinterpret_cast<uint4*>(O + row * H + head * HD + g * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          op[i] = make_uint4(pack_bf16x2(o[8 * i], o[8 * i + 1]), pack_bf16x2(o[8 * i + 2], o[8 * i + 3]),
                             pack_bf16x2(o[8 * i + 4], o[8 * i + 5]), pack_bf16x2(o[8 * i + 6], o[8 * i + 7]));
        if (g == 0) LSE[(long long)head * M + row] = m_all + log2f(l_all);
      }
      named_bar_sync(1, 256);                          // the merge buffers become P[g] again
      STAMP();
#ifdef ATTN_TIMING
      if (it == 1 && (threadIdx.x == 64 || threadIdx.x == 192) && blockIdx.x == 3) {
        printf("attn fwd thread %d T=%d nkv=%d stamps:", threadIdx.x, T, nkv);
        for (int i = 0; i < ntk; ++i) printf(" %lld", tk[i]);
        printf("\n");
      }
#endif
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

int attention_forward(const bf16* qkv, bf16* O, float* LSE, const int4* blk_tab, int n_blk, int H, int heads, long long M,
                         cudaStream_t stream) {
  SUTA_CHECK_ARG(H == heads * HD);
  if (n_blk <= 0) return SUTA_OK;
  static const bool shared_items = getenv("SUTA_ATTN_FWD_V1") != nullptr;    // debug switch: this file's kernel
  if (!shared_items) return attention_forward_v2(qkv, O, LSE, blk_tab, n_blk, H, heads, M, stream);
  static bool attr = false;
  static int n_sm = 148;
  if (!attr) {
    CUDA_TRY(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    attr = true;
  }
  CUtensorMap tmq, tmkv;
  SUTA_TRY(make_map(&tmq, qkv, M, 3LL * H, 3LL * H, BQ));
  SUTA_TRY(make_map(&tmkv, qkv, M, 3LL * H, 3LL * H, BKV));
  const float scale_log2 = 0.125f * 1.4426950408889634f;   // 64^-0.5 * log2(e)
  const long long n_items = (long long)n_blk * heads;
  const int grid = (int)(n_items < n_sm ? n_items : n_sm);
  attn_fwd_tc_kernel<<<grid, FWD_THREADS, F_SMEM, stream>>>(tmq, tmkv, O, LSE, blk_tab, n_blk, heads, H, M, scale_log2);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}
