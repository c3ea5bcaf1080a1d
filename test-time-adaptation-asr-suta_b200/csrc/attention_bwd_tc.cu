// Attention backward on tcgen05 / TMEM (head_dim 64, per-utterance, non-causal).  Layouts as in attention_fwd2_tc.cu.
//
//   S = Q K^T,  P = exp2(S * scale * log2e - LSE),  dP = dO V^T,  dS = P o (dP - D),  D = rowsum(dO o O)
//   dV = P^T dO,   dK = scale * dS^T Q,   dQ = scale * dS K
//
// One templated persistent kernel (one CTA per SM: tensor-memory kernels do not share an SM) serves both halves:
//   DKV = true   item = (128 keys, head).  The 128 rows of X1 = K_j, X2 = V_j stay in shared memory; the queries stream
//                by in blocks of 64 (Y1 = Q_i, Y2 = dO_i).  Score MMAs give the TRANSPOSED tiles S^T = K Q^T and
//                dP^T = V dO^T (M = 128 keys on the TMEM lanes, N = 64 queries), so each thread owns one key row and
//                the accumulators dV += P^T dO_i, dK += dS^T Q_i need no cross-CTA reduction.  LSE and D belong to the
//                columns here: staged in shared memory, +inf / 0 for queries past the utterance end (their P is 0).
//   DKV = false  item = (128 queries, head): X1 = Q_i, X2 = dO_i resident, keys stream by (Y1 = K_j, Y2 = V_j);
//                S = Q K^T, dP = dO V^T, one thread per query row (LSE, D in registers), dQ += dS K_j.
// Warp roles as in the forward: warp 0 TMA producer (X double buffered across items, Y through a 4-stage ring that
// runs across item boundaries), warp 1 score-MMA issuer, warps 10 / 11 output-MMA issuers, warps 2..5 / 6..9 two groups that take the even / odd streamed
// blocks (each with its own score tiles in TMEM and its own A-operand buffers in shared memory) and ping-pong on the
// MUFU.  Unlike the forward there is no running maximum: LSE is known, both groups add into the SAME accumulators.
// The streamed tiles are used twice with two descriptors: K-major as the B operand of the score MMAs and MN-major
// as the B operand of the output MMAs.
#include "attention_tc.cuh"

namespace {

using namespace attn_tc;

constexpr int HD = 64;
constexpr int BX = 128;                      // resident rows per item (= TMEM lanes)
constexpr int BY = 64;                       // streamed rows per block
constexpr int XTILE = BX * HD * 2;           // 16 KB
constexpr int YTILE = BY * HD * 2;           // 8 KB
constexpr int NS = 4;                        // Y ring depth
constexpr int BWD_THREADS = 384;          // TMA, score MMA, 8 elementwise warps, 2 output-MMA warps

// shared-memory map (offsets from a 1024-byte aligned base)
constexpr int B_X = 0;                       // [2 item buffers][X1 | X2]
constexpr int B_Y = B_X + 4 * XTILE;         // [NS][Y1 | Y2]
constexpr int B_A = B_Y + NS * 2 * YTILE;    // [2 groups][dS | P] bf16 [128 rows][64], K-major, 128B swizzle
constexpr int B_C = B_A + 4 * XTILE;         // [2 groups][2 buffers][lse[64] | D[64]] fp32 (DKV only)
constexpr int B_S = B_C + 4 * 512;           // [2 groups] result staging [128 rows][128 B] for the TMA stores
constexpr int B_BAR = B_S + 2 * XTILE;
constexpr int B_SMEM = B_BAR + 256;

template <bool DKV>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv_x, const __grid_constant__ CUtensorMap tm_qkv_y,
                   const __grid_constant__ CUtensorMap tm_do_x, const __grid_constant__ CUtensorMap tm_do_y,
                   const __grid_constant__ CUtensorMap tm_out, const float* __restrict__ LSE, const float* __restrict__ Dv, bf16* __restrict__ dqkv,
                   const int4* __restrict__ tab, int n_blk, int heads, int H, long long M, float scale, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = n_blk * heads;

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_BAR);
  uint64_t* x_full = bars;               // [2]  TMA: resident tiles of item parity landed
  uint64_t* x_empty = bars + 2;          // [2]  MMA: every MMA of that item that reads X has retired
  uint64_t* y_full = bars + 4;           // [NS] TMA: streamed tiles landed
  uint64_t* y_empty = bars + 4 + NS;     // [NS] MMA: output MMAs of that block retired
  uint64_t* sc_full = bars + 4 + 2 * NS; // [3]  MMA: both score tiles of set (block % 3) complete
  uint64_t* sc_free = sc_full + 3;       // [3]  consuming group (128 arrivals): the set's scores are in registers
  uint64_t* a_full = sc_free + 3;        // [2]  group g (128 arrivals): A operands written
  uint64_t* a_empty = a_full + 2;        // [g][b] MMA: output MMAs that read group g's A operands (buffer b) retired
  uint64_t* acc_full = a_empty + 4;      //      MMA: accumulators of the item are final
  uint64_t* acc_empty = acc_full + 1;    //      groups (256 arrivals): accumulators read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    tma_prefetch_desc(&tm_qkv_x);
    tma_prefetch_desc(&tm_qkv_y);
    tma_prefetch_desc(&tm_do_x);
    tma_prefetch_desc(&tm_do_y);
    tma_prefetch_desc(&tm_out);
    constexpr int NOUT = DKV ? 2 : 1;          // output issuers: each commits its own arrival
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 1); mbar_init(&a_full[i], 128);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&a_empty[i], NOUT);
    for (int i = 0; i < 3; ++i) { mbar_init(&sc_full[i], 1); mbar_init(&sc_free[i], 128); }
    for (int i = 0; i < NS; ++i) { mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], NOUT); }
    mbar_init(acc_full, NOUT);
    mbar_init(acc_empty, 256);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // TMEM columns: three sets of score tiles (Sc1 | Sc2, 128 columns) at 0, 128 and 384, used by the streamed blocks in
  // rotation (block c -> set c % 3) so that the scores of a group's NEXT block are produced while it works on the current
  // one; accumulators at 256 (dV | dQ) and 320 (dK)
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sX = smem_u32(smem + B_X), sY = smem_u32(smem + B_Y), sA = smem_u32(smem + B_A), sS = smem_u32(smem + B_S);
  // column offsets of the four operand tiles inside qkv / dO
  const int colQ = 0, colK = H, colV = 2 * H;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      uint32_t yc = 0;
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int item = n_items - 1 - w;
        const int4 t = __ldg(&tab[item / heads]);
        const int head = item - (item / heads) * heads;
        const int urow0 = t.x, T = t.y, m0 = t.z;
        const int ny = (T + BY - 1) / BY;
        const int xb = it & 1;
        mbar_wait(&x_empty[xb], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&x_full[xb], 2 * XTILE);
        uint8_t* xs = smem + B_X + xb * 2 * XTILE;
        if (DKV) {
          tma_load_2d(xs, &tm_qkv_x, &x_full[xb], colK + head * HD, urow0 + m0);
          tma_load_2d(xs + XTILE, &tm_qkv_x, &x_full[xb], colV + head * HD, urow0 + m0);
        } else {
          tma_load_2d(xs, &tm_qkv_x, &x_full[xb], colQ + head * HD, urow0 + m0);
          tma_load_2d(xs + XTILE, &tm_do_x, &x_full[xb], head * HD, urow0 + m0);
        }
        for (int j = 0; j < ny; ++j, ++yc) {
          const int s = yc % NS;
          mbar_wait(&y_empty[s], ((yc / NS) & 1) ^ 1);
          mbar_expect_tx(&y_full[s], 2 * YTILE);
          uint8_t* ys = smem + B_Y + s * 2 * YTILE;
          if (DKV) {
            tma_load_2d(ys, &tm_qkv_y, &y_full[s], colQ + head * HD, urow0 + j * BY);
            tma_load_2d(ys + YTILE, &tm_do_y, &y_full[s], head * HD, urow0 + j * BY);
          } else {
            tma_load_2d(ys, &tm_qkv_y, &y_full[s], colK + head * HD, urow0 + j * BY);
            tma_load_2d(ys + YTILE, &tm_qkv_y, &y_full[s], colV + head * HD, urow0 + j * BY);
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      // ===================== score issuer (whole warp runs the loop, one elected lane issues, see elect_one()): Sc1 = X1 Y1_j^T, Sc2 = X2 Y2_j^T into group (j & 1)'s tiles =================
      // (one tcgen05.mma issue costs the issuing thread ~130 cycles of dependent uniform-datapath work, so scores and the
      // two output products each have their own issuing warp; descriptors advance by adds in the address field.  A fourth
      // issuing warp would push the CTA to 13 warps and the register cap to 128: measured slower.)
      constexpr uint32_t idesc_sc = idesc_bf16(128, BY, false, false);     // X Y^T: both K-major
      uint32_t yc = 0;                         // ring position of block 0 of the current item
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int item = n_items - 1 - w;
        const int T = __ldg(&tab[item / heads]).y;
        const int ny = (T + BY - 1) / BY;
        const int xb = it & 1;
        const uint64_t x1 = umma_desc_sw128(sX + xb * 2 * XTILE), x2 = umma_desc_sw128(sX + xb * 2 * XTILE + XTILE);
        mbar_wait(&x_full[xb], (it >> 1) & 1);
        for (int j = 0; j < ny; ++j) {
          const uint32_t c = yc + j;
          const int s = c % NS;
          const int set = c % 3;
          const uint32_t use = c / 3;
          const uint32_t sc_col = set == 2 ? 384u : (uint32_t)set * 128u;
          if (use >= 1) mbar_wait(&sc_free[set], (use - 1) & 1);   // the block that used this set last is in registers
          mbar_wait(&y_full[s], (c / NS) & 1);
          tc_fence_after();
          const uint64_t y1 = umma_desc_sw128(sY + s * 2 * YTILE), y2 = umma_desc_sw128(sY + s * 2 * YTILE + YTILE);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) {                // two independent accumulation chains, interleaved
              umma_bf16_ss(tmem_base + sc_col, x1 + 2 * k, y1 + 2 * k, idesc_sc, k != 0);
              umma_bf16_ss(tmem_base + sc_col + 64, x2 + 2 * k, y2 + 2 * k, idesc_sc, k != 0);
            }
            umma_commit(&sc_full[set]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(&x_empty[xb]);
        __syncwarp();
        yc += ny;
      }
    }
  } else if (warp == 10 || warp == 11) {
    // ===================== output issuers: warp 10 -> accumulator 0 (dV | dQ), warp 11 -> accumulator 1 (dK, DKV only) ====
    const int which = warp - 10;
    if (DKV || which == 0) {
      constexpr uint32_t idesc_out = idesc_bf16(128, HD, false, true);     // A (K-major) x Y (MN-major)
      uint32_t yc = 0;
      uint32_t ac0 = 0, ac1 = 0;               // A-operand hand-overs from group 0 / 1 so far
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int item = n_items - 1 - w;
        const int T = __ldg(&tab[item / heads]).y;
        const int ny = (T + BY - 1) / BY;
        for (int j = 0; j < ny; ++j) {
          const int g = j & 1;
          const uint32_t c = yc + j;
          const int s = c % NS;
          const int nvalid = min(BY, T - j * BY);
          const int ksteps = (nvalid + 15) >> 4;           // reduction over the streamed rows that exist
          const uint32_t n = g ? ac1 : ac0;
          if (g) ++ac1; else ++ac0;
          mbar_wait(&a_full[g], n & 1);
          if (j == 0 && it > 0) mbar_wait(acc_empty, (it - 1) & 1);   // previous item's accumulators have been read
          tc_fence_after();
          // DKV: dV += P^T dO (A = P, B = Y2) on warp 10, dK += dS^T Q (A = dS, B = Y1) on warp 11;  DQ: dQ += dS K (A = dS, B = Y1)
          const bool use_p = DKV && which == 0;
          // DKV: [dS | P] single-buffered; DQ needs no P, so its dS operand is double-buffered in the same space
          const int ab = DKV ? 0 : (int)(n & 1);
          const uint64_t a_desc = umma_desc_sw128(sA + g * 2 * XTILE + (DKV ? (use_p ? XTILE : 0) : ab * XTILE));
          const uint64_t y_desc = umma_desc_sw128_mn(sY + s * 2 * YTILE + (use_p ? YTILE : 0));
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              if (ks < ksteps) umma_bf16_ss(tmem_base + 256 + which * 64, a_desc + 2 * ks, y_desc + 128 * ks, idesc_out, (j > 0 || ks > 0) ? 1u : 0u);
            umma_commit(&y_empty[s]);
            umma_commit(&a_empty[g * 2 + ab]);
          }
          __syncwarp();
        }
        if (elect_one()) umma_commit(acc_full);
        __syncwarp();
        yc += ny;
      }
    }
  } else {
    // ===================== elementwise groups: one resident row per thread =====================
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;                       // resident row = TMEM lane
    const int gi = (warp - 2 - 4 * g) * 32 + lane;     // thread index inside the group
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t a_ds0 = sA + g * 2 * XTILE + r * 128, a_p = a_ds0 + XTILE;
    float* cst0 = reinterpret_cast<float*>(smem + B_C + g * 1024);   // 2 x (lse[64] | D[64]) of the streamed blocks (DKV)
    float sc = scale_log2;
    asm volatile("" : "+f"(sc));
    uint32_t cnt = 0;                                  // blocks this group has processed (a_full / a_empty phase)
    uint32_t blk0 = 0;                                 // streamed blocks of this CTA before the current item (score-set rotation)
    int it = 0;
#ifdef ATTN_TIMING
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};      // constants + barrier | sc_full wait | a_empty wait | elementwise | fence + arrive | acc_full wait | epilogue + set-up
    long long tprev = clock64();
    const long long tstart = tprev;
    int nblocks = 0;
#define LAP(k) do { const long long tn = clock64(); tacc[k] += tn - tprev; tprev = tn; } while (0)
#else
#define LAP(k) do {} while (0)
#endif
    // ---- accumulators -> bf16 -> dqkv for a finished item.  DEFERRED: an item is drained after the group's first block of
    // the NEXT item, when the last output MMAs have long retired (waiting for them right away cost ~1.2 k cycles per item);
    // the next item's output MMAs wait for acc_empty meanwhile.  Results go through a staging tile of their own. ----
    struct Pending { int valid, it, head, urow0, m0, T; };
    Pending pend = {0, 0, 0, 0, 0, 0};
    auto drain = [&](const Pending& pd) {
      mbar_wait(acc_full, pd.it & 1);
      tc_fence_after();
      LAP(5);
      uint32_t o[32], o2[32];
      if (DKV) {                                       // group 0 stores dV, group 1 stores dK (* scale)
        tmem_ld_32x32(tmem_base + lane_off + 256 + g * 64, o);
        tmem_ld_32x32(tmem_base + lane_off + 256 + g * 64 + 32, o2);
      } else {                                         // group g stores columns [32 g, 32 g + 32) of dQ (* scale)
        tmem_ld_32x32(tmem_base + lane_off + 256 + g * 32, o);
      }
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(acc_empty);
      // Results leave through the warp's own 32 rows of the group's staging tile and one TMA store per warp when all 32
      // rows exist; scattered 16-byte stores (rows are 6 H bytes apart) cost ~2 k cycles per item.
      const bool warp_full = pd.m0 + q * 32 + 32 <= pd.T;
      const bool row_ok = pd.m0 + r < pd.T;
      const long long row = (long long)pd.urow0 + pd.m0 + r;
      const int head = pd.head, urow0 = pd.urow0, m0 = pd.m0;
      const uint32_t stage_w = sS + g * XTILE + q * 32 * 128, stage_r = sS + g * XTILE + r * 128;
      if (lane == 0) bulk_wait_read<0>();               // the previous item's store has finished reading the staging rows
      __syncwarp();
      if (DKV) {
        const float f = g ? scale : 1.0f;              // group 0 stores dV, group 1 stores dK (* scale)
        uint4 ov[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ov[i] = make_uint4(pack_bf16x2(__uint_as_float(o[8 * i]) * f, __uint_as_float(o[8 * i + 1]) * f),
                             pack_bf16x2(__uint_as_float(o[8 * i + 2]) * f, __uint_as_float(o[8 * i + 3]) * f),
                             pack_bf16x2(__uint_as_float(o[8 * i + 4]) * f, __uint_as_float(o[8 * i + 5]) * f),
                             pack_bf16x2(__uint_as_float(o[8 * i + 6]) * f, __uint_as_float(o[8 * i + 7]) * f));
          ov[4 + i] = make_uint4(pack_bf16x2(__uint_as_float(o2[8 * i]) * f, __uint_as_float(o2[8 * i + 1]) * f),
                                 pack_bf16x2(__uint_as_float(o2[8 * i + 2]) * f, __uint_as_float(o2[8 * i + 3]) * f),
                                 pack_bf16x2(__uint_as_float(o2[8 * i + 4]) * f, __uint_as_float(o2[8 * i + 5]) * f),
                                 pack_bf16x2(__uint_as_float(o2[8 * i + 6]) * f, __uint_as_float(o2[8 * i + 7]) * f));
        }
        const int col = (g ? colK : colV) + head * HD;
        if (warp_full) {
#pragma unroll
          for (int i = 0; i < 8; ++i) st_shared_v4(stage_r + ((i ^ (r & 7)) << 4), ov[i].x, ov[i].y, ov[i].z, ov[i].w);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tm_out, stage_w, col, (int)(urow0 + m0 + q * 32));
            bulk_commit();
          }
        } else if (row_ok) {
          uint4* op = reinterpret_cast<uint4*>(dqkv + row * (3LL * H) + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) op[i] = ov[i];
        }
      } else {
        uint4 ov[4];                                   // group g stores columns [32 g, 32 g + 32) of dQ (* scale)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          ov[i] = make_uint4(pack_bf16x2(__uint_as_float(o[8 * i]) * scale, __uint_as_float(o[8 * i + 1]) * scale),
                             pack_bf16x2(__uint_as_float(o[8 * i + 2]) * scale, __uint_as_float(o[8 * i + 3]) * scale),
                             pack_bf16x2(__uint_as_float(o[8 * i + 4]) * scale, __uint_as_float(o[8 * i + 5]) * scale),
                             pack_bf16x2(__uint_as_float(o[8 * i + 6]) * scale, __uint_as_float(o[8 * i + 7]) * scale));
        const int col = colQ + head * HD + g * 32;
        if (warp_full) {                               // 32 rows x 64 bytes, 64-byte swizzle
          const uint32_t wrow = stage_w + lane * 64;
          const int swz = (lane >> 1) & 3;
#pragma unroll
          for (int i = 0; i < 4; ++i) st_shared_v4(wrow + ((i ^ swz) << 4), ov[i].x, ov[i].y, ov[i].z, ov[i].w);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tm_out, stage_w, col, (int)(urow0 + m0 + q * 32));
            bulk_commit();
          }
        } else if (row_ok) {
          uint4* op = reinterpret_cast<uint4*>(dqkv + row * (3LL * H) + col);
#pragma unroll
          for (int i = 0; i < 4; ++i) op[i] = ov[i];
        }
      }
    };
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int item = n_items - 1 - w;
      const int4 t = __ldg(&tab[item / heads]);
      const int head = item - (item / heads) * heads;
      const int urow0 = t.x, T = t.y, m0 = t.z;
      const int ny = (T + BY - 1) / BY;
      const bool row_ok = m0 + r < T;
      const long long row = (long long)urow0 + m0 + r;
      float lse_r = INFINITY, d_r = 0.f;               // DQ: this thread's query row (+inf: rows past the end give P = 0)
      if (!DKV && row_ok) {
        lse_r = __ldg(LSE + (long long)head * M + row);
        d_r = __ldg(Dv + (long long)head * M + row);
      }
      // column constants (LSE, D of the streamed query block) of block jj for this thread's slot; past the utterance
      // end LSE = +inf makes P = 0, so those (foreign) query rows contribute nothing
      auto load_cst = [&](int jj) -> float {
        const int c = gi & 63;
        const bool ok = jj < ny && jj * BY + c < T;
        const long long qrow = (long long)head * M + urow0 + jj * BY + c;
        return gi < 64 ? (ok ? __ldg(LSE + qrow) : INFINITY) : (ok ? __ldg(Dv + qrow) : 0.f);
      };
      float cst_next = 0.f;
      if (DKV) cst_next = load_cst(g);
      bool drained = false;
      for (int j = g; j < ny; j += 2, ++cnt) {
        const int nvalid = min(BY, T - j * BY);
        float* cst = cst0 + (cnt & 1) * 128;
        const uint32_t cst_u = smem_u32(cst);
        LAP(6);
        if (DKV) {
          // double-buffered: the value was fetched one block ahead, so its global latency is off the chain, and one
          // barrier per block suffices (whoever passes it has finished reading the buffer written two blocks ago)
          cst[gi] = cst_next;
          cst_next = load_cst(j + 2);
          named_bar_sync(2 + g, 128);
        }
        LAP(0);
        const uint32_t c = blk0 + j;
        const int set = c % 3;
        const uint32_t t1 = tmem_base + lane_off + (set == 2 ? 384u : (uint32_t)set * 128u), t2 = t1 + 64;
        mbar_wait(&sc_full[set], (c / 3) & 1);
        tc_fence_after();
        LAP(1);
        // the MMAs that read the operand buffer about to be rewritten have retired (DKV: the previous block's; DQ: the
        // dS buffer alternates, so the block before the previous one's)
        const int ab = DKV ? 0 : (int)(cnt & 1);
        const uint32_t a_ds = a_ds0 + ab * XTILE;
        mbar_wait(&a_empty[g * 2 + ab], (DKV ? (cnt & 1) : ((cnt >> 1) & 1)) ^ 1);
        LAP(2);
#pragma unroll
        for (int h = 0; h < 2; ++h) {                  // two halves of 32 streamed rows (columns of the score tiles)
          uint32_t s1[32], s2[32];
          tmem_ld_32x32(t1 + 32 * h, s1);
          tmem_ld_32x32(t2 + 32 * h, s2);
          tmem_ld_wait();
          if (h == 1) {                                // the whole set is in registers: the score issuer may reuse it
            tc_fence_before();
            mbar_arrive(&sc_free[set]);
          }
          uint32_t pds[16], pp[16];                    // packed bf16 pairs: dS (and P for DKV)
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float l0, l1, d0, d1;
            if (DKV) {
              const float2 lv = ld_shared_v2(cst_u + (32 * h + 2 * i) * 4);
              const float2 dv = ld_shared_v2(cst_u + 256 + (32 * h + 2 * i) * 4);
              l0 = lv.x; l1 = lv.y; d0 = dv.x; d1 = dv.y;
            } else {
              l0 = l1 = lse_r; d0 = d1 = d_r;
            }
            float p0 = ex2_approx(fmaf(__uint_as_float(s1[2 * i]), sc, -l0));
            float p1 = ex2_approx(fmaf(__uint_as_float(s1[2 * i + 1]), sc, -l1));
            if (!DKV && nvalid < BY) {                 // keys past the utterance end
              p0 = 32 * h + 2 * i < nvalid ? p0 : 0.f;
              p1 = 32 * h + 2 * i + 1 < nvalid ? p1 : 0.f;
            }
            const float e0 = p0 * (__uint_as_float(s2[2 * i]) - d0), e1 = p1 * (__uint_as_float(s2[2 * i + 1]) - d1);
            pds[i] = pack_bf16x2(e0, e1);
            if (DKV) pp[i] = pack_bf16x2(p0, p1);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            st_shared_v4(a_ds + (((4 * h + i) ^ (r & 7)) << 4), pds[4 * i], pds[4 * i + 1], pds[4 * i + 2], pds[4 * i + 3]);
            if (DKV) st_shared_v4(a_p + (((4 * h + i) ^ (r & 7)) << 4), pp[4 * i], pp[4 * i + 1], pp[4 * i + 2], pp[4 * i + 3]);
          }
        }
        LAP(3);
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&a_full[g]);
        if (!drained) {                                // first block of the item done: now drain the previous item
          if (pend.valid) drain(pend);
          drained = true;
        }
        LAP(4);
#ifdef ATTN_TIMING
        ++nblocks;
#endif
      }
      blk0 += ny;
      if (!drained && pend.valid) drain(pend);         // this group had no block in the item
      pend = Pending{1, it, head, urow0, m0, T};
    }
    if (pend.valid) drain(pend);
    if (lane == 0) bulk_wait<0>();                     // shared memory must outlive the last store's reads
#ifdef ATTN_TIMING
    LAP(6);
    if ((threadIdx.x == 64 || threadIdx.x == 192) && blockIdx.x == 3)
      printf("attn bwd dkv=%d cta %d thread %d: items %d blocks %d total %lld | cst %lld sc_wait %lld a_wait %lld elementwise %lld fence %lld acc_wait %lld epi %lld\n",
             (int)DKV, blockIdx.x, threadIdx.x, it, nblocks, clock64() - tstart, tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5], tacc[6]);
#endif
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// D[h][row] = sum_d dO[row, h, d] * O[row, h, d]
// 8 threads per (row, head): one 16-byte piece of O and dO each, 3-step shuffle reduction.  A block covers 32 consecutive
// rows of one head, so the reads are 128-byte segments and the 32 results are one contiguous 128-byte store (one thread
// per (row, head) wrote D with a stride of M floats between neighbouring threads).
__global__ void __launch_bounds__(256)
attn_bwd_prep_tc_kernel(const bf16* __restrict__ O, const bf16* __restrict__ dO, float* __restrict__ D, int H, long long M) {
  const int h = blockIdx.y, sub = threadIdx.x & 7;
  const long long row = (long long)blockIdx.x * 32 + (threadIdx.x >> 3);
  float acc = 0.f;
  if (row < M) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(O + row * H + h * HD) + sub);
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(dO + row * H + h * HD) + sub);
    float2 x, y;
    x = unpack_bf16x2(a.x); y = unpack_bf16x2(b.x); acc += x.x * y.x + x.y * y.y;
    x = unpack_bf16x2(a.y); y = unpack_bf16x2(b.y); acc += x.x * y.x + x.y * y.y;
    x = unpack_bf16x2(a.z); y = unpack_bf16x2(b.z); acc += x.x * y.x + x.y * y.y;
    x = unpack_bf16x2(a.w); y = unpack_bf16x2(b.w); acc += x.x * y.x + x.y * y.y;
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (sub == 0 && row < M) D[(long long)h * M + row] = acc;
}

}  // namespace

int attention_backward(const bf16* qkv, const bf16* O, const bf16* dO, const float* LSE, float* D, bf16* dqkv,
                          const int4* blk_tab, int n_blk, int H, int heads, long long M, cudaStream_t stream) {
  SUTA_CHECK_ARG(H == heads * HD);
  if (n_blk <= 0) return SUTA_OK;
  static bool attr = false;
  static int n_sm = 148;
  if (!attr) {
    CUDA_TRY(cudaFuncSetAttribute(attn_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    CUDA_TRY(cudaFuncSetAttribute(attn_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    attr = true;
  }
  CUtensorMap qx, qy, dx, dy;
  SUTA_TRY(make_map(&qx, qkv, M, 3LL * H, 3LL * H, BX));
  SUTA_TRY(make_map(&qy, qkv, M, 3LL * H, 3LL * H, BY));
  SUTA_TRY(make_map(&dx, dO, M, H, H, BX));
  SUTA_TRY(make_map(&dy, dO, M, H, H, BY));
  CUtensorMap out128, out64;                                 // result stores: 32-row boxes of 64 (dK, dV) / 32 (dQ halves) columns
  SUTA_TRY(make_map(&out128, dqkv, M, 3LL * H, 3LL * H, 32, 64));
  SUTA_TRY(make_map(&out64, dqkv, M, 3LL * H, 3LL * H, 32, 32));
  const float scale = 0.125f, scale_log2 = scale * 1.4426950408889634f;
  attn_bwd_prep_tc_kernel<<<dim3((unsigned)((M + 31) / 32), (unsigned)heads), 256, 0, stream>>>(O, dO, D, H, M);
  const long long n_items = (long long)n_blk * heads;
  const int grid = (int)(n_items < n_sm ? n_items : n_sm);
  attn_bwd_tc_kernel<true><<<grid, BWD_THREADS, B_SMEM, stream>>>(qx, qy, dx, dy, out128, LSE, D, dqkv, blk_tab, n_blk, heads, H, M, scale, scale_log2);
  attn_bwd_tc_kernel<false><<<grid, BWD_THREADS, B_SMEM, stream>>>(qx, qy, dx, dy, out64, LSE, D, dqkv, blk_tab, n_blk, heads, H, M, scale, scale_log2);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}
