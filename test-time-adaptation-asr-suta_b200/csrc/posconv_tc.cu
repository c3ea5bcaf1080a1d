// Grouped positional convolution (HF/modeling_wav2vec2.py:360-368: Conv1d(H -> H, k = 128, groups = 16), and its dgrad with
// the flipped weights) as a tcgen05 kernel whose A operand is NOT re-fetched per tap.
//
// As a plain GEMM (gemm_tc.cu, N = 48 per group, K = 128 taps x 48 channels) every 64-deep k-block re-loads a 128-row
// window tile of 16 KB although consecutive taps read the same rows shifted by one: the kernel ran at 30 % of the tensor
// pipe, bound by the ~70 B/clk one SM can ingest through TMA.  Here a CTA keeps the input window of its 256 output rows
// (256 + K - 1 rows x 48 channels) in shared memory in a PLANAR, un-swizzled layout [channel octet][row][8 channels]:
// rows are 16 bytes apart inside an octet plane, so in the canonical K-major no-swizzle UMMA layout (core matrix = 8 rows x
// 16 B, stride between 8-row groups = 128 B, leading offset between octets = one plane) the A tile of tap t is the
// tile of tap 0 moved by t * 16 bytes -- one add on the descriptor.  Only the weights stream: 4.6 KB per tap for two
// 128 x 48 x 48 MMAs groups (32 B/clk).
//
//   out[r, g*CG + o] = bias[g*CG + o] + sum_{tap, c} xg[g][r + tap][c] * w[g*CG + o][tap*CG + c]      r < Rm
//
// Warps: 0 TMA producer (window per item, weight ring), 1 MMA issuer, 2-5 epilogue (TMEM -> registers -> fp32 rows).
#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace {

constexpr int PC_THREADS = 192;
constexpr int PC_MT = 256;                 // output rows per item (two 128-row accumulators)
constexpr int PC_WIN_ROWS = 384;           // window rows held per item (>= PC_MT + pos_k - 1)
constexpr int PC_WSTAGES = 3;

template <int CG>
struct PcCfg {
  static constexpr int PLANES = CG / 8;
  static constexpr int PLANE_BYTES = PC_WIN_ROWS * 16;
  static constexpr int WIN_BYTES = PLANES * PLANE_BYTES;
  static constexpr int TPS = CG == 48 ? 8 : 4;                     // taps per weight stage
  static constexpr int TAP_BYTES = PLANES * CG * 16;               // [octet][o][8 c]
  static constexpr int WSTAGE_BYTES = TPS * TAP_BYTES;
  static constexpr int SMEM_BYTES = 2 * WIN_BYTES + PC_WSTAGES * WSTAGE_BYTES + 256;
};

struct PcParams {
  int R, Rm, G, pos_k, n_items;
  float* out;
  int out_ld;
  const float* bias;
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// K-major, no swizzle: core matrices of 8 rows x 16 bytes; `lbo` = bytes between core matrices adjacent in K,
// `sbo` = bytes between core matrices adjacent in M|N
__device__ __forceinline__ uint64_t umma_desc_plain(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;                                              // layout type 0: SWIZZLE_NONE
}

template <int CG>
__global__ void __launch_bounds__(PC_THREADS, 1)
posconv_tc_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w, const PcParams p) {
  using C = PcCfg<CG>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* win = smem;                                   // [2][PLANES][PC_WIN_ROWS][16 B]
  uint8_t* wst = smem + 2 * C::WIN_BYTES;                // [PC_WSTAGES][TPS][PLANES][CG][16 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(wst + PC_WSTAGES * C::WSTAGE_BYTES);
  uint64_t* win_full = bars;                             // [2]
  uint64_t* win_empty = bars + 2;                        // [2]
  uint64_t* w_full = bars + 4;                           // [PC_WSTAGES]
  uint64_t* w_empty = bars + 4 + PC_WSTAGES;             // [PC_WSTAGES]
  uint64_t* acc_full = bars + 4 + 2 * PC_WSTAGES;        // [2]
  uint64_t* acc_empty = acc_full + 2;                    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_x);
    tma_prefetch_desc(&tma_w);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&win_full[i], 1);
      mbar_init(&win_empty[i], 1);
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);
    }
    for (int i = 0; i < PC_WSTAGES; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_wstages = p.pos_k / C::TPS;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int ws = 0;
      uint32_t wphase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
        const int g = item % p.G, r0 = (item / p.G) * PC_MT;
        const int wb = it & 1;
        mbar_wait(&win_empty[wb], ((it >> 1) & 1) ^ 1);
        mbar_expect_tx(&win_full[wb], C::WIN_BYTES);
        uint8_t* wdst = win + wb * C::WIN_BYTES;
#pragma unroll 1
        for (int pl = 0; pl < C::PLANES; ++pl)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            tma_load_2d(wdst + pl * C::PLANE_BYTES + h * (PC_WIN_ROWS / 2) * 16, &tma_x, &win_full[wb], pl * 8,
                        g * p.R + r0 + h * (PC_WIN_ROWS / 2));
        for (int s = 0; s < n_wstages; ++s) {
          mbar_wait(&w_empty[ws], wphase ^ 1);
          mbar_expect_tx(&w_full[ws], C::WSTAGE_BYTES);
          tma_load_3d(wst + ws * C::WSTAGE_BYTES, &tma_w, &w_full[ws], 0, g * CG, s * C::TPS * C::PLANES);
          if (++ws == PC_WSTAGES) { ws = 0; wphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(CG);
    int ws = 0;
    uint32_t wphase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      const int wb = it & 1, ab = it & 1;
      mbar_wait(&acc_empty[ab], ((it >> 1) & 1) ^ 1);
      mbar_wait(&win_full[wb], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t d0 = tmem_base + ab * 128;
      const uint64_t a_base = umma_desc_plain(smem_u32(win + wb * C::WIN_BYTES), C::PLANE_BYTES, 128);
      for (int s = 0; s < n_wstages; ++s) {
        mbar_wait(&w_full[ws], wphase);
        tc_fence_after();
        const uint64_t b_base = umma_desc_plain(smem_u32(wst + ws * C::WSTAGE_BYTES), CG * 16, 128);
        if (elect_one()) {
          // fully unrolled: every descriptor is (per-stage base) + (compile-time constant); a 128 x 48 x 16 MMA takes
          // ~24 cycles, so the issue loop itself must stay at a handful of uniform-datapath instructions per MMA
          const uint64_t a_stage = a_base + (uint64_t)(s * C::TPS);
          const uint32_t acc0 = s != 0 ? 1u : 0u;
#pragma unroll
          for (int t = 0; t < C::TPS; ++t) {
#pragma unroll
            for (int kk = 0; kk < CG / 16; ++kk) {
              // descriptor start addresses are in 16-byte units: one row = 1, 128 rows = 128, two octet planes
              const uint64_t da = a_stage + (uint64_t)(t + kk * 2 * (C::PLANE_BYTES >> 4));
              const uint64_t db = b_base + (uint64_t)((t * C::TAP_BYTES + kk * 2 * CG * 16) >> 4);
              const uint32_t acc = (t | kk) != 0 ? 1u : acc0;
              umma_bf16_ss(d0, da, db, idesc, acc);
              umma_bf16_ss(d0 + 64, da + 128, db, idesc, acc);
            }
          }
          umma_commit(&w_empty[ws]);
        }
        __syncwarp();
        if (++ws == PC_WSTAGES) { ws = 0; wphase ^= 1; }
      }
      if (elect_one()) {
        umma_commit(&win_empty[wb]);
        umma_commit(&acc_full[ab]);
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue: accumulators -> fp32 rows (+ bias) =====================
    const int q = warp & 3;
    int it = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++it) {
      const int g = item % p.G, r0 = (item / p.G) * PC_MT;
      const int ab = it & 1;
      mbar_wait(&acc_full[ab], (it >> 1) & 1);
      tc_fence_after();
      uint32_t v[2][CG];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 128 + h * 64;
        uint32_t r32[32];
        tmem_ld_32x32(ta, r32);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[h][i] = r32[i];
        if constexpr (CG == 48) {
          uint32_t r16[16];
          tmem_ld_32x16(ta + 32, r16);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[h][32 + i] = r16[i];
        } else {
          uint32_t r2[32];
          tmem_ld_32x32(ta + 32, r2);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[h][32 + i] = r2[i];
        }
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[ab]);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const long long row = (long long)r0 + h * 128 + q * 32 + lane;
        if (row < p.Rm) {
          float* o = p.out + row * p.out_ld + g * CG;
#pragma unroll
          for (int c = 0; c < CG; c += 4) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) b = __ldg(reinterpret_cast<const float4*>(p.bias + g * CG + c));
            *reinterpret_cast<float4*>(o + c) = make_float4(__uint_as_float(v[h][c]) + b.x, __uint_as_float(v[h][c + 1]) + b.y,
                                                            __uint_as_float(v[h][c + 2]) + b.z, __uint_as_float(v[h][c + 3]) + b.w);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

template <int CG>
int launch_pc(const bf16* xg, const bf16* w, const float* bias, float* out, int out_ld, int G, long long R, long long Rm, int pos_k,
              cudaStream_t stream) {
  using C = PcCfg<CG>;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(posconv_tc_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap tx, tw;
  {  // window source: [G*R (+8) rows][CG] bf16, boxes of 8 channels x 192 rows
    const long long dims[2] = {CG, (long long)G * R + 8};
    const long long strides[1] = {CG * 2};
    const int box[2] = {8, PC_WIN_ROWS / 2};
    SUTA_TRY(gemm_encode_tmap_nd(&tx, xg, 2, dims, strides, box));
  }
  {  // weights [G*CG rows (o)][pos_k*CG] as (8 channels, o rows, channel octets): one box = TPS taps of one group
    const long long dims[3] = {8, (long long)G * CG, (long long)pos_k * CG / 8};
    const long long strides[2] = {(long long)pos_k * CG * 2, 16};
    const int box[3] = {8, CG, C::TPS * C::PLANES};
    SUTA_TRY(gemm_encode_tmap_nd(&tw, w, 3, dims, strides, box));
  }
  PcParams p;
  p.R = (int)R; p.Rm = (int)Rm; p.G = G; p.pos_k = pos_k;
  p.n_items = (int)((Rm + PC_MT - 1) / PC_MT) * G;
  p.out = out; p.out_ld = out_ld; p.bias = bias;
  const int grid = p.n_items < gemm_num_sms() ? p.n_items : gemm_num_sms();
  posconv_tc_kernel<CG><<<grid, PC_THREADS, C::SMEM_BYTES, stream>>>(tx, tw, p);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

}  // namespace

bool posconv_tc_supported(int CG, int pos_k) {
  if (CG != 48 && CG != 64) return false;
  const int tps = CG == 48 ? 8 : 4;
  return pos_k % tps == 0 && PC_MT + pos_k - 1 <= PC_WIN_ROWS;
}

// xg: [G][R][CG] bf16 group slabs; w: [G*CG][pos_k*CG] bf16 ((tap, c) along k); out: fp32 [Rm][out_ld], group g at columns g*CG
int posconv_tc(const bf16* xg, const bf16* w, const float* bias, float* out, int out_ld, int G, int CG, long long R, long long Rm,
               int pos_k, cudaStream_t stream) {
  SUTA_CHECK_ARG(xg && w && out && posconv_tc_supported(CG, pos_k) && out_ld % 4 == 0 && Rm > 0 && R >= Rm);
  if (CG == 48) return launch_pc<48>(xg, w, bias, out, out_ld, G, R, Rm, pos_k, stream);
  return launch_pc<64>(xg, w, bias, out, out_ld, G, R, Rm, pos_k, stream);
}
