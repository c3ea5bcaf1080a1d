// tcgen05 + TMEM + TMA GEMM for sm_100a.  See gemm_tc.cuh for what it computes.
//
// Structure (one persistent CTA per SM, 6 warps, warp-specialised):
//   warp 0      TMA producer: cp.async.bulk.tensor 128x64 (A) and BNx64 (B) bf16 boxes, 128B swizzle,
//               into a STAGES-deep shared-memory ring guarded by full/empty mbarriers
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M=128, N=BN, K=16) x4 per stage into a
//               double-buffered fp32 accumulator in tensor memory; tcgen05.commit releases the stage
//   warps 2..9  epilogue: two warps per TMEM lane quarter, each owning half of the tile's columns;
//               tcgen05.ld the accumulator (one row per thread), apply bias / GELU / GELU' /
//               residual, store fp32 and/or bf16 straight to global; overlaps the next tile's mainloop
#include "gemm_tc.cuh"

#include <mutex>

namespace {

constexpr int BM = 128;
constexpr int BK = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int GEMM_THREADS = 320;   // TMA warp + MMA warp + 8 epilogue warps

struct GemmKernelParams {
  int M, N, K;
  int num_mblk, num_nblk, nz;
  long long a_z_rows, b_z_rows;
  int c_z_cols;
  const int4* mblk;
  const int4* ztab;
  long long out_z_stride;
  GemmEpilogue epi;
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (196 * 1024 / STAGE_BYTES) > 8 ? 8 : (196 * 1024 / STAGE_BYTES);
  static constexpr int ACC_STRIDE = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  static constexpr int CH = (BN % 32 == 0) ? 32 : 16;       // epilogue column chunk
  static constexpr int EPI_SPLIT = (BN >= 128) ? 2 : 1;      // epilogue warps per TMEM lane quarter
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int CH>
__device__ __forceinline__ void epilogue_chunk(const GemmEpilogue& e, float (&v)[CH], long long row, int oc,
                                               long long bias_off, long long out_off) {
  if (e.bias) {
    const float4* bp = reinterpret_cast<const float4*>(e.bias + bias_off + oc);
#pragma unroll
    for (int i = 0; i < CH / 4; ++i) {
      float4 b = __ldg(bp + i);
      v[4 * i + 0] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
    }
  }
  if (e.act == 1) {
    if (e.aux_out) {
      uint4* ap = reinterpret_cast<uint4*>(e.aux_out + row * e.aux_ld + oc);
#pragma unroll
      for (int i = 0; i < CH / 8; ++i) {
        uint4 u;
        u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]); u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
        u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]); u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
        ap[i] = u;
      }
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] = gelu_erf(v[i]);
  } else if (e.act == 2) {
    const uint4* ap = reinterpret_cast<const uint4*>(e.aux_in + row * e.aux_ld + oc);
#pragma unroll
    for (int i = 0; i < CH / 8; ++i) {
      uint4 u = __ldg(ap + i);
      float2 f0 = unpack_bf16x2(u.x), f1 = unpack_bf16x2(u.y), f2 = unpack_bf16x2(u.z), f3 = unpack_bf16x2(u.w);
      v[8 * i + 0] *= gelu_erf_grad(f0.x); v[8 * i + 1] *= gelu_erf_grad(f0.y);
      v[8 * i + 2] *= gelu_erf_grad(f1.x); v[8 * i + 3] *= gelu_erf_grad(f1.y);
      v[8 * i + 4] *= gelu_erf_grad(f2.x); v[8 * i + 5] *= gelu_erf_grad(f2.y);
      v[8 * i + 6] *= gelu_erf_grad(f3.x); v[8 * i + 7] *= gelu_erf_grad(f3.y);
    }
  }
  if (e.residual) {
    const float4* rp = reinterpret_cast<const float4*>(e.residual + row * e.res_ld + oc);
#pragma unroll
    for (int i = 0; i < CH / 4; ++i) {
      float4 r = __ldg(rp + i);
      v[4 * i + 0] += r.x; v[4 * i + 1] += r.y; v[4 * i + 2] += r.z; v[4 * i + 3] += r.w;
    }
  }
  if (e.out_f32) {
    float4* op = reinterpret_cast<float4*>(e.out_f32 + out_off + row * e.out_ld + oc);
#pragma unroll
    for (int i = 0; i < CH / 4; ++i) op[i] = make_float4(v[4 * i + 0], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
  if (e.out_bf16) {
    uint4* op = reinterpret_cast<uint4*>(e.out_bf16 + out_off + row * e.out_ld + oc);
#pragma unroll
    for (int i = 0; i < CH / 8; ++i) {
      uint4 u;
      u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]); u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
      u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]); u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
      op[i] = u;
    }
  }
}

// A_MN / B_MN: the operand is MN-major in memory ([K rows][M|N contiguous]); its tile is fetched as 64x64 TMA boxes
// (64 M|N elements = one 128-byte swizzle row, 64 k-rows) laid 8 KB apart, described to the tensor core with
// leading-dimension byte offset 8192 (next 64 M|N elements) and stride byte offset 1024 (next 8 k-rows).
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const GemmKernelParams p) {
  using C = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles must start on 1024-byte boundaries
  uint8_t* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + C::STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + C::STAGES;          // [STAGES]
  uint64_t* acc_full = bars + 2 * C::STAGES;       // [2]
  uint64_t* acc_empty = bars + 2 * C::STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], 4 * C::EPI_SPLIT);
    }
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
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
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
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_2d(sa + C::A_BYTES + i * 8192, &tma_b, &full_bar[stage], n_blk * BN + (int)(z * p.b_z_rows) + i * 64,
                          b_off + b_k0 + kb * BK);
          } else {
            tma_load_2d(sa + C::A_BYTES, &tma_b, &full_bar[stage], b_k0 + kb * BK, b_row);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BN) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int num_kblk = num_kblk_all;
        if (p.ztab) num_kblk = (__ldg(&p.ztab[tile / tiles_per_z]).z + BK - 1) / BK;
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * C::ACC_STRIDE;
        for (int kb = 0; kb < num_kblk; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(tiles + stage * C::STAGE_BYTES);
          const uint32_t b_addr = a_addr + C::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = A_MN ? umma_desc_sw128_mn(a_addr + k * 2048) : umma_desc_sw128(a_addr + k * 32);
            const uint64_t db = B_MN ? umma_desc_sw128_mn(b_addr + k * 2048) : umma_desc_sw128(b_addr + k * 32);
            umma_bf16_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);          // frees the smem stage once these MMAs retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[acc]);               // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;              // which column half of the tile this warp owns
    constexpr int COLS_PER_WARP = BN / C::EPI_SPLIT;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; half < C::EPI_SPLIT && tile < total_tiles; tile += gridDim.x) {
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
      const int row_in_tile = q * 32 + lane;
      const bool row_ok = row_in_tile < rows_valid;
      const long long row = (long long)out_row0 + row_in_tile;

      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * C::ACC_STRIDE;
#pragma unroll 1
      for (int c0 = half * COLS_PER_WARP; c0 < (half + 1) * COLS_PER_WARP; c0 += C::CH) {
        uint32_t raw[C::CH];
        if constexpr (C::CH == 32) tmem_ld_32x32(t_addr + c0, raw); else tmem_ld_32x16(t_addr + c0, raw);
        tmem_ld_wait();
        const int col = n_blk * BN + c0;
        if (row_ok && col < p.N) {
          float v[C::CH];
#pragma unroll
          for (int i = 0; i < C::CH; ++i) v[i] = __uint_as_float(raw[i]);
          const long long bias_off = p.epi.bias_utt_stride ? (long long)(b_off / p.N) * p.epi.bias_utt_stride : (long long)b_off;
          epilogue_chunk<C::CH>(p.epi, v, row, z * p.c_z_cols + col, bias_off, (long long)z * p.out_z_stride);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
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

int make_tmap(CUtensorMap* tm, const GemmOperand& op, int K, int box_rows);

// MN-major operand: dims {cols (M|N, contiguous), rows (K)}, 64x64 boxes
int make_tmap_mn(CUtensorMap* tm, const GemmOperand& op) {
  GemmOperand t = op;
  return make_tmap(tm, t, (int)op.cols, 64);
}

int make_tmap(CUtensorMap* tm, const GemmOperand& op, int K, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    suta_set_last_error("cuTensorMapEncodeTiled entry point not available");
    return SUTA_ERR_DRIVER;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)op.rows};
  cuuint64_t strides[1] = {(cuuint64_t)op.row_stride * sizeof(bf16)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  if ((reinterpret_cast<uintptr_t>(op.ptr) & 15) || (strides[0] & 15) || op.rows <= 0) {
    suta_set_last_error("gemm operand not TMA-compatible: ptr=%p row_stride=%lld rows=%lld", (const void*)op.ptr,
                        op.row_stride, op.rows);
    return SUTA_ERR_ARG;
  }
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(op.ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    suta_set_last_error("cuTensorMapEncodeTiled failed (%d): K=%d rows=%lld stride=%lld box_rows=%d", (int)r, K,
                        op.rows, op.row_stride, box_rows);
    return SUTA_ERR_DRIVER;
  }
  return SUTA_OK;
}

template <int BN, bool A_MN, bool B_MN>
int launch(const GemmProblem& p, cudaStream_t stream) {
  using C = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_TRY(cudaFuncSetAttribute(gemm_bf16_tc_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap ta, tb;
  if (A_MN) SUTA_TRY(make_tmap_mn(&ta, p.a)); else SUTA_TRY(make_tmap(&ta, p.a, p.K, BM));
  if (B_MN) SUTA_TRY(make_tmap_mn(&tb, p.b)); else SUTA_TRY(make_tmap(&tb, p.b, p.K, BN));
  GemmKernelParams kp;
  kp.M = p.M; kp.N = p.N; kp.K = p.K;
  kp.num_mblk = p.mblk ? p.num_mblk : ceil_div(p.M, BM);
  kp.num_nblk = ceil_div(p.N, BN);
  kp.nz = p.nz;
  kp.a_z_rows = p.a_z_rows; kp.b_z_rows = p.b_z_rows; kp.c_z_cols = p.c_z_cols;
  kp.mblk = p.mblk;
  kp.ztab = p.ztab;
  kp.out_z_stride = p.out_z_stride;
  kp.epi = p.epi;
  long long total = (long long)kp.num_mblk * kp.num_nblk * kp.nz;
  if (total <= 0) return SUTA_OK;
  int grid = (int)(total < gemm_num_sms() ? total : gemm_num_sms());
  gemm_bf16_tc_kernel<BN, A_MN, B_MN><<<grid, GEMM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, kp);
  CUDA_TRY(cudaGetLastError());
  return SUTA_OK;
}

}  // namespace

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
  if (p.a.mn_major || p.b.mn_major) {
    SUTA_CHECK_ARG(p.N % 64 == 0 && (!p.a.mn_major || p.b.mn_major));
    if (p.a.mn_major) {
      if (p.N % 256 == 0) return launch<256, true, true>(p, stream);
      if (p.N % 128 == 0) return launch<128, true, true>(p, stream);
      return launch<64, true, true>(p, stream);
    }
    if (p.N % 256 == 0) return launch<256, false, true>(p, stream);
    if (p.N % 128 == 0) return launch<128, false, true>(p, stream);
    return launch<64, false, true>(p, stream);
  }
  if (p.N % 256 == 0) return launch<256, false, false>(p, stream);
  if (p.N % 128 == 0) return launch<128, false, false>(p, stream);
  if (p.N % 64 == 0) return launch<64, false, false>(p, stream);
  if (p.N % 48 == 0) return launch<48, false, false>(p, stream);
  if (p.N % 32 == 0) return launch<32, false, false>(p, stream);
  suta_set_last_error("gemm: N=%d must be a multiple of 32 or 48", p.N);
  return SUTA_ERR_ARG;
}
