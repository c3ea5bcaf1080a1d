// Shared device/host helpers for the suta_b200 kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#ifndef __CUDA_ARCH_FEAT_SM100_ALL
#if defined(__CUDA_ARCH__)
#error "suta_b200 kernels must be compiled with -gencode arch=compute_100a,code=sm_100a"
#endif
#endif

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

#define SUTA_OK 0
#define SUTA_ERR_CUDA 1
#define SUTA_ERR_ARG 2
#define SUTA_ERR_NOMEM 3
#define SUTA_ERR_DRIVER 4

extern "C" void suta_set_last_error(const char* fmt, ...);

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      suta_set_last_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return SUTA_ERR_CUDA;                                                                 \
    }                                                                                       \
  } while (0)

#define SUTA_TRY(expr)            \
  do {                            \
    int _s = (expr);              \
    if (_s != SUTA_OK) return _s; \
  } while (0)

#define SUTA_CHECK_ARG(cond)                                                    \
  do {                                                                          \
    if (!(cond)) {                                                              \
      suta_set_last_error("%s:%d argument check failed: %s", __FILE__, __LINE__, #cond); \
      return SUTA_ERR_ARG;                                                      \
    }                                                                           \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// erf(|z|) = 1 - 2^p(|z|), p a degree-5 polynomial fitted (weighted minimax, tools/fit_erf.py) to log2(erfc) on [0,4]:
// |error| <= 6.8e-7 in fp32 over the whole line, p is monotone decreasing so large |z| saturates to 1.  One MUFU.EX2 +
// 6 FMA-pipe instructions -- the GEMM epilogues evaluate it for every element of the FFN / conv activations, where the
// instruction budget per element decides whether the tensor pipe or the epilogue warps set the pace.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float erf_abs_fast(float t) {   // t >= 0
  float p = fmaf(-0.002944157226011157f, t, 0.02959004044532776f);
  p = fmaf(p, t, -0.1486656218767166f);
  p = fmaf(p, t, -0.9185093641281128f);
  p = fmaf(p, t, -1.6278890371322632f);
  return 1.0f - ex2_approx(p * t);
}
__device__ __forceinline__ float erf_fast(float z) { return copysignf(erf_abs_fast(fabsf(z)), z); }
__device__ __forceinline__ float gelu_erf(float x) {
  const float hx = 0.5f * x;
  return fmaf(hx, erf_fast(x * 0.70710678118654752f), hx);
}
// d/dx [ x * Phi(x) ] = Phi(x) + x * phi(x)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = fmaf(0.5f, erf_fast(x * 0.70710678118654752f), 0.5f);
  const float pdf = 0.39894228040143268f * ex2_approx(-0.72134752044448170f * x * x);
  return fmaf(x, pdf, cdf);
}
// both at once (shared erf): the forward epilogues store GELU'(pre) in bf16 so the backward only multiplies
__device__ __forceinline__ void gelu_erf_both(float x, float& g, float& dg) {
  const float e = erf_fast(x * 0.70710678118654752f);
  const float hx = 0.5f * x;
  g = fmaf(hx, e, hx);
  const float pdf = 0.39894228040143268f * ex2_approx(-0.72134752044448170f * x * x);
  dg = fmaf(x, pdf, fmaf(0.5f, e, 0.5f));
}

// ---- packed fp32 pairs (Blackwell FFMA2 / FMUL2: two fp32 lanes per instruction) ------------------------------
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint64_t dup2(float a) { return pk2(a, a); }
__device__ __forceinline__ void upk2(uint64_t r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// gelu_erf / gelu_erf_grad on two values (packed fp32): the same rounded operations as the scalar forms
__device__ __forceinline__ uint64_t erf_fast2(uint64_t z, float z0, float z1) {      // z = (z0, z1)
  const uint64_t t = pk2(fabsf(z0), fabsf(z1));
  uint64_t p = fma2(dup2(-0.002944157226011157f), t, dup2(0.02959004044532776f));
  p = fma2(p, t, dup2(-0.1486656218767166f));
  p = fma2(p, t, dup2(-0.9185093641281128f));
  p = fma2(p, t, dup2(-1.6278890371322632f));
  p = mul2(p, t);
  float p0, p1;
  upk2(p, p0, p1);
  return pk2(copysignf(1.0f - ex2_approx(p0), z0), copysignf(1.0f - ex2_approx(p1), z1));
}
__device__ __forceinline__ uint64_t gelu_erf2(uint64_t x) {
  const uint64_t z = mul2(x, dup2(0.70710678118654752f));
  float z0, z1;
  upk2(z, z0, z1);
  const uint64_t hx = mul2(x, dup2(0.5f));
  return fma2(hx, erf_fast2(z, z0, z1), hx);
}
__device__ __forceinline__ uint64_t gelu_erf_grad2(uint64_t x) {
  const uint64_t z = mul2(x, dup2(0.70710678118654752f));
  float z0, z1, q0, q1;
  upk2(z, z0, z1);
  const uint64_t cdf = fma2(dup2(0.5f), erf_fast2(z, z0, z1), dup2(0.5f));
  upk2(mul2(mul2(x, x), dup2(-0.72134752044448170f)), q0, q1);
  const uint64_t pdf = mul2(pk2(ex2_approx(q0), ex2_approx(q1)), dup2(0.39894228040143268f));
  return fma2(x, pdf, cdf);
}
// gelu_erf_both on two values: identical arithmetic (every step is the same rounded fp32 operation), half the
// FMA-pipe instructions -- the GELU epilogues and conv0 are bound by instruction issue, not by memory
__device__ __forceinline__ void gelu_erf_both2(float x0, float x1, float& g0, float& g1, float& d0, float& d1) {
  const uint64_t x = pk2(x0, x1);
  const uint64_t z = mul2(x, dup2(0.70710678118654752f));
  float z0, z1;
  upk2(z, z0, z1);
  const uint64_t t = pk2(fabsf(z0), fabsf(z1));
  uint64_t p = fma2(dup2(-0.002944157226011157f), t, dup2(0.02959004044532776f));
  p = fma2(p, t, dup2(-0.1486656218767166f));
  p = fma2(p, t, dup2(-0.9185093641281128f));
  p = fma2(p, t, dup2(-1.6278890371322632f));
  p = mul2(p, t);
  const uint64_t q = mul2(mul2(x, x), dup2(-0.72134752044448170f));
  float p0, p1, q0, q1;
  upk2(p, p0, p1);
  upk2(q, q0, q1);
  const uint64_t e = pk2(copysignf(1.0f - ex2_approx(p0), z0), copysignf(1.0f - ex2_approx(p1), z1));
  const uint64_t hx = mul2(x, dup2(0.5f));
  const uint64_t g = fma2(hx, e, hx);
  const uint64_t cdf = fma2(e, dup2(0.5f), dup2(0.5f));
  const uint64_t pdf = mul2(pk2(ex2_approx(q0), ex2_approx(q1)), dup2(0.39894228040143268f));
  const uint64_t dg = fma2(x, pdf, cdf);
  upk2(g, g0, g1);
  upk2(dg, d0, d1);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  bf162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  bf162 v = *reinterpret_cast<bf162*>(&u);
  return __bfloat1622float2(v);
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ float2 ld_shared_v2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (an error the host sees) instead of a hung GPU.  No printf here: a call
// in the slow path makes ptxas spill every register that is live across the wait (measured: 64 spilled registers in the
// attention softmax loop).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// Wait for a barrier whose arrivals come from the PEER CTA of a pair (remote mbarrier.arrive, multicast tcgen05.commit):
// polls with test_wait at cluster scope.  try_wait suspends the thread in hardware and, measured on B200, a sleeping
// waiter is only woken promptly by arrivals from its own CTA -- remote arrivals cost thousands of cycles per hand-over.
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// ---- TMA ------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// pull one box of the tensor into L2 only (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_l2_2d(const void* tmap, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// TMA stores: shared memory (written through the generic proxy, so fence first) -> global, bulk-group completion
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
// global[box] += shared[box] (fp32 add performed at the L2, no load on the SM side)
__device__ __forceinline__ void tma_reduce_add_2d(const void* tmap, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// One lane of a CONVERGED warp (the lowest one).  tcgen05.mma / commit are single-thread instructions that execute on the
// uniform datapath: issued from a region the compiler sees as divergent (if (lane == 0) {...}) every one of them is wrapped
// in an ELECT / BRA.U.ANY loop with its descriptors rebuilt through a dependent chain of uniform ops -- ~130 cycles per
// MMA, as long as a 128x256x16 MMA takes to execute.  With the whole warp running the role loop and only the issue under
// elect_one(), ptxas emits the UTCHMMAs back to back with UIADD3.64 descriptor updates.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- tcgen05 / TMEM -------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {     // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed (implies fence::before).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (base_lane + t), columns [col, col+32).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory matrix descriptor (rows of 64 bf16 = 128 B; 8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address        bits [0,14)
  d |= (uint64_t)1 << 16;                               // leading byte offset  bits [16,30) (unused for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset   bits [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version 1 bits [46,48)
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B         bits [61,64)
  return d;
}
// MN-major, 128-byte-swizzled: 64 M|N elements per 128-byte row, 8 k-rows per 1024-byte atom; atoms along M|N are
// `lbo` bytes apart (here 8192: one 64x64 TMA box), atoms along K 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(8192 >> 4) << 16;                     // leading byte offset: next 64 M|N elements
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset: next 8 k-rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, M=128.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

#endif  // __CUDACC__
