// Bandwidth experiment for the fused Adam update: 4 read streams (P, G, M, V) + 4 write streams (P, M, V, bf16 copy)
// = 30 B/param.  Variants: cache hints, ILP, interleaved P/M/V layout, persistent grid.  nvcc -arch=sm_100a adam_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ float4 upd(float4 p, float4 g, float4& m, float4& v) {
  float* pp = &p.x; float* gg = &g.x; float* mm = &m.x; float* vv = &v.x;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    mm[t] = mm[t] + (gg[t] - mm[t]) * 0.1f;
    vv[t] = vv[t] * 0.999f + 0.001f * gg[t] * gg[t];
    pp[t] = pp[t] - 1e-3f * (mm[t] / (sqrtf(vv[t]) / 0.03f + 1e-8f));
  }
  return p;
}
__device__ __forceinline__ uint2 pk(float4 p) {
  __nv_bfloat162 a = __floats2bfloat162_rn(p.x, p.y), b = __floats2bfloat162_rn(p.z, p.w);
  return make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
}

template <int HINT>
__global__ void __launch_bounds__(256) k_base(float4* P, const float4* G, float4* M, float4* V, uint2* S, long long n4) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  float4 p = P[i];
  float4 g = HINT ? __ldcs(G + i) : G[i];
  float4 m = HINT ? __ldcs(M + i) : M[i], v = HINT ? __ldcs(V + i) : V[i];
  p = upd(p, g, m, v);
  if (HINT) { __stcs(M + i, m); __stcs(V + i, v); } else { M[i] = m; V[i] = v; }
  P[i] = p;
  S[i] = pk(p);
}
// two float4 per thread, loads first
__global__ void __launch_bounds__(256) k_ilp2(float4* P, const float4* G, float4* M, float4* V, uint2* S, long long n4) {
  const long long i = (long long)blockIdx.x * 512 + threadIdx.x;
  if (i + 256 >= n4) return;
  float4 p0 = P[i], p1 = P[i + 256], g0 = __ldcs(G + i), g1 = __ldcs(G + i + 256);
  float4 m0 = __ldcs(M + i), m1 = __ldcs(M + i + 256), v0 = __ldcs(V + i), v1 = __ldcs(V + i + 256);
  p0 = upd(p0, g0, m0, v0); p1 = upd(p1, g1, m1, v1);
  __stcs(M + i, m0); __stcs(M + i + 256, m1); __stcs(V + i, v0); __stcs(V + i + 256, v1);
  P[i] = p0; P[i + 256] = p1; S[i] = pk(p0); S[i + 256] = pk(p1);
}
// persistent grid-stride
__global__ void __launch_bounds__(256) k_persist(float4* P, const float4* G, float4* M, float4* V, uint2* S, long long n4) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 p = P[i], g = __ldcs(G + i), m = __ldcs(M + i), v = __ldcs(V + i);
    p = upd(p, g, m, v);
    __stcs(M + i, m); __stcs(V + i, v); P[i] = p; S[i] = pk(p);
  }
}
// interleaved state: PMV[i/32][3][32] float4 (a warp's P, M, V are three consecutive 512-byte runs)
__global__ void __launch_bounds__(256) k_inter(float4* PMV, const float4* G, uint2* S, long long n4) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  float4* b = PMV + (i >> 5) * 96 + (i & 31);
  float4 p = b[0], m = b[32], v = b[64], g = __ldcs(G + i);
  p = upd(p, g, m, v);
  b[0] = p; b[32] = m; b[64] = v; S[i] = pk(p);
}
__global__ void __launch_bounds__(256) k_copy(float4* D, const float4* A, long long n4) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i < n4) D[i] = A[i];
}
// read 4 / write 3 without the bf16 copy and without math
__global__ void __launch_bounds__(256) k_rw43(float4* P, const float4* G, float4* M, float4* V, long long n4) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  float4 p = P[i], g = G[i], m = M[i], v = V[i];
  p.x += g.x; m.x += g.y; v.x += g.z;
  P[i] = p; M[i] = m; V[i] = v;
}

int main() {
  const long long n = 4640000LL * 64, n4 = n / 4;
  float4 *P, *G, *M, *V, *PMV; uint2* S;
  CK(cudaMalloc(&P, n * 4)); CK(cudaMalloc(&G, n * 4)); CK(cudaMalloc(&M, n * 4)); CK(cudaMalloc(&V, n * 4));
  CK(cudaMalloc(&S, n * 2)); CK(cudaMalloc(&PMV, n * 12));
  CK(cudaMemset(P, 0, n * 4)); CK(cudaMemset(G, 0, n * 4)); CK(cudaMemset(M, 0, n * 4)); CK(cudaMemset(V, 0, n * 4)); CK(cudaMemset(PMV, 0, n * 12));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const unsigned gb = (unsigned)((n4 + 255) / 256);
  auto run = [&](const char* name, double bytes, auto&& f) {
    for (int i = 0; i < 2; ++i) f();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    printf("%-28s %7.3f ms  %7.1f GB/s  (%s)\n", name, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  run("copy 4B+4B", n * 8.0, [&] { k_copy<<<gb, 256>>>(M, V, n4); });
  run("rw 4 in / 3 out, no math", n * 28.0, [&] { k_rw43<<<gb, 256>>>(P, G, M, V, n4); });
  run("adam base no hints", n * 30.0, [&] { k_base<0><<<gb, 256>>>(P, G, M, V, S, n4); });
  run("adam base ldcs/stcs", n * 30.0, [&] { k_base<1><<<gb, 256>>>(P, G, M, V, S, n4); });
  run("adam ilp2", n * 30.0, [&] { k_ilp2<<<gb / 2, 256>>>(P, G, M, V, S, n4); });
  run("adam persistent 148x8", n * 30.0, [&] { k_persist<<<148 * 8, 256>>>(P, G, M, V, S, n4); });
  run("adam persistent 148x4", n * 30.0, [&] { k_persist<<<148 * 4, 256>>>(P, G, M, V, S, n4); });
  run("adam interleaved PMV", n * 30.0, [&] { k_inter<<<gb, 256>>>(PMV, G, S, n4); });
  return 0;
}
