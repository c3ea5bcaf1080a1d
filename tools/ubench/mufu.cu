// micro-benchmark: MUFU.EX2 / FFMA issue rates per SM sub-partition on sm_100a (cycles per warp instruction)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) a[i] = ex2a(a[i]);
      if (MODE == 1) a[i] = fmaf(a[i], 1.0001f, 0.5f);
      if (MODE == 2) a[i] = ex2a(fmaf(a[i], 0.999f, -0.1f));
      if (MODE == 3) { a[i] = ex2a(fmaf(a[i], 0.999f, -0.1f)); a[(i + 1) & 15] += a[i]; }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 1000;
  const char* names[4] = {"MUFU.EX2", "FFMA", "FFMA+EX2", "FFMA+EX2+FADD"};
  for (int mode = 0; mode < 4; ++mode)
    for (int warps : {4, 8, 16, 32}) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<1, warps * 32>>>(out, cyc, iters);
        if (mode == 1) k<1><<<1, warps * 32>>>(out, cyc, iters);
        if (mode == 2) k<2><<<1, warps * 32>>>(out, cyc, iters);
        if (mode == 3) k<3><<<1, warps * 32>>>(out, cyc, iters);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      double per = (double)h / (iters * 16.0);      // cycles per (one instruction-group of every warp)
      printf("%-14s warps/SM %2d (per SMSP %d): %.2f cycles per element-step per warp, %.2f per SMSP-warp-instr\n", names[mode], warps,
             warps / 4, per, per / (warps / 4));
    }
  return 0;
}
