// Cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) as a function of N, operands resident in shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_n mma_n.cu
#include "../../test-time-adaptation-asr-suta_b200/csrc/common.cuh"

template <int N>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    const uint64_t da = umma_desc_sw128(smem_u32(smem)), db = umma_desc_sw128(smem_u32(smem) + 16384);
    constexpr uint32_t idesc = umma_idesc_bf16(N);
    __shared__ long long t0s;
    if (elect_one()) {
      t0s = clock64();
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) umma_bf16_ss(tm + (i & 1) * 256, da + 2 * k2, db + 2 * k2, idesc, 1u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    __syncwarp();
    if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0s;
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tm);
}

template <int N>
void run(long long* d, int grid) {
  const int iters = 2000;
  cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  k<N><<<grid, 128, 64 * 1024>>>(d, iters);
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d grid=%3d: %.1f cycles per MMA (M=128, K=16)   [%s]\n", N, grid, (double)h / (iters * 4), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  for (int grid : {1, 148}) {
    run<16>(d, grid); run<32>(d, grid); run<48>(d, grid); run<64>(d, grid); run<96>(d, grid); run<128>(d, grid); run<192>(d, grid); run<256>(d, grid);
  }
  return 0;
}
