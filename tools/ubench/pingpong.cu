// micro-benchmark: remote mbarrier arrive round trip between the two CTAs of a cluster (sm_100a)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity, bool test) {
  uint32_t ok;
  if (test) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
  else asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
  return ok;
}
template <bool TEST>
__global__ void __cluster_dims__(2, 1, 1) k(long long* out, int iters) {
  __shared__ uint64_t bar;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0) {
    uint32_t peer; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer) : "r"(s32(&bar)), "r"(rank ^ 1));
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (rank == 0) {
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(peer) : "memory");
        while (!try_wait(&bar, i & 1, TEST)) {}
      } else {
        while (!try_wait(&bar, i & 1, TEST)) {}
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(peer) : "memory");
      }
    }
    if (rank == 0) out[0] = (clock64() - t0) / iters;
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
int main() {
  long long* out; cudaMalloc(&out, 8);
  long long h;
  k<false><<<2, 32>>>(out, 1000); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("remote mbarrier round trip, try_wait (hw sleep): %lld cycles (%s)\n", h, cudaGetErrorString(cudaGetLastError()));
  k<true><<<2, 32>>>(out, 1000); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("remote mbarrier round trip, test_wait (poll):    %lld cycles (%s)\n", h, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
