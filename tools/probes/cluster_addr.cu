// Probe: how shared::cta addresses relate to shared::cluster addresses in a 4-CTA cluster (sm_100a).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void __cluster_dims__(4, 1, 1) probe() {
  __shared__ uint64_t bar;
  uint32_t rank, local = (uint32_t)__cvta_generic_to_shared(&bar);
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    uint32_t m[4];
    for (uint32_t r = 0; r < 4; ++r) asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(m[r]) : "r"(local), "r"(r));
    printf("block %d rank %u local 0x%08x masked 0x%08x mapa: 0x%08x 0x%08x 0x%08x 0x%08x\n", blockIdx.x, rank, local,
           local & 0xFEFFFFFFu, m[0], m[1], m[2], m[3]);
  }
}
int main() {
  probe<<<8, 32>>>();
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
