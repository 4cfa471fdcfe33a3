#include <cstdio>
#include <cooperative_groups.h>
#include <cstdint>
namespace cg = cooperative_groups;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }
__global__ void __cluster_dims__(4,1,1) kd(double* o) {
  __shared__ __align__(16) double s[64];
  __shared__ __align__(8) uint64_t bar[2];
  auto c = cg::this_cluster();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  c.sync();
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar[0])), "r"(4u * 16u) : "memory");
  if (threadIdx.x < 4) {
    uint32_t ra = mapa(smem_u32(&s[2 * c.block_rank()]), threadIdx.x), rb = mapa(smem_u32(&bar[0]), threadIdx.x);
    double a = 1.0 + c.block_rank(), b = 2.0;
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" :: "r"(ra), "d"(a), "d"(b), "r"(rb) : "memory");
  }
  uint32_t done = 0; int spins = 0;
  while (!done && ++spins < (1 << 22)) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar[0])), "r"(0u) : "memory");
  }
  o[blockIdx.x * 64 + threadIdx.x % 64] = s[threadIdx.x % 64] + (done ? 0 : 1e9);
  c.sync();
}
int main() {
  double* o; cudaMalloc(&o, 8 * 64 * 4); cudaMemset(o, 0, 8*64*4);
  kd<<<4, 64>>>(o);
  cudaError_t e = cudaDeviceSynchronize();
  double h[256]; cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%s:", cudaGetErrorString(e));
  for (int b = 0; b < 4; ++b) { for (int i = 0; i < 8; ++i) printf(" %g", h[b * 64 + i]); printf(" |"); }
  printf("\n");
  return 0;
}
