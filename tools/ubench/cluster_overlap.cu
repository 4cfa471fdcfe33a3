// Do thread-block-cluster kernels launched from different program graphs run side by side?
// k graphs, each a chain of N kernels; kernel = one cluster of C CTAs (256 threads, `smem` bytes of dynamic shared memory)
// spinning `spin` clocks, with a cluster barrier at both ends.   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
namespace cg = cooperative_groups;

__global__ void cluster_spin_kernel(long long spin, double* out) {
  extern __shared__ double sm[];
  cg::cluster_group cl = cg::this_cluster();
  cl.sync();
  const long long t0 = clock64();
  double x = 0.0;
  while (clock64() - t0 < spin) x += 1.0;
  sm[threadIdx.x] = x;
  cl.sync();
  if (x < 0) out[0] = sm[0];
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 200;
  double* out;
  cudaMalloc(&out, 8);
  cudaFuncSetAttribute(cluster_spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int ks[] = {1, 2, 3, 6, 12};
  const int Cs[] = {1, 4, 8};
  const size_t smems[] = {30 * 1024, 140 * 1024};
  const long long spin = 200000;                             // ~100 us
  for (size_t smem : smems)
    for (int C : Cs)
      for (int k : ks) {
        std::vector<cudaStream_t> st(k);
        std::vector<cudaGraphExec_t> ex(k);
        bool ok = true;
        for (int i = 0; i < k; ++i) {
          cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
          cudaGraph_t g;
          cudaStreamBeginCapture(st[i], cudaStreamCaptureModeThreadLocal);
          for (int n = 0; n < N; ++n) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(C);
            cfg.blockDim = dim3(256);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = st[i];
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            if (cudaLaunchKernelEx(&cfg, cluster_spin_kernel, spin, out) != cudaSuccess) ok = false;
          }
          cudaStreamEndCapture(st[i], &g);
          if (cudaGraphInstantiate(&ex[i], g, 0) != cudaSuccess) ok = false;
          cudaGraphDestroy(g);
        }
        if (!ok) { printf("smem %zu C %d k %d: launch failed (%s)\n", smem, C, k, cudaGetErrorString(cudaGetLastError())); continue; }
        for (int i = 0; i < k; ++i) cudaGraphLaunch(ex[i], st[i]);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, st[0]);
        for (int i = 1; i < k; ++i) cudaStreamWaitEvent(st[i], e0, 0);
        for (int i = 0; i < k; ++i) cudaGraphLaunch(ex[i], st[i]);
        std::vector<cudaEvent_t> done(k);
        for (int i = 1; i < k; ++i) { cudaEventCreate(&done[i]); cudaEventRecord(done[i], st[i]); cudaStreamWaitEvent(st[0], done[i], 0); }
        cudaEventRecord(e1, st[0]);
        cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("smem %3zu KB  cluster %d  k=%2d graphs x %d kernels of ~100 us: %8.2f ms  (%.2f x one graph's ideal %.1f ms)\n", smem / 1024, C, k, N, ms,
               ms / (N * 0.105), N * 0.105);
        for (int i = 0; i < k; ++i) { cudaGraphExecDestroy(ex[i]); cudaStreamDestroy(st[i]); }
      }
  return 0;
}
