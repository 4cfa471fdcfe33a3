// How many dependent kernel launches per second does the GPU sustain when k program graphs run side by side?
// Each graph is a chain of N kernels (one CTA of 64 threads spinning `spin` clocks); k graphs are launched on k streams.
// If the aggregate node rate saturates as k grows, concurrently running side programs are bound by launch processing,
// not by kernel work.   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o launch_rate launch_rate.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__global__ void spin_kernel(long long spin, int ctas_dummy, double* out) {
  const long long t0 = clock64();
  double x = 0.0;
  while (clock64() - t0 < spin) x += 1.0;
  if (x < 0) out[0] = x;
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 4000;
  double* out;
  cudaMalloc(&out, 8);
  const int ks[] = {1, 2, 3, 6, 12};
  const long long spins[] = {0, 4000, 20000};             // ~0, 2, 10 us at 1.9 GHz
  const int grids[] = {1, 148, 512};
  for (long long spin : spins)
    for (int grid : grids) {
      for (int k : ks) {
        std::vector<cudaStream_t> st(k);
        std::vector<cudaGraphExec_t> ex(k);
        for (int i = 0; i < k; ++i) {
          cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
          cudaGraph_t g;
          cudaStreamBeginCapture(st[i], cudaStreamCaptureModeThreadLocal);
          for (int n = 0; n < N; ++n) spin_kernel<<<grid, 64, 0, st[i]>>>(spin, grid, out);
          cudaStreamEndCapture(st[i], &g);
          cudaGraphInstantiate(&ex[i], g, 0);
          cudaGraphDestroy(g);
        }
        for (int i = 0; i < k; ++i) cudaGraphLaunch(ex[i], st[i]);     // warm
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
        printf("spin %6lld clk  grid %3d  k=%2d graphs x %d nodes: %8.2f ms  -> %6.2f us per node per graph, aggregate %7.1f k nodes/s\n", spin, grid, k, N, ms,
               1e3 * ms / N, (double)k * N / ms);
        for (int i = 0; i < k; ++i) { cudaGraphExecDestroy(ex[i]); cudaStreamDestroy(st[i]); }
      }
    }
  return 0;
}
