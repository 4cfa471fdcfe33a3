// FP64 latency / throughput probe (DFMA, DADD, MUFU.RSQ64H-based rsqrt, drcp, SHFL of a double) on one SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat_kernel(double* out, long long* cyc, double x0, int iters) {
  double x = x0 + threadIdx.x * 1e-9, y = 1.0000001;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) x = fma(x, y, 1e-9);          // dependent chain
  }
  long long t1 = clock64();
  double z = x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) z = z + y;
  }
  long long t2 = clock64();
  double r = fabs(z) + 1.5;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k) r = rsqrt(r) + 1.5;
  }
  long long t3 = clock64();
  double c = r;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k) c = __drcp_rn(c) + 1.5;
  }
  long long t4 = clock64();
  double s = c;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) s = __shfl_xor_sync(0xffffffffu, s, 1) + 1.0;
  }
  long long t5 = clock64();
  float f = (float)s;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) f = fmaf(f, 1.0000001f, 1e-9f);
  }
  long long t6 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + z + r + c + s + f;
}

// throughput: many independent chains per thread, many warps
__global__ void thr_kernel(double* out, long long* cyc, int iters) {
  double a[8];
  for (int k = 0; k < 8; ++k) a[k] = 1.0 + threadIdx.x * 1e-9 + k;
  const double y = 1.0000001;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fma(a[k], y, 1e-9);
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  double s = 0;
  for (int k = 0; k < 8; ++k) s += a[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, sizeof(double) * 1024 * 256);
  cudaMallocManaged(&cyc, sizeof(long long) * 8);
  const int iters = 1000;
  lat_kernel<<<1, 32>>>(out, cyc, 1.0, iters);
  cudaDeviceSynchronize();
  printf("latency (1 warp, dependent): DFMA %.1f  DADD %.1f  rsqrt(double)+DADD %.1f  __drcp_rn+DADD %.1f  SHFL.f64+DADD %.1f  FFMA %.1f cycles\n",
         cyc[0] / (16.0 * iters), cyc[1] / (16.0 * iters), cyc[2] / (4.0 * iters), cyc[3] / (4.0 * iters), cyc[4] / (8.0 * iters), cyc[5] / (16.0 * iters));
  for (int threads = 32; threads <= 1024; threads *= 2) {
    thr_kernel<<<1, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    printf("throughput %4d threads x 8 chains: %.2f DFMA/clk/SM\n", threads, 8.0 * iters * threads / (double)cyc[0]);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
