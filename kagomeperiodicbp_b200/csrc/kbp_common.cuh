// Shared device helpers for the Kagome block-BP kernels (sm_100a, complex128 throughout).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kbp {

typedef double2 cplx;   // interleaved (re, im) -- the layout numpy complex128 uses

__host__ __device__ __forceinline__ cplx cmake(double r, double i) { return make_double2(r, i); }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) {
  return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
// a * conj(b)
__device__ __forceinline__ cplx cmulc(cplx a, cplx b) {
  return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y));
}
// conj(a) * b
__device__ __forceinline__ cplx ccmul(cplx a, cplx b) {
  return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.x, b.y, -a.y * b.x));
}
__device__ __forceinline__ cplx cconj(cplx a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ cplx cscale(cplx a, double s) { return make_double2(a.x * s, a.y * s); }
__device__ __forceinline__ double cabs2(cplx a) { return fma(a.x, a.x, a.y * a.y); }
__device__ __forceinline__ cplx cfma(cplx a, cplx b, cplx c) {  // a*b + c
  return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ cplx warp_sum(cplx v) {
  v.x = warp_sum(v.x);
  v.y = warp_sum(v.y);
  return v;
}

// FP64 tensor-core tile: D(16x8) += A(16x8, row) * B(8x8, col).  Fragment layout (lane = 4*g + q):
//   a[v0 + 2*v1] = A[g + 8*v0][q + 4*v1]      b[v] = B[k = q + 4*v][n = g]
//   c[v0 + 2*v1] = C[g + 8*v1][2*q + v0]
// SASS: DMMA.  (sm_90+ shape; tcgen05 has no FP64 kind, so DMMA is the FP64 tensor path on sm_100a.)
__device__ __forceinline__ void dmma16x8x8(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// block-wide sum of a double; every thread gets the result.  `red` = >= 33 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = lane < nw ? red[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

}  // namespace kbp
