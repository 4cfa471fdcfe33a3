// Batched complex128 GEMM on the FP64 tensor pipe (DMMA m16n8k8):  C = opA(A) * opB(B), row-major.
// One CTA computes a 64x64 tile of one chain's C; 8 warps (4 x 2), each owning a 16x32 sub-tile as
// four 16x8 accumulator tiles (re + im).  Operands are staged through shared memory split into
// real/imaginary planes (k-major, row stride 68 doubles -> conflict-free fragment reads), next k-slab
// prefetched into registers while the current one is multiplied.  The complex product is four real
// DMMAs per (tile, k8):  Cr += Ar*Br + (-Ai)*Bi ;  Ci += Ar*Bi + Ai*Br.
#include "kbp_common.cuh"
#include "kbp_ops.cuh"

namespace kbp {

constexpr int BM = 64, BN = 64, BK = 16, LDS = 68;

struct GemmArgs {
  long long C, A, B;
  int m, n, k, opA, opB;
};

__device__ __forceinline__ cplx load_op(const cplx* __restrict__ M, int r, int c, int rows, int cols, int op) {
  // element (r, c) of op(M) where op(M) is rows x cols
  if (r >= rows || c >= cols) return cmake(0.0, 0.0);
  cplx v;
  if (op == OP_N || op == OP_J) v = M[(long long)r * cols + c];
  else v = M[(long long)c * rows + r];
  if (op == OP_C || op == OP_J) v.y = -v.y;
  return v;
}

__global__ void __launch_bounds__(256) zgemm_dmma_kernel(cplx* __restrict__ base, long long chain_stride, GemmArgs g) {
  __shared__ double As_re[BK][LDS], As_im[BK][LDS], Bs_re[BK][LDS], Bs_im[BK][LDS];
  cplx* Cb = base + (long long)blockIdx.z * chain_stride + g.C;
  const cplx* Ab = base + (long long)blockIdx.z * chain_stride + g.A;
  const cplx* Bb = base + (long long)blockIdx.z * chain_stride + g.B;
  const int row0 = blockIdx.y * BM, col0 = blockIdx.x * BN;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int wm = w & 3, wn = w >> 2;
  const int gq = lane >> 2, q = lane & 3;

  // per-thread staging coordinates (4 elements of each operand per k-slab)
  int ai[4], ak[4], bj[4], bk[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    if (g.opA == OP_N || g.opA == OP_J) { ak[r] = t & 15; ai[r] = (t >> 4) + 16 * r; }
    else { ai[r] = t & 63; ak[r] = (t >> 6) + 4 * r; }
    if (g.opB == OP_N || g.opB == OP_J) { bj[r] = t & 63; bk[r] = (t >> 6) + 4 * r; }
    else { bk[r] = t & 15; bj[r] = (t >> 4) + 16 * r; }
  }

  double cr[4][4], ci[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) cr[i][j] = ci[i][j] = 0.0;

  cplx ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      ra[r] = load_op(Ab, row0 + ai[r], k0 + ak[r], g.m, g.k, g.opA);
      rb[r] = load_op(Bb, k0 + bk[r], col0 + bj[r], g.k, g.n, g.opB);
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < g.k; k0 += BK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      As_re[ak[r]][ai[r]] = ra[r].x; As_im[ak[r]][ai[r]] = ra[r].y;
      Bs_re[bk[r]][bj[r]] = rb[r].x; Bs_im[bk[r]][bj[r]] = rb[r].y;
    }
    __syncthreads();
    if (k0 + BK < g.k) fetch(k0 + BK);
#pragma unroll
    for (int ks = 0; ks < BK / 8; ++ks) {
      double ar[4], aim[4], an[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int mm = wm * 16 + gq + 8 * (v & 1), kk = ks * 8 + q + 4 * (v >> 1);
        ar[v] = As_re[kk][mm];
        aim[v] = As_im[kk][mm];
        an[v] = -aim[v];
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        double br[2], bi[2];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const int kk = ks * 8 + q + 4 * v, nn = wn * 32 + nt * 8 + gq;
          br[v] = Bs_re[kk][nn];
          bi[v] = Bs_im[kk][nn];
        }
        dmma16x8x8(cr[nt], ar, br);
        dmma16x8x8(cr[nt], an, bi);
        dmma16x8x8(ci[nt], ar, bi);
        dmma16x8x8(ci[nt], aim, br);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int r = row0 + wm * 16 + gq + 8 * (v >> 1);
      const int c = col0 + wn * 32 + nt * 8 + 2 * q + (v & 1);
      if (r < g.m && c < g.n) Cb[(long long)r * g.n + c] = cmake(cr[nt][v], ci[nt][v]);
    }
}

void gemm(const Arena& a, int64_t C, int64_t A, int64_t B, int64_t m, int64_t n, int64_t k, int opA, int opB) {
  if (m == 0 || n == 0) return;
  GemmArgs g{C, A, B, (int)m, (int)n, (int)k, opA, opB};
  dim3 grid((unsigned)((n + BN - 1) / BN), (unsigned)((m + BM - 1) / BM), (unsigned)a.nb);
  zgemm_dmma_kernel<<<grid, 256, 0, a.stream>>>(a.base, a.chain_stride, g);
  ++*a.launches;
}

}  // namespace kbp
