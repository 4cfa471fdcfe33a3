// Batched complex128 GEMM on the FP64 tensor pipe (DMMA m16n8k8):  C = opA(A) * opB(B), row-major.
//
// The boundary-MPS chain is a long sequence of SMALL dependent products (m, n of a few hundred at most,
// k up to a few thousand), so what bounds a product is latency, not the tensor pipe: one CTA computes a
// BM x BN tile with 8 warps and streams its k-slabs through a 4-stage cp.async ring in shared memory
// (16-byte copies of whole complex128 elements, zero-filled past the edges), one barrier per slab.
// Tiles are 64x64 (warps 4 x 2, four 16x8 accumulators each) when that already fills the machine, else
// 32x32 (warps 2 x 4); products with a long k and a tiny output (Gram matrices Y^H Y) can additionally be
// split along k into `ksplit` partial results written side by side (summed by the consumer: deterministic,
// no atomics).  Operands stay in the layout they have in HBM -- "k-contiguous" [row][k] or
// "m-contiguous" [k][row] depending on op -- so every copy is coalesced and transposition / conjugation
// happen in the fragment loads: one LDS.128 brings (re, im) of an element, conflict-free for both layouts
// (row strides of 20 resp. BM+2 elements).  The complex product is four real DMMAs per (tile, k8):
//     Cr += Ar*Br - Ai*Bi ;  Ci += Ar*Bi + Ai*Br.
#include "kbp_common.cuh"
#include "kbp_ops.cuh"

namespace kbp {

constexpr int BK = 16;
constexpr int LDK = BK + 4;       // row stride (elements) of a k-contiguous tile

struct GemmArgs {
  long long C, A, B;
  int m, n, k, opA, opB;
  int ksplit;                     // partial results: C + s * m * n, s < ksplit
  int rows_on_x;                  // grid.x (2^31 limit) carries the row tiles instead of the column tiles
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int BM, int BN, int STAGES>
__global__ void __launch_bounds__(256) zgemm_dmma_kernel(cplx* __restrict__ base, long long chain_stride, GemmArgs g) {
  constexpr int WMS = BM / 16, WNS = 8 / WMS;    // warp grid
  constexpr int NT = BN / WNS / 8;               // 16x8 accumulator tiles per warp
  constexpr int TA = (BM * LDK > BK * (BM + 2)) ? BM * LDK : BK * (BM + 2);   // elements per A tile (either layout)
  constexpr int TB = (BN * LDK > BK * (BN + 2)) ? BN * LDK : BK * (BN + 2);
  constexpr int LDM = BM + 2, LDN = BN + 2;
  extern __shared__ __align__(16) unsigned char gemm_smem[];
  cplx* As = reinterpret_cast<cplx*>(gemm_smem);                 // [STAGES][TA]
  cplx* Bs = As + STAGES * TA;                                   // [STAGES][TB]

  const int chain = blockIdx.z / g.ksplit, split = blockIdx.z - chain * g.ksplit;
  cplx* Cb = base + (long long)chain * chain_stride + g.C + (long long)split * g.m * g.n;
  const cplx* Ab = base + (long long)chain * chain_stride + g.A;
  const cplx* Bb = base + (long long)chain * chain_stride + g.B;
  const int row0 = (g.rows_on_x ? blockIdx.x : blockIdx.y) * BM, col0 = (g.rows_on_x ? blockIdx.y : blockIdx.x) * BN;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int wm = w % WMS, wn = w / WMS;
  const int gq = lane >> 2, q = lane & 3;
  const bool a_kc = (g.opA == OP_N || g.opA == OP_J);   // A stored [m][k]
  const bool b_kc = (g.opB == OP_T || g.opB == OP_C);   // B stored [n][k]
  const double sa = (g.opA == OP_C || g.opA == OP_J) ? -1.0 : 1.0;
  const double sb = (g.opB == OP_C || g.opB == OP_J) ? -1.0 : 1.0;

  const int slabs_total = (g.k + BK - 1) / BK;
  const int per = (slabs_total + g.ksplit - 1) / g.ksplit;
  const int s_begin = split * per;
  int s_end = s_begin + per;
  if (s_end > slabs_total) s_end = slabs_total;
  const int nslab = s_end > s_begin ? s_end - s_begin : 0;

  auto issue = [&](int slab, int stage) {
    const int k0 = (s_begin + slab) * BK;
    cplx* at = As + stage * TA;
    cplx* bt = Bs + stage * TB;
#pragma unroll
    for (int r = 0; r < BM * BK / 256; ++r) {
      const int c = t + 256 * r;
      if (a_kc) {
        const int mm = c / BK, kk = c % BK;
        const bool ok = row0 + mm < g.m && k0 + kk < g.k;
        cp_async16(at + mm * LDK + kk, ok ? Ab + (long long)(row0 + mm) * g.k + k0 + kk : Ab, ok);
      } else {
        const int kk = c / BM, mm = c % BM;
        const bool ok = row0 + mm < g.m && k0 + kk < g.k;
        cp_async16(at + kk * LDM + mm, ok ? Ab + (long long)(k0 + kk) * g.m + row0 + mm : Ab, ok);
      }
    }
#pragma unroll
    for (int r = 0; r < BN * BK / 256; ++r) {
      const int c = t + 256 * r;
      if (b_kc) {
        const int nn = c / BK, kk = c % BK;
        const bool ok = col0 + nn < g.n && k0 + kk < g.k;
        cp_async16(bt + nn * LDK + kk, ok ? Bb + (long long)(col0 + nn) * g.k + k0 + kk : Bb, ok);
      } else {
        const int kk = c / BN, nn = c % BN;
        const bool ok = col0 + nn < g.n && k0 + kk < g.k;
        cp_async16(bt + kk * LDN + nn, ok ? Bb + (long long)(k0 + kk) * g.n + col0 + nn : Bb, ok);
      }
    }
  };

  double cr[NT][4], ci[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) cr[i][j] = ci[i][j] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nslab) issue(s, s);
    cp_async_commit();
  }
  for (int slab = 0; slab < nslab; ++slab) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();                                   // slab's data visible; everyone finished slab-1's stage
    if (slab + STAGES - 1 < nslab) issue(slab + STAGES - 1, (slab + STAGES - 1) % STAGES);
    cp_async_commit();
    const cplx* at = As + (slab % STAGES) * TA;
    const cplx* bt = Bs + (slab % STAGES) * TB;
#pragma unroll
    for (int ks = 0; ks < BK / 8; ++ks) {
      double ar[4], aim[4], an[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int mm = wm * 16 + gq + 8 * (v & 1), kk = ks * 8 + q + 4 * (v >> 1);
        const cplx x = a_kc ? at[mm * LDK + kk] : at[kk * LDM + mm];
        ar[v] = x.x;
        aim[v] = sa * x.y;
        an[v] = -aim[v];
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        double br[2], bi[2];
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const int kk = ks * 8 + q + 4 * v, nn = wn * (NT * 8) + nt * 8 + gq;
          const cplx x = b_kc ? bt[nn * LDK + kk] : bt[kk * LDN + nn];
          br[v] = x.x;
          bi[v] = sb * x.y;
        }
        dmma16x8x8(cr[nt], ar, br);
        dmma16x8x8(cr[nt], an, bi);
        dmma16x8x8(ci[nt], ar, bi);
        dmma16x8x8(ci[nt], aim, br);
      }
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int r = row0 + wm * 16 + gq + 8 * (v >> 1);
      const int c = col0 + wn * (NT * 8) + nt * 8 + 2 * q + (v & 1);
      if (r < g.m && c < g.n) Cb[(long long)r * g.n + c] = cmake(cr[nt][v], ci[nt][v]);
    }
}

template <int BM, int BN, int STAGES>
static void launch_gemm(const Arena& a, GemmArgs g) {
  constexpr int TA = (BM * LDK > BK * (BM + 2)) ? BM * LDK : BK * (BM + 2);
  constexpr int TB = (BN * LDK > BK * (BN + 2)) ? BN * LDK : BK * (BN + 2);
  constexpr size_t smem = sizeof(double2) * (size_t)STAGES * (TA + TB);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(zgemm_dmma_kernel<BM, BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  const unsigned tn = (unsigned)((g.n + BN - 1) / BN), tm = (unsigned)((g.m + BM - 1) / BM);
  g.rows_on_x = tm > tn;
  dim3 grid(g.rows_on_x ? tm : tn, g.rows_on_x ? tn : tm, (unsigned)(a.nb * g.ksplit));
  zgemm_dmma_kernel<BM, BN, STAGES><<<grid, 256, smem, a.stream>>>(a.base, a.chain_stride, g);
  ++*a.launches;
}

void gemm_splitk(const Arena& a, int64_t C, int64_t A, int64_t B, int64_t m, int64_t n, int64_t k, int opA, int opB, int ksplit) {
  if (m == 0 || n == 0) return;
  GemmArgs g{C, A, B, (int)m, (int)n, (int)k, opA, opB, ksplit < 1 ? 1 : ksplit, 0};
  const int64_t ctas64 = ((n + 63) / 64) * ((m + 63) / 64) * a.nb * g.ksplit;
  if (ctas64 >= 96) launch_gemm<64, 64, 3>(a, g);
  else launch_gemm<32, 32, 4>(a, g);
}

void gemm(const Arena& a, int64_t C, int64_t A, int64_t B, int64_t m, int64_t n, int64_t k, int opA, int opB) {
  gemm_splitk(a, C, A, B, m, n, k, opA, opB, 1);
}

}  // namespace kbp
