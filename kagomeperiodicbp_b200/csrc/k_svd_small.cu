// Truncated SVD of a SMALL complex128 matrix entirely in one CTA's shared memory: one-sided (Hestenes)
// Jacobi on the rows of X = A (m <= n) or X = A^T (m > n), p = min(m, n) rows of length q = max(m, n),
// p*q*16 B <= ~220 KB.  This is the kernel behind every truncation at D = 2 and D = 3 (matrices up to
// 81 x 162) and behind the Rayleigh-Ritz stage of the subspace-iteration SVD that larger D uses
// (k_tsvd.cu) -- it replaces numpy.linalg.svd in mps.right_canonical (src/libs/bmpslib.py:733-772).
//
// One launch does the whole factorisation: load, sweeps until every pair's |<x_i,x_j>| / (|x_i||x_j|)
// is below tol (the convergence test stays on chip: no host round trip), rank sort, outputs.  A sweep is a
// round-robin tournament of pp-1 steps; in each step the pp/2 disjoint row pairs are rotated concurrently,
// one lane group (8/16/32 lanes, picked so that all pairs fit the 1024 threads) per pair: the group
// forms the 2x2 Gram matrix with shuffle reductions, builds the rotation in full double precision and
// applies it, writing the larger row first (de Rijk ordering, which also sorts the rows as a side effect).
//
// Outputs (same contract as the block-Jacobi kernel in k_svd.cu):
//   m <= n:  rows converge to s_i v_i^H ->  Vh_k = rows / s  (orthonormal to rounding),  US = A Vh_k^H
//   m >  n:  the rows carry an identity block behind them, [A^T | I], rotated along (Gram sums over the A^T part only):
//            rows converge to [s_i u_i^T | conj(v_i)^T]  ->  US_k = first part, Vh_k = conj(second part), both exact
// so US Vh is exactly the projection of A on the span found, whatever the rounding of the other factor.
#include "kbp_common.cuh"
#include "kbp_ops.cuh"

#include <math.h>
#include <stdio.h>

namespace kbp {

constexpr double SMALL_TOL = 1e-14;
constexpr int SMALL_MAX_SWEEPS = 60;
constexpr size_t SMALL_SMEM_MAX = 225 * 1024;

__device__ __forceinline__ void rr_pair_s(int n, int r, int t, int& a, int& b) {
  const int m = n - 1;
  if (t == 0) { a = m; b = r; }
  else { a = (r + t) % m; b = (r - t + m) % m; }
  if (a > b) { int x = a; a = b; b = x; }
}

struct SmallArgs {
  long long A, US, Vh;      // arena offsets; A is m x n with row stride lda
  int m, n, lda, keep;
  int nr_bulk, slot_lognorm, slot_trunc;
  int group;                // lanes per row pair
};

size_t svd_small_smem(int64_t m, int64_t n) {
  const int64_t p = m < n ? m : n, q = m <= n ? n : m + n;       // tall: rows of [A^T | I]
  return (size_t)(p * q) * sizeof(double2) + (size_t)p * (sizeof(double) + sizeof(int)) + 64;
}

bool svd_small_fits(int64_t m, int64_t n) {
  const int64_t p = m < n ? m : n;
  return p <= 128 && svd_small_smem(m, n) <= SMALL_SMEM_MAX;
}

// rotate one row pair with both rows held in registers between the Gram sums and the update (CPL columns per lane)
template <int CPL>
__device__ __forceinline__ void jacobi_pair_cached(cplx* __restrict__ xi, cplx* __restrict__ xj, int q, int qx, int G, int gl, unsigned gmask,
                                                   double floor2, double& my_off) {
  cplx u[CPL], v[CPL];
  double a = 0.0, b = 0.0, cr = 0.0, ci = 0.0;
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    const int c = gl + k * G;
    u[k] = c < q ? xi[c] : cmake(0.0, 0.0);
    v[k] = c < q ? xj[c] : cmake(0.0, 0.0);
    if (c < qx) {                                  // Gram sums over the data columns only (not the accumulated rotations)
      a = fma(u[k].x, u[k].x, fma(u[k].y, u[k].y, a));
      b = fma(v[k].x, v[k].x, fma(v[k].y, v[k].y, b));
      cr = fma(u[k].x, v[k].x, fma(u[k].y, v[k].y, cr));
      ci = fma(u[k].y, v[k].x, fma(-u[k].x, v[k].y, ci));
    }
  }
  for (int o = G >> 1; o > 0; o >>= 1) {
    a += __shfl_xor_sync(gmask, a, o);
    b += __shfl_xor_sync(gmask, b, o);
    cr += __shfl_xor_sync(gmask, cr, o);
    ci += __shfl_xor_sync(gmask, ci, o);
  }
  const double r2 = fma(cr, cr, ci * ci);
  if (!(a > floor2 && b > floor2 && r2 > 0.0)) return;
  // the pair counts as unconverged iff |c|^2 > tol^2 a b (no division / square root on the decision path); the rotation
  // needs three dependent special-function evaluations: rsqrt(|c|^2) || rsqrt(d^2 + 4|c|^2), a reciprocal, rsqrt(1 + t^2)
  if (!(r2 > (SMALL_TOL * SMALL_TOL) * a * b)) return;
  my_off = fmax(my_off, r2 * __drcp_rn(a * b));        // cos^2 of the pair's angle (off the rotation's critical path)
  const double inv = rsqrt(r2), ab = r2 * inv;
  const double er = cr * inv, ei = ci * inv;
  const double d = a - b;
  const double h = fma(d, d, 4.0 * r2);
  double tt = 2.0 * ab * __drcp_rn(fabs(d) + h * rsqrt(h));
  if (d < 0.0) tt = -tt;
  const double cs = rsqrt(fma(tt, tt, 1.0)), sn = tt * cs;
  const bool swap = d < 0.0;
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    const int c = gl + k * G;
    if (c < q) {
      const double vr = er * v[k].x - ei * v[k].y, vi = er * v[k].y + ei * v[k].x;
      const cplx yi = make_double2(fma(cs, u[k].x, sn * vr), fma(cs, u[k].y, sn * vi));
      const cplx yj = make_double2(fma(cs, vr, -sn * u[k].x), fma(cs, vi, -sn * u[k].y));
      xi[c] = swap ? yj : yi;
      xj[c] = swap ? yi : yj;
    }
  }
}

template <bool CACHED>
__global__ void __launch_bounds__(CACHED ? 768 : 1024) svd_small_kernel(cplx* __restrict__ base, long long chain_stride, double* __restrict__ slots,
                                                                        int n_slots, SmallArgs g) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int m = g.m, n = g.n;
  const bool mode_t = m > n;
  const int p = mode_t ? n : m, qx = mode_t ? m : n, q = mode_t ? m + n : n;   // q: row length incl. the identity block of a tall matrix
  cplx* X = reinterpret_cast<cplx*>(sm_raw);                                 // p x q
  double* s2 = reinterpret_cast<double*>(sm_raw + sizeof(cplx) * (size_t)p * q);   // p
  int* idx = reinterpret_cast<int*>(s2 + p);                                 // p
  __shared__ double red[34];
  __shared__ double sh_flag;

  cplx* cb = base + (long long)blockIdx.x * chain_stride;
  const cplx* A = cb + g.A;
  cplx* US = cb + g.US;
  cplx* Vh = cb + g.Vh;
  const int t = threadIdx.x, nt = blockDim.x, lane = t & 31, w = t >> 5, nw = nt >> 5;

  // ---- load (transposing when m > n) and ||A||_F^2
  double fro = 0.0;
  if (!mode_t) {
    for (int e = t; e < p * q; e += nt) {
      const int r = e / q, c = e - r * q;
      const cplx v = A[(long long)r * g.lda + c];
      X[e] = v;
      fro += cabs2(v);
    }
  } else {
    for (int e = t; e < m * n; e += nt) {           // coalesced read of A[r][c], scattered smem write
      const int r = e / n, c = e - r * n;
      const cplx v = A[(long long)r * g.lda + c];
      X[c * q + r] = v;
      fro += cabs2(v);
    }
    for (int e = t; e < n * n; e += nt) X[(e / n) * q + m + e % n] = cmake(e / n == e % n ? 1.0 : 0.0, 0.0);
  }
  const double fro2 = block_sum(fro, red);          // (contains the barrier that publishes X)
  const double floor2 = 1e-34 * fro2;

  // ---- Jacobi sweeps
  const int pp = (p + 1) & ~1, npairs = pp / 2, G = g.group;
  const int slot = t / G, gl = t - slot * G;
  // groups of one warp can take different branches (phantom row, dead rows): shuffles name their own group only
  const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));
  bool converged = p < 2;
#ifdef KBP_SMALL_DEBUG
  int nsweeps = 0;
  long long tstart = clock64();
#endif
  for (int sweep = 0; sweep < SMALL_MAX_SWEEPS && !converged; ++sweep) {
#ifdef KBP_SMALL_DEBUG
    ++nsweeps;
#endif
    double my_off = 0.0;
    for (int step = 0; step < pp - 1; ++step) {
      if (slot < npairs) {
        int i, j;
        rr_pair_s(pp, step, slot, i, j);
        if (CACHED && j < p) {
          cplx* xi = X + i * q;
          cplx* xj = X + j * q;
          switch ((q + G - 1) / G) {
            case 1: jacobi_pair_cached<1>(xi, xj, q, qx, G, gl, gmask, floor2, my_off); break;
            case 2: jacobi_pair_cached<2>(xi, xj, q, qx, G, gl, gmask, floor2, my_off); break;
            case 3: jacobi_pair_cached<3>(xi, xj, q, qx, G, gl, gmask, floor2, my_off); break;
            case 4: jacobi_pair_cached<4>(xi, xj, q, qx, G, gl, gmask, floor2, my_off); break;
            case 5: jacobi_pair_cached<5>(xi, xj, q, qx, G, gl, gmask, floor2, my_off); break;
            default: jacobi_pair_cached<6>(xi, xj, q, qx, G, gl, gmask, floor2, my_off); break;
          }
        } else if (j < p) {
          cplx* xi = X + i * q;
          cplx* xj = X + j * q;
          double a = 0.0, b = 0.0, cr = 0.0, ci = 0.0;
          for (int c = gl; c < qx; c += G) {
            const cplx u = xi[c], v = xj[c];
            a = fma(u.x, u.x, fma(u.y, u.y, a));
            b = fma(v.x, v.x, fma(v.y, v.y, b));
            cr = fma(u.x, v.x, fma(u.y, v.y, cr));          // u conj(v)
            ci = fma(u.y, v.x, fma(-u.x, v.y, ci));
          }
          for (int o = G >> 1; o > 0; o >>= 1) {
            a += __shfl_xor_sync(gmask, a, o);
            b += __shfl_xor_sync(gmask, b, o);
            cr += __shfl_xor_sync(gmask, cr, o);
            ci += __shfl_xor_sync(gmask, ci, o);
          }
          const double r2 = fma(cr, cr, ci * ci);
          if (a > floor2 && b > floor2 && r2 > 0.0) {
            if (r2 > (SMALL_TOL * SMALL_TOL) * a * b) {
              my_off = fmax(my_off, r2 * __drcp_rn(a * b));
              // x_j' = e x_j with e = c/|c| makes <x_i, x_j'> = |c| real; then a real rotation by theta,
              // tan(2 theta) = 2|c| / (a - b), small-angle root
              const double inv = rsqrt(r2), ab = r2 * inv;
              const double er = cr * inv, ei = ci * inv;
              const double d = a - b;
              const double h = fma(d, d, 4.0 * r2);
              double tt = 2.0 * ab * __drcp_rn(fabs(d) + h * rsqrt(h));
              if (d < 0.0) tt = -tt;
              const double cs = rsqrt(fma(tt, tt, 1.0)), sn = tt * cs;
              const bool swap = d < 0.0;            // keep the larger row first
              for (int c = gl; c < q; c += G) {
                const cplx u = xi[c], v = xj[c];
                const double vr = er * v.x - ei * v.y, vi = er * v.y + ei * v.x;     // e x_j
                const cplx yi = make_double2(fma(cs, u.x, sn * vr), fma(cs, u.y, sn * vi));
                const cplx yj = make_double2(fma(cs, vr, -sn * u.x), fma(cs, vi, -sn * u.y));
                xi[c] = swap ? yj : yi;
                xj[c] = swap ? yi : yj;
              }
            }
          }
        }
      }
      __syncthreads();
    }
    // sweep-wide maximum of the pair measure
    my_off = warp_max(my_off);
    if (lane == 0) red[w] = my_off;
    __syncthreads();
    if (w == 0) {
      double x = lane < nw ? red[lane] : 0.0;
      x = warp_max(x);
      if (lane == 0) sh_flag = x;
    }
    __syncthreads();
    // my_off = largest cos^2 rotated away in this sweep (0: nothing was above tol).  One-sided Jacobi converges quadratically
    // at the end: once every pair of a sweep started below sqrt(tol), the next sweep would find nothing to do -- skip it
    converged = sh_flag <= SMALL_TOL;
#ifdef KBP_SMALL_DEBUG
    if (t == 0) printf("[small %dx%d] sweep %d max off %.3e  cycles %lld\n", p, q, sweep, sh_flag, clock64() - tstart);
#endif
  }

  // ---- singular values = row norms, rank sort (descending, stable)
  for (int i = w; i < p; i += nw) {
    double acc = 0.0;
    const cplx* row = X + i * q;
    for (int c = lane; c < qx; c += 32) acc += cabs2(row[c]);
    acc = warp_sum(acc);
    if (lane == 0) s2[i] = (acc == acc && acc < 1e300) ? acc : 0.0;
  }
  for (int k = t; k < p; k += nt) idx[k] = k;       // identity first: non-finite input can never index out of bounds
  __syncthreads();
  double disc_part = 0.0;
  for (int i = t; i < p; i += nt) {
    const double si = s2[i];
    int rank = 0;
    for (int j = 0; j < p; ++j) {
      const double sj = s2[j];
      rank += (sj > si) || (sj == si && j < i);
    }
    if (rank >= g.keep) disc_part += si;        // summed directly: total - kept would cancel
    idx[rank] = i;
  }
  const double disc = block_sum(disc_part, red);
  const double frob = sqrt(fro2);
  const double scale = (g.nr_bulk && frob > 0.0) ? 1.0 / frob : 1.0;
  const int keep = g.keep;
  if (!mode_t) {
    // Vh_k = rows / s ;  US = A Vh_k^H  (A re-read from global memory: X was overwritten)
    for (int e = t; e < keep * n; e += nt) {
      const int k = e / n, c = e - k * n;
      const double sk2 = s2[idx[k]];
      Vh[e] = sk2 > floor2 ? cscale(X[idx[k] * q + c], rsqrt(sk2)) : cmake(0.0, 0.0);
    }
    for (int e = w; e < m * keep; e += nw) {
      const int r = e / keep, k = e - r * keep;
      const double sk2 = s2[idx[k]];
      const cplx* xr = X + idx[k] * q;
      const cplx* ar = A + (long long)r * g.lda;
      cplx acc = cmake(0.0, 0.0);
      for (int c = lane; c < n; c += 32) acc = cadd(acc, cmulc(ar[c], xr[c]));
      acc = warp_sum(acc);
      if (lane == 0) US[e] = sk2 > floor2 ? cscale(acc, rsqrt(sk2) * scale) : cmake(0.0, 0.0);
    }
  } else {
    // US_k = data part of the rows ;  Vh_k = conj of the accumulated-rotation part (orthonormal to rounding)
    for (int e = t; e < m * keep; e += nt) {
      const int r = e / keep, k = e - r * keep;
      US[e] = cscale(X[idx[k] * q + r], scale);
    }
    for (int e = t; e < keep * n; e += nt) {
      const int k = e / n, c = e - k * n;
      Vh[e] = s2[idx[k]] > floor2 ? cconj(X[idx[k] * q + m + c]) : cmake(0.0, 0.0);
    }
  }
  if (t == 0) {
    double* sl = slots + (long long)blockIdx.x * n_slots;
    if (g.nr_bulk && g.slot_lognorm >= 0 && frob > 0.0) sl[g.slot_lognorm] += log(frob);
    if (g.slot_trunc >= 0 && fro2 > 0.0) sl[g.slot_trunc] += sqrt(disc / fro2);
    if (!converged) sl[n_slots - 1] += 1.0;        // engine-reserved status slot: Jacobi did not converge
  }
}

void svd_small(const Arena& a, int64_t A, int64_t lda, int64_t US, int64_t Vh, int64_t m, int64_t n, int64_t keep, int nr_bulk,
               int slot_lognorm, int slot_trunc) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(svd_small_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM_MAX);
    cudaFuncSetAttribute(svd_small_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMALL_SMEM_MAX);
    attr_set = true;
  }
  SmallArgs g;
  g.A = A; g.US = US; g.Vh = Vh; g.m = (int)m; g.n = (int)n; g.lda = (int)lda; g.keep = (int)keep;
  g.nr_bulk = nr_bulk; g.slot_lognorm = slot_lognorm; g.slot_trunc = slot_trunc;
  const int p = (int)(m < n ? m : n), pp = (p + 1) & ~1, npairs = pp / 2 > 0 ? pp / 2 : 1;
  int G = 32;
  while (G > 8 && npairs * G > 1024) G >>= 1;
  g.group = G;
  int threads = npairs * G;
  threads = (threads + 31) / 32 * 32;
  if (threads < 256) threads = 256;
  if (threads > 1024) threads = 1024;
  const int q = (int)(m <= n ? n : m + n);
  if (threads <= 768 && (q + G - 1) / G <= 6)
    svd_small_kernel<true><<<a.nb, threads, svd_small_smem(m, n), a.stream>>>(a.base, a.chain_stride, a.slots, a.n_slots, g);
  else
    svd_small_kernel<false><<<a.nb, threads, svd_small_smem(m, n), a.stream>>>(a.base, a.chain_stride, a.slots, a.n_slots, g);
  ++*a.launches;
}

}  // namespace kbp
