// Batched complex128 Householder QR, one CTA per chain.  A(m x n, row-major) = Q(m x r) R(r x n),
// r = min(m, n), Q with orthonormal columns even when A is rank deficient (the boundary MPS is
// rank deficient on the first swallows of every chain, so Cholesky-type QR is not an option).
// The matrix is worked on column-major in a scratch buffer (L2-resident: <= a few MB), one warp per
// trailing column, lanes strided over rows, dot products reduced with warp shuffles.
#include "kbp_common.cuh"
#include "kbp_ops.cuh"

namespace kbp {

__global__ void __launch_bounds__(512) qr_householder_kernel(cplx* __restrict__ base, long long chain_stride, long long A_,
                                                             long long Q_, long long R_, long long work_, int m, int n,
                                                             const int* __restrict__ mask, int mask_want) {
  if (mask && mask[blockIdx.x] != mask_want) return;
  __shared__ double red[34];
  __shared__ cplx sh_inv_u0, sh_s;
  __shared__ double sh_tau;
  cplx* cb = base + (long long)blockIdx.x * chain_stride;
  const cplx* A = cb + A_;
  cplx* Q = cb + Q_;
  cplx* R = cb + R_;
  cplx* W = cb + work_;                        // m x n, column-major
  const int kk = m < n ? m : n;
  cplx* Qw = W + (long long)m * n;             // m x kk, column-major
  double* tau = reinterpret_cast<double*>(Qw + (long long)m * kk);  // kk doubles
  const int t = threadIdx.x, nt = blockDim.x, lane = t & 31, w = t >> 5, nw = nt >> 5;

  for (long long i = t; i < (long long)m * n; i += nt) {
    int r = (int)(i / n), c = (int)(i % n);
    W[(long long)c * m + r] = A[i];
  }
  __syncthreads();

  for (int j = 0; j < kk; ++j) {
    cplx* x = W + (long long)j * m;
    // column scale first: boundary-MPS columns can sit at 1e-150 (rows that the truncation emptied), where
    // squaring underflows -- all norms are formed from entries divided by the column's largest component
    double cmx = 0.0;
    for (int i = j + t; i < m; i += nt) cmx = fmax(cmx, fmax(fabs(x[i].x), fabs(x[i].y)));
    cmx = warp_max(cmx);
    __syncthreads();
    if (lane == 0) red[w] = cmx;
    __syncthreads();
    if (w == 0) {
      double v = lane < nw ? red[lane] : 0.0;
      v = warp_max(v);
      if (lane == 0) red[33] = v;
    }
    __syncthreads();
    cmx = red[33];
    const double cinv = cmx > 0.0 ? 1.0 / cmx : 0.0;
    double sig = 0.0;                                   // scaled:  sum_{i>j} |x_i / cmx|^2
    for (int i = j + 1 + t; i < m; i += nt) { const cplx y = cscale(x[i], cinv); sig += cabs2(y); }
    sig = block_sum(sig, red);
    if (t == 0) {
      const cplx alpha = x[j];
      const cplx as = cscale(alpha, cinv);
      const double absa_s = sqrt(cabs2(as));
      const double nrm_s = sqrt(fma(absa_s, absa_s, sig));
      if (!(cmx > 0.0) || nrm_s == 0.0) {
        sh_tau = 0.0; sh_s = cmake(0.0, 0.0); sh_inv_u0 = cmake(0.0, 0.0);
      } else {
        const cplx ph = absa_s > 0.0 ? cscale(as, 1.0 / absa_s) : cmake(1.0, 0.0);
        const double u0_s = absa_s + nrm_s;               // |u0| / cmx : no cancellation, no squaring of tiny numbers
        const double r = sqrt(sig) / u0_s;                // <= 1
        sh_tau = 2.0 / (1.0 + r * r);                     // = 2 |u0|^2 / (|u0|^2 + sum_{i>j} |x_i|^2)
        sh_inv_u0 = cscale(cconj(ph), cinv / u0_s);       // 1 / u0,  u0 = ph (|alpha| + ||x||)
        sh_s = cscale(ph, -nrm_s * cmx);                  // R_jj
      }
      tau[j] = sh_tau;
    }
    __syncthreads();
    const cplx inv_u0 = sh_inv_u0;
    const double tj = sh_tau;
    for (int i = j + 1 + t; i < m; i += nt) x[i] = cmul(x[i], inv_u0);
    if (t == 0) x[j] = sh_s;
    __syncthreads();
    if (tj != 0.0) {
      for (int c = j + 1 + w; c < n; c += nw) {
        cplx* a = W + (long long)c * m;
        cplx d = cmake(0.0, 0.0);
        for (int i = j + 1 + lane; i < m; i += 32) d = cadd(d, ccmul(x[i], a[i]));
        d = warp_sum(d);
        d = cadd(d, a[j]);
        d = cscale(d, tj);
        for (int i = j + 1 + lane; i < m; i += 32) a[i] = csub(a[i], cmul(x[i], d));
        __syncwarp();
        if (lane == 0) a[j] = csub(a[j], d);
      }
    }
    __syncthreads();
  }

  // R (kk x n, row-major): upper triangle of W
  for (long long i = t; i < (long long)kk * n; i += nt) {
    int r = (int)(i / n), c = (int)(i % n);
    R[i] = r <= c ? W[(long long)c * m + r] : cmake(0.0, 0.0);
  }
  // Q = H_0 H_1 ... H_{kk-1} applied to the first kk columns of the identity
  for (long long i = t; i < (long long)m * kk; i += nt) {
    int c = (int)(i / m), r = (int)(i % m);
    Qw[i] = r == c ? cmake(1.0, 0.0) : cmake(0.0, 0.0);
  }
  __syncthreads();
  for (int j = kk - 1; j >= 0; --j) {
    const cplx* v = W + (long long)j * m;
    const double tj = tau[j];
    if (tj != 0.0) {
      for (int c = j + w; c < kk; c += nw) {
        cplx* qc = Qw + (long long)c * m;
        cplx d = cmake(0.0, 0.0);
        for (int i = j + 1 + lane; i < m; i += 32) d = cadd(d, ccmul(v[i], qc[i]));
        d = warp_sum(d);
        d = cadd(d, qc[j]);
        d = cscale(d, tj);
        for (int i = j + 1 + lane; i < m; i += 32) qc[i] = csub(qc[i], cmul(v[i], d));
        __syncwarp();
        if (lane == 0) qc[j] = csub(qc[j], d);
      }
    }
    __syncthreads();
  }
  for (long long i = t; i < (long long)m * kk; i += nt) {
    int r = (int)(i / kk), c = (int)(i % kk);
    Q[i] = Qw[(long long)c * m + r];
  }
}

bool qr_cluster_fits(int64_t m, int64_t n);
void qr_householder(const Arena& a, int64_t A, int64_t Q, int64_t R, int64_t work, int64_t m, int64_t n);

// Tall matrices the cluster kernel does not take (1024 x 64 at D = 4, 2592 x 72 at D = 6): TSQR.  The rows are cut into
// nbk blocks that it does take, A_i = Q_i R_i; the stacked R_i (nbk n x n) are factored again, [R_1; ...; R_nbk] = Qs R; then
// Q rows of block i = Q_i Qs_i.  Householder at both levels: orthonormal Q also for rank-deficient input.
static bool qr_tsqr(const Arena& a, int64_t A, int64_t Q, int64_t R, int64_t work, int64_t m, int64_t n) {
  if (m < 2 * n) return false;
  for (int nbk = 2; nbk <= 8; ++nbk) {
    const int64_t mb = (m + nbk - 1) / nbk, last = m - mb * (nbk - 1);
    if (last < n || 4 * nbk * n > m) break;
    if (!qr_cluster_fits(mb, n) || !qr_cluster_fits(last, n) || !qr_cluster_fits(nbk * n, n)) continue;
    const int64_t T = work, S = T + m * n, Qs = S + nbk * n * n, W2 = Qs + nbk * n * n;
    for (int i = 0; i < nbk; ++i) {
      const int64_t rows = i + 1 < nbk ? mb : last;
      if (!qr_cluster(a, A + i * mb * n, T + i * mb * n, S + i * n * n, rows, n)) return false;   // (nothing launched yet if i == 0)
    }
    qr_householder(a, S, Qs, R, W2, nbk * n, n);
    for (int i = 0; i < nbk; ++i) {
      const int64_t rows = i + 1 < nbk ? mb : last;
      gemm(a, Q + i * mb * n, T + i * mb * n, Qs + i * n * n, rows, n, n, OP_N, OP_N);
    }
    return true;
  }
  return false;
}

// the unconditionally stable path: Householder
void qr_householder(const Arena& a, int64_t A, int64_t Q, int64_t R, int64_t work, int64_t m, int64_t n) {
  if (qr_cluster(a, A, Q, R, m, n)) return;                // shared-memory resident, rows over a cluster (k_qr_cluster.cu)
  if (qr_tsqr(a, A, Q, R, work, m, n)) return;
  qr_householder_kernel<<<a.nb, 512, 0, a.stream>>>(a.base, a.chain_stride, A, Q, R, work, (int)m, (int)n, a.mask, a.mask_want);
  ++*a.launches;
}

bool qr_cholqr2(const Arena& a, int64_t A, int64_t Q, int64_t R, int64_t work, int64_t m, int64_t n);   // k_tsvd.cu
bool qr_blocked(const Arena& a, int64_t A, int64_t Q, int64_t R, int64_t work, int64_t m, int64_t n);   // k_tsvd.cu

void qr(const Arena& a, int64_t A, int64_t Q, int64_t R, int64_t work, int64_t m, int64_t n) {
  if (m == 0 || n == 0) return;
  if (qr_cholqr2(a, A, Q, R, work, m, n)) return;          // tall and skinny: Cholesky-QR twice, Householder only if the Gram matrix is too ill conditioned
  if (qr_blocked(a, A, Q, R, work, m, n)) return;          // large, not skinny: block Gram-Schmidt over tall-skinny blocks
  qr_householder(a, A, Q, R, work, m, n);
}

}  // namespace kbp
