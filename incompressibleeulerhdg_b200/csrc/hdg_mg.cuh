// Geometric-trace multigrid preconditioner for the condensed trace system: the GPU apply of the
// reference's firedrake.GTMGPC (src/timesteppers/hdg_imex.py:138-169):
//   fine level   trace space, Chebyshev(n) + facet-block-Jacobi (ASMStarPC patches, :143-152)
//   coarse space conforming P1 on the same mesh (:97-106), transfer = trace of the P1 function
//   coarse solve one V-cycle over a nested P1 hierarchy (stand-in for GAMG, :153-167) with
//                Chebyshev(n) + point-Jacobi smoothing and a dense pseudo-inverse on the coarsest level
// All matrices of the hierarchy are built on the host at setup (multigrid.py) and uploaded as CSR;
// every kernel here is deterministic (thread-per-row, no atomics).
#pragma once
#include <vector>

struct DevCsr {
  int nrows = 0, ncols = 0, nnz = 0;
  int *rowptr = nullptr, *col = nullptr;
  double* val = nullptr;
};

struct MgLevel {
  int n = 0;      // vector length (local entries: owned + ghost on a distributed level)
  int nrows = 0;  // rows this rank computes (owned entries; == n on replicated levels / single GPU)
  DevCsr A, P, R;  // P: level l+1 -> l (n_l x n_{l+1}), R = P^T
  double* dinv = nullptr;
  double *x = nullptr, *x2 = nullptr, *b = nullptr, *r = nullptr, *d = nullptr;
  double lmax = 2.0;
};

struct MgState {
  bool enabled = false;
  int nlevels = 0;
  std::vector<MgLevel> L;
  DevCsr T, Tt;  // trace <- P1 level 0 and its transpose
  double* pinv = nullptr;
  int n_last = 0;
  int ns_fine = 1, ns_coarse = 1;
  double ratio = 10.0;
  double fine_lmax = 2.0;
  double *fx = nullptr, *fx2 = nullptr, *fd = nullptr, *fr = nullptr;  // fine work vectors [b*nf]
  // multi-GPU: levels l < repl are row-distributed (halo plan PLAN_P1 + l), levels >= repl replicated;
  // the owned rows of level `repl` are all-gathered (padded to gmax per rank) and scattered by gid
  int repl = 0;
  int gmax = 0;
  int *gptr = nullptr, *ggid = nullptr;  // device [nranks+1], [sum counts]
  double *gsend = nullptr, *gbuf = nullptr;
};

struct ChebCoef {
  double cd, cr;
};
// coefficients of sweep j (0-based) of Chebyshev iteration on [lmax/ratio, 1.1 lmax]:
//   d <- cd * d + cr * (D^-1 residual);  x <- x + d
static inline void cheb_coefs(double lmax, double ratio, int nsweeps, std::vector<ChebCoef>& out) {
  double a = lmax / ratio, b = 1.1 * lmax;
  double theta = 0.5 * (b + a), delta = 0.5 * (b - a), sigma = theta / delta, rho = 1.0 / sigma;
  out.clear();
  out.push_back({0.0, 1.0 / theta});
  for (int j = 1; j < nsweeps; ++j) {
    double rho_new = 1.0 / (2.0 * sigma - rho);
    out.push_back({rho_new * rho, 2.0 * rho_new / delta});
    rho = rho_new;
  }
}

__global__ void k_csr_diag_inv(int n, const int* __restrict__ rowptr, const int* __restrict__ col,
                               const double* __restrict__ val, double* __restrict__ dinv) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double d = 1.0;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
      if (col[k] == i) d = val[k];
    dinv[i] = 1.0 / d;
  }
}

// mode 0: y = A x     mode 1: y += A x     mode 2: y = b - A x
__global__ void k_csr_spmv(int n, const int* __restrict__ rowptr, const int* __restrict__ col,
                           const double* __restrict__ val, const double* __restrict__ x, const double* __restrict__ b,
                           double* __restrict__ y, int mode) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) s = fma(val[k], x[col[k]], s);
    if (mode == 0)
      y[i] = s;
    else if (mode == 1)
      y[i] += s;
    else
      y[i] = b[i] - s;
  }
}

// one Chebyshev/Jacobi sweep on a CSR level: r = dinv (b - A x) (x == 0 if zero), d = cd d + cr r,
// xout = x + d.  Jacobi-type: xout must not alias x.
__global__ void k_csr_cheb(int n, const int* __restrict__ rowptr, const int* __restrict__ col,
                           const double* __restrict__ val, const double* __restrict__ dinv,
                           const double* __restrict__ b, const double* __restrict__ x, double* __restrict__ d,
                           double* __restrict__ xout, double cd, double cr, int zero) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double s = b[i];
    double xi = 0.0;
    if (!zero) {
      for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) s = fma(-val[k], x[col[k]], s);
      xi = x[i];
    }
    double di = cr * dinv[i] * s;
    if (cd != 0.0) di = fma(cd, d[i], di);
    d[i] = di;
    xout[i] = xi + di;
  }
}

// x = pinv b on the coarsest level (dense, row per thread).  The pseudo-inverse of the symmetric
// coarse operator is symmetric, so row i is read as column i: consecutive threads touch consecutive
// addresses (coalesced) and the summation order over j stays fixed.
__global__ void k_dense_matvec(int n, const double* __restrict__ M, const double* __restrict__ b,
                               double* __restrict__ x) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int j = 0; j < n; ++j) s = fma(M[(size_t)j * n + i], b[j], s);
    x[i] = s;
  }
}

// fine level (blocked ELL, P = -S): one Chebyshev / block-Jacobi sweep
//   r = Dinv (bv - P x), d = cd d + cr r, xout = x + d      (xout must not alias x)
template <int b>
__global__ void __launch_bounds__(256) k_ell_cheb(int nf, const double* __restrict__ val, const int* __restrict__ col,
                                                  const double* __restrict__ dinv, const double* __restrict__ bv,
                                                  const double* __restrict__ x, double* __restrict__ d,
                                                  double* __restrict__ xout, double cd, double cr, int zero) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    double r[b], xo[b];
    HDG_UNROLL
    for (int i = 0; i < b; ++i) {
      r[i] = bv[(size_t)i * nf + f];
      xo[i] = 0.0;
    }
    if (!zero) {
      HDG_UNROLL
      for (int j = 0; j < 5; ++j) {
        int cj = col[(size_t)j * nf + f];
        double xv[b];
        HDG_UNROLL
        for (int c = 0; c < b; ++c) xv[c] = x[(size_t)c * nf + cj];
        if (j == 0) {
          HDG_UNROLL
          for (int c = 0; c < b; ++c) xo[c] = xv[c];
        }
        HDG_UNROLL
        for (int rr = 0; rr < b; ++rr)
          HDG_UNROLL
          for (int c = 0; c < b; ++c) r[rr] = fma(-val[(size_t)((j * b + rr) * b + c) * nf + f], xv[c], r[rr]);
      }
    }
    HDG_UNROLL
    for (int i = 0; i < b; ++i) {
      double z = 0.0;
      HDG_UNROLL
      for (int j = 0; j < b; ++j) z = fma(dinv[(size_t)(i * b + j) * nf + f], r[j], z);
      double di = cr * z;
      if (cd != 0.0) di = fma(cd, d[(size_t)i * nf + f], di);
      d[(size_t)i * nf + f] = di;
      xout[(size_t)i * nf + f] = xo[i] + di;
    }
  }
}

// r = bv - P x on the fine level
template <int b>
__global__ void __launch_bounds__(256) k_ell_residual(int nf, const double* __restrict__ val,
                                                      const int* __restrict__ col, const double* __restrict__ bv,
                                                      const double* __restrict__ x, double* __restrict__ r) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    double y[b];
    HDG_UNROLL
    for (int i = 0; i < b; ++i) y[i] = bv[(size_t)i * nf + f];
    HDG_UNROLL
    for (int j = 0; j < 5; ++j) {
      int cj = col[(size_t)j * nf + f];
      double xv[b];
      HDG_UNROLL
      for (int c = 0; c < b; ++c) xv[c] = x[(size_t)c * nf + cj];
      HDG_UNROLL
      for (int rr = 0; rr < b; ++rr)
        HDG_UNROLL
        for (int c = 0; c < b; ++c) y[rr] = fma(-val[(size_t)((j * b + rr) * b + c) * nf + f], xv[c], y[rr]);
    }
    HDG_UNROLL
    for (int i = 0; i < b; ++i) r[(size_t)i * nf + f] = y[i];
  }
}

// z = Dinv r (block-Jacobi apply), used by the power iteration for lambda_max(Dinv P)
template <int b>
__global__ void __launch_bounds__(256) k_blockjac(int nf, const double* __restrict__ dinv, const double* __restrict__ r,
                                                  double* __restrict__ z) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    double rv[b];
    HDG_UNROLL
    for (int i = 0; i < b; ++i) rv[i] = r[(size_t)i * nf + f];
    HDG_UNROLL
    for (int i = 0; i < b; ++i) {
      double s = 0.0;
      HDG_UNROLL
      for (int j = 0; j < b; ++j) s = fma(dinv[(size_t)(i * b + j) * nf + f], rv[j], s);
      z[(size_t)i * nf + f] = s;
    }
  }
}

__global__ void k_scale(size_t n, double c, double* __restrict__ x) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] *= c;
}

// x += alpha p ; r -= alpha q   with alpha = <r,z>/<p,q> (plain CG update without preconditioner)
// part_q0 (optional): partial sums of the mode-0 coefficients of q = P p.  The assembled P annihilates the
// constants only to ~1e-13 (S_K 1 = 0 holds to the accuracy of the per-cell Cholesky), i.e. it is a
// slightly perturbed singular matrix; once the residual reaches ~1e-12 CG starts chasing that spurious
// near-null mode (alpha explodes, the residual bounces).  Removing the constant from q makes the operator
// the CG sees exactly (I - Pi) P (I - Pi) -- what PETSc's MatNullSpace does for the reference
// (`nullspace=` at hdg_imex.py:186,196,218).
__global__ void __launch_bounds__(256) k_cg_update_plain(size_t n, const double* __restrict__ p,
                                                         const double* __restrict__ q, double* __restrict__ x,
                                                         double* __restrict__ r, const double* __restrict__ part_pq,
                                                         const CgScalars* __restrict__ s,
                                                         const double* __restrict__ part_q0 = nullptr,
                                                         double inv_nf_glob = 0.0, size_t nf = 0) {
  if (s->done) return;
  double pq = reduce_partials(part_pq, gridDim.x);
  double alpha = s->rz / pq;
  const double qmean = part_q0 ? reduce_partials(part_q0, gridDim.x) * inv_nf_glob : 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    x[i] = fma(alpha, p[i], x[i]);
    r[i] = fma(-alpha, q[i] - (i < nf ? qmean : 0.0), r[i]);
  }
}
