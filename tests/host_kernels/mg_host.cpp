// TEST INFRASTRUCTURE -- not part of the product (see cuda_shim.h).  The smoother / transfer / coarse-solve kernels of
// the geometric-trace multigrid preconditioner (csrc/hdg_mg.cuh, the GPU apply of firedrake.GTMGPC,
// hdg_imex.py:138-169) as plain C functions; tests/test_mg_host.py strings them together the way mg_apply /
// mg_vcycle / mg_smooth_csr (csrc/hdg_engine.cu) do.
#include "cuda_shim.h"
#include "hdg_local.cuh"
// declarations hdg_mg.cuh expects from csrc/hdg_engine.cu (only the kernels called below are exercised)
struct CgScalars {
  double rz0, rz, tol2;
  int iters, done, maxit, pad;
};
constexpr int BLOCK = 256;
static inline double block_reduce(double v) { return v; }
static inline double reduce_partials(const double* part, int n) {
  double v = 0.0;
  for (int i = 0; i < n; ++i) v += part[i];
  return v;
}
#include "hdg_mg.cuh"

#define BY_B(b_, ...)               \
  switch (b_) {                     \
    case 2: { constexpr int b = 2; __VA_ARGS__; } return 0; \
    case 3: { constexpr int b = 3; __VA_ARGS__; } return 0; \
    case 4: { constexpr int b = 4; __VA_ARGS__; } return 0; \
    case 5: { constexpr int b = 5; __VA_ARGS__; } return 0; \
    default: return 1;              \
  }

extern "C" {

int mh_ell_cheb(int bs, int nf, const double* val, const int* col, const double* dinv, const double* bv, const double* x,
                double* d, double* xout, double cd, double cr, int zero) {
  BY_B(bs, k_ell_cheb<b>(nf, val, col, dinv, bv, x, d, xout, cd, cr, zero))
}

int mh_ell_residual(int bs, int nf, const double* val, const int* col, const double* bv, const double* x, double* r) {
  BY_B(bs, k_ell_residual<b>(nf, val, col, bv, x, r))
}

int mh_blockjac(int bs, int nf, const double* dinv, const double* r, double* z) {
  BY_B(bs, k_blockjac<b>(nf, dinv, r, z))
}

int mh_csr_diag_inv(int n, const int* rowptr, const int* col, const double* val, double* dinv) {
  k_csr_diag_inv(n, rowptr, col, val, dinv);
  return 0;
}

int mh_csr_spmv(int n, const int* rowptr, const int* col, const double* val, const double* x, const double* b, double* y,
                int mode) {
  k_csr_spmv(n, rowptr, col, val, x, b, y, mode);
  return 0;
}

int mh_csr_cheb(int n, const int* rowptr, const int* col, const double* val, const double* dinv, const double* b,
                const double* x, double* d, double* xout, double cd, double cr, int zero) {
  k_csr_cheb(n, rowptr, col, val, dinv, b, x, d, xout, cd, cr, zero);
  return 0;
}

int mh_dense_matvec(int n, const double* M, const double* b, double* x) {
  k_dense_matvec(n, M, b, x);
  return 0;
}
}
