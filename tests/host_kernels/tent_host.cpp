// TEST INFRASTRUCTURE -- not part of the product (see cuda_shim.h).  The velocity-side kernels of the tentative
// solve (csrc/hdg_flow.cuh, hdg_tent.cuh, hdg_advblock.cuh) as plain C functions, all arrays in the engine's SoA
// layouts; tests/test_tent_host.py strings them together the way run_tentative_aug (csrc/hdg_engine.cu) does.
#include "cuda_shim.h"
#include "hdg_flow.cuh"
#include "hdg_tent.cuh"
#include "hdg_advblock.cuh"

#define BY_K(k, ...)                \
  switch (k) {                      \
    case 1: { constexpr int K = 1; __VA_ARGS__; } return 0; \
    case 2: { constexpr int K = 2; __VA_ARGS__; } return 0; \
    case 3: { constexpr int K = 3; __VA_ARGS__; } return 0; \
    case 4: { constexpr int K = 4; __VA_ARGS__; } return 0; \
    default: return 1;              \
  }

extern "C" {

int th_setup(int nc, int nf, const double* xy, const int* cell_facet, const int* cell_flip, const int* facet_cell,
             const int* facet_local, int* nbr, int* nbr_e, double* tc, int* tcol, int* tbits) {
  k_build_nbr(cell_facet, facet_cell, facet_local, nc, nf, nbr, nbr_e);
  k_tent_setup(xy, cell_facet, cell_flip, facet_cell, facet_local, nc, nf, tc, tcol, tbits);
  return 0;
}

int th_fimpl(int k, int upwind, int nc, const double* xy, const int* nbr, const int* nbr_e, double alpha,
             const double* Qstar, const double* X, const double* Z, double c0, double c1, double* Y) {
  BY_K(k, if (upwind) k_fimpl<K, true>(xy, nbr, nbr_e, nc, alpha, Qstar, X, Z, c0, c1, Y);
          else k_fimpl<K, false>(xy, nbr, nbr_e, nc, alpha, Qstar, X, Z, c0, c1, Y))
}

// FP32 instantiations of the mixed-precision solver (run_tentative_mixed): operator in FP32 arithmetic, the
// bandwidth-bound kernels with float storage and FP64 registers
int th_fimpl32(int k, int upwind, int nc, const double* xy, const int* nbr, const int* nbr_e, double alpha,
               const float* Qstar, const float* X, const float* Z, float c0, float c1, float* Y) {
  BY_K(k, if (upwind) k_fimpl<K, true, float>(xy, nbr, nbr_e, nc, alpha, Qstar, X, Z, c0, c1, Y);
          else k_fimpl<K, false, float>(xy, nbr, nbr_e, nc, alpha, Qstar, X, Z, c0, c1, Y))
}
int th_moments32(int k, int nc, const double* xy, const int* flip, const float* Y, float* cm) {
  BY_K(k, (k_tent_moments<K, float>(xy, flip, nc, Y, cm)))
}
int th_trhs32(int k, int nc, int nf, const float* cm, const int* facet_cell, const int* facet_local, const float* ymu,
              float* t, float* nyx) {
  BY_K(k, (k_tent_trhs<K, float>(cm, facet_cell, facet_local, nc, nf, ymu, t, nyx)))
}
int th_xhat32(int k, int nc, int nf, const double* xy, const int* flip, const int* cell_facet, const float* Y,
              const float* mu, float* Xh, const double* sK, float* Zout) {
  BY_K(k, (k_tent_xhat<K, float>(xy, flip, cell_facet, nc, nf, Y, mu, Xh, 0, sK, Zout)))
}
int th_sweep_f(int k, int nf, const int* facet_local, const double* tc, const int* tcol, const int* tbits,
               double inv_aalpha, const float* rhs, const float* x, float* xout, int mode) {
  BY_K(k, (k_tent_sweep<K, 5, float>(nf, facet_local, tc, tcol, tbits, inv_aalpha, rhs, (const float*)nullptr, x,
                                     (float*)nullptr, xout, 0.0, 0.0, 0, mode)))
}

// operator with the Q*-dependent factors tabulated once per solve (k_fimpl_pre + k_fimpl_q); npre = FimplPre<K>::N
int th_fimpl_pre(int k, int nc, const double* xy, const double* Qstar, double* pre, int* npre) {
  BY_K(k, *npre = FimplPre<K>::N; if (pre) k_fimpl_pre<K>(xy, nc, Qstar, pre))
}
int th_fimpl_q(int k, int upwind, int nc, const double* xy, const int* nbr, const int* nbr_e, double alpha,
               const double* pre, const double* X, const double* Z, double c0, double c1, double* Y) {
  BY_K(k, if (upwind) k_fimpl_q<K, true>(xy, nbr, nbr_e, nc, alpha, pre, X, Z, c0, c1, Y);
          else k_fimpl_q<K, false>(xy, nbr, nbr_e, nc, alpha, pre, X, Z, c0, c1, Y))
}

// penalty-free operator with one thread per (cell, component) (k_fimpl_c), reading the table of k_fimpl_pre
int th_fimpl_c(int k, int upwind, int nc, const double* xy, const int* nbr, const int* nbr_e, const double* pre,
               const double* X, const double* Z, double c0, double c1, double* Y) {
  BY_K(k, if (upwind) k_fimpl_c<K, true>(xy, nbr, nbr_e, nc, pre, X, Z, c0, c1, Y);
          else k_fimpl_c<K, false>(xy, nbr, nbr_e, nc, pre, X, Z, c0, c1, Y))
}

// entry (j, l) of the reference Gram block GG(e, f) = BF_e BF_f^T the Schur-complement sweeps are built from
int th_gram(int k, int e, int f, int j, int l, double* out) {
  BY_K(k, *out = RefTables<K>::GG(e, f, j, l))
}

int th_moments(int k, int nc, const double* xy, const int* flip, const double* Y, double* cm) {
  BY_K(k, k_tent_moments<K>(xy, flip, nc, Y, cm))
}

int th_trhs(int k, int nc, int nf, const double* cm, const int* facet_cell, const int* facet_local, const double* ymu,
            double* t, double* nyx) {
  BY_K(k, k_tent_trhs<K>(cm, facet_cell, facet_local, nc, nf, ymu, t, nyx))
}

int th_sweep(int k, int nf, const int* facet_local, const double* tc, const int* tcol, const int* tbits,
             double inv_aalpha, const double* rhs, const double* rhs2, const double* x, double* d, double* xout,
             double cd, double cr, int zero, int mode) {
  BY_K(k, (k_tent_sweep<K, 5>(nf, facet_local, tc, tcol, tbits, inv_aalpha, rhs, rhs2, x, d, xout, cd, cr, zero, mode)))
}

// FP32-stored sweep (mode 0): exactly one of xout32 / xout64 is given
int th_sweep32(int k, int nf, const int* facet_local, const double* tc, const int* tcol, const int* tbits,
               double inv_aalpha, const double* rhs, const float* x, float* d, float* xout32, double* xout64, double cd,
               double cr, int zero) {
  BY_K(k, (k_tent_sweep32<K, 5>(nf, facet_local, tc, tcol, tbits, inv_aalpha, rhs, x, d, xout32, xout64, cd, cr, zero)))
}

int th_xhat(int k, int nc, int nf, const double* xy, const int* flip, const int* cell_facet, const double* Y,
            const double* mu, double* Xh, int mode) {
  BY_K(k, k_tent_xhat<K>(xy, flip, cell_facet, nc, nf, Y, mu, Xh, mode))
}

int th_advblock(int k, int upwind, int nc, const double* xy, const int* nbr, const double* Qstar, double adt,
                double* blk, float* blk32) {
  BY_K(k, for (int cell = 0; cell < nc; ++cell) {
    if (upwind) advblock_build_cell<K, true>(xy, nbr, nc, cell, Qstar, adt, blk);
    else advblock_build_cell<K, false>(xy, nbr, nc, cell, Qstar, adt, blk);
    advblock_invert_cell<Dims<K>::NQ1>(nc, cell, blk, blk32);
  })
}

// as th_advblock, also returning sK[cell] = tr(C_K) / NQ1 (scaled Schur complement)
int th_advblock_sk(int k, int upwind, int nc, const double* xy, const int* nbr, const double* Qstar, double adt,
                   double* blk, float* blk32, double* sK) {
  BY_K(k, for (int cell = 0; cell < nc; ++cell) {
    if (upwind) advblock_build_cell<K, true>(xy, nbr, nc, cell, Qstar, adt, blk);
    else advblock_build_cell<K, false>(xy, nbr, nc, cell, Qstar, adt, blk);
    advblock_invert_cell<Dims<K>::NQ1>(nc, cell, blk, blk32, sK);
  })
}

int th_scale_tc(int nf, const int* facet_cell, const double* tc, const double* sK, double* tcs) {
  k_tent_scale_tc(nf, facet_cell, tc, sK, tcs);
  return 0;
}

int th_xhat_scaled(int k, int nc, int nf, const double* xy, const int* flip, const int* cell_facet, const double* Y,
                   const double* mu, double* Xh, int mode, const double* sK, double* Zout) {
  BY_K(k, k_tent_xhat<K>(xy, flip, cell_facet, nc, nf, Y, mu, Xh, mode, sK, Zout))
}

int th_elem_bound(int k, int nc, const double* xy, double* lam) {
  BY_K(k, k_tent_elem_bound<K>(xy, nc, lam))
}

int th_advblock_apply(int k, int nc, const float* blk32, const double* X, double* Y) {
  BY_K(k, for (int cell = 0; cell < nc; ++cell) advblock_apply_cell<K>(nc, cell, blk32, X, Y))
}
}
