// TEST INFRASTRUCTURE -- not part of the product (see cuda_shim.h).  Velocity-side kernels of csrc/hdg_flow.cuh
// (BDM projection, weak divergence, pressure gradient, trace reconstruction) as plain C functions, all arrays in
// the engine's SoA layouts; the call sequences are those of the hdg_*_dev entry points in csrc/hdg_engine.cu.
#include "cuda_shim.h"
#include "hdg_flow.cuh"

#define BY_K(k, ...)                \
  switch (k) {                      \
    case 1: { constexpr int K = 1; __VA_ARGS__; } return 0; \
    case 2: { constexpr int K = 2; __VA_ARGS__; } return 0; \
    case 3: { constexpr int K = 3; __VA_ARGS__; } return 0; \
    case 4: { constexpr int K = 4; __VA_ARGS__; } return 0; \
    default: return 1;              \
  }

extern "C" {

int fh_build_nbr(int nc, int nf, const int* cell_facet, const int* facet_cell, const int* facet_local, int* nbr,
                 int* nbr_e) {
  k_build_nbr(cell_facet, facet_cell, facet_local, nc, nf, nbr, nbr_e);
  return 0;
}

// hdg_project_bdm_dev: fm is scratch [2 (k + 2)][nf]
int fh_project_bdm(int k, int nc, int nf, const double* xy, const int* cell_facet, const int* facet_cell,
                   const double* Q, double* fm, double* Qs) {
  BY_K(k, k_bdm_moments<K>(xy, cell_facet, facet_cell, nc, nf, Q, fm);
          k_bdm_lift<K>(xy, cell_facet, facet_cell, nc, nf, Q, fm, Qs))
}

int fh_weak_div(int k, int nc, const double* xy, const int* nbr, const int* nbr_e, const double* Q, double scale,
                int mode, double* Rp) {
  BY_K(k, k_weak_div<K>(xy, nbr, nbr_e, nc, Q, scale, mode, Rp))
}

int fh_pgrad(int k, int nc, int nf, const double* xy, const int* flip, const int* cell_facet, const double* p,
             const double* lam, double c0, double c1, double* Y) {
  BY_K(k, k_pgrad<K>(xy, flip, cell_facet, nc, nf, p, lam, c0, c1, Y))
}

// hdg_gamma_apply_dev: constraint rows Gamma(psi, mu; u, phi, lambda) of the mixed operator; gK is scratch
int fh_gamma(int k, int nc, int nf, const double* xy, const int* flip, const int* cell_facet, const int* facet_cell,
             const int* facet_local, double tau, const double* Q, const double* p, const double* lam, double* Rp,
             double* gK, double* Rl) {
  BY_K(k, k_gamma_cell<K>(xy, flip, cell_facet, nc, nf, tau, Q, p, lam, Rp, gK);
          k_facet_sum<K>(gK, facet_cell, facet_local, nc, nf, Rl))
}

// hdg_reconstruction_rhs_dev: Rl must be zero on entry
int fh_recon_rhs(int k, int nc, int nf, const double* xy, const int* nbr, const int* nbr_e, const int* cell_facet,
                 const int* flip, const double* Q, const double* B, double* Rp, double* Rl) {
  BY_K(k, k_recon_rhs<K>(xy, nbr, nbr_e, cell_facet, flip, nc, nf, Q, B, Rp, Rl))
}

// hdg_reconstruct_trace_dev: gK is scratch [3 (k + 1)][nc]
int fh_reconstruct_trace(int k, int nc, int nf, const double* xy, const int* flip, const int* facet_cell,
                         const int* facet_local, double tau, const double* Q, const double* p, double* gK,
                         double* lam) {
  BY_K(k, k_trace_moments<K>(xy, flip, nc, tau, Q, p, gK); k_trace_avg<K>(gK, facet_cell, facet_local, nc, nf, lam))
}
}
