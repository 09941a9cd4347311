// TEST INFRASTRUCTURE -- not part of the product (see cuda_shim.h).  The per-cell / per-facet kernels of the condensed
// mixed-Poisson path (csrc/hdg_poisson.cuh) as plain C functions, all arrays in the engine's SoA layouts.
#include "cuda_shim.h"
#include "hdg_poisson.cuh"
#include "hdg_poisson_s.cuh"

#define BY_K(k, ...)                \
  switch (k) {                      \
    case 1: { constexpr int K = 1; __VA_ARGS__; } return 0; \
    case 2: { constexpr int K = 2; __VA_ARGS__; } return 0; \
    case 3: { constexpr int K = 3; __VA_ARGS__; } return 0; \
    case 4: { constexpr int K = 4; __VA_ARGS__; } return 0; \
    default: return 1;              \
  }
// the shared-memory-factor variants exist for k >= 3 only (hdg_poisson_s.cuh); on the host the "shared" column is a
// static array and threadIdx.x = 0, so the [NH][BD] indexing is exercised with its real strides
#define BY_K34(k, ...)              \
  switch (k) {                      \
    case 3: { constexpr int K = 3; __VA_ARGS__; } return 0; \
    case 4: { constexpr int K = 4; __VA_ARGS__; } return 0; \
    default: return 1;              \
  }

extern "C" {

int ph_condense(int k, int nc, const double* xy, const int* flip, double tau, double* SK) {
  BY_K(k, k_condense<K>(xy, flip, nc, tau, SK))
}

int ph_assemble(int k, int nc, int nf, const double* SK, const int* cell_facet, const int* facet_cell,
                const int* facet_local, double* val, int* col, double* dinv) {
  BY_K(k, k_assemble<K>(SK, cell_facet, facet_cell, facet_local, nc, nf, val, col, dinv))
}

int ph_forward(int k, int nc, const double* xy, const int* flip, double tau, const double* Ru, const double* Rp,
               double* gK) {
  BY_K(k, k_forward<K>(xy, flip, nc, tau, Ru, Rp, gK))
}

int ph_back(int k, int nc, int nf, const double* xy, const int* flip, const int* cell_facet, double tau,
            const double* Ru, const double* Rp, const double* lam, double* uo, double* po) {
  BY_K(k, k_back<K>(xy, flip, cell_facet, nc, nf, tau, Ru, Rp, lam, uo, po))
}

// back-substitution fused with the caller's update (k_back_update); partial[0] = sum over the owned cells of detJ phi_0
int ph_back_update(int k, int nc, int nc_own, int nf, const double* xy, const int* flip, const int* cell_facet, double tau,
                   const double* Ru, const double* Rp, const double* lam, double cq, double cb, double cu, double cp,
                   const double* Qbase, double* Qacc, double* pacc, double* partial) {
  BackUpdate U{cq, cb, cu, cp, Qbase, Qacc, pacc, partial, nc_own};
  BY_K(k, k_back_update<K>(xy, flip, cell_facet, nc, nf, tau, Ru, Rp, lam, U))
}

int ph_condense_s(int k, int nc, const double* xy, const int* flip, double tau, double* SK) {
  BY_K34(k, k_condense_b<K>(xy, flip, nc, tau, SK))
}

int ph_forward_s(int k, int nc, const double* xy, const int* flip, double tau, const double* Ru, const double* Rp,
                 double* gK) {
  BY_K34(k, k_forward_s<K>(xy, flip, nc, tau, Ru, Rp, gK))
}

int ph_back_s(int k, int nc, int nf, const double* xy, const int* flip, const int* cell_facet, double tau,
              const double* Ru, const double* Rp, const double* lam, double* uo, double* po) {
  BY_K34(k, k_back_s<K>(xy, flip, cell_facet, nc, nf, tau, Ru, Rp, lam, uo, po))
}

int ph_back_update_s(int k, int nc, int nc_own, int nf, const double* xy, const int* flip, const int* cell_facet,
                     double tau, const double* Ru, const double* Rp, const double* lam, double cq, double cb, double cu,
                     double cp, const double* Qbase, double* Qacc, double* pacc, double* partial) {
  BackUpdate U{cq, cb, cu, cp, Qbase, Qacc, pacc, partial, nc_own};
  BY_K34(k, k_back_update_s<K>(xy, flip, cell_facet, nc, nf, tau, Ru, Rp, lam, U))
}
}
