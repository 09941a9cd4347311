// TEST INFRASTRUCTURE -- not part of the product (see cuda_shim.h).  The passive-tracer kernels of
// csrc/hdg_tracer.cuh on the CPU: the matrix-free Jacobi-PCG projection onto [CG_{k+1}]^2 (a line-by-line port of
// the launch sequence of run_project_cg in csrc/hdg_engine.cu, with one block of one thread, so every two-stage
// reduction has a single partial) and the two advection kernels.
#include <vector>

#include "cuda_shim.h"
// the engine's block-level reduction helpers (csrc/hdg_engine.cu), as they act for a block of one thread
constexpr int BLOCK = 256;
static inline double block_reduce(double v) { return v; }
static inline double reduce_partials(const double* part, int n) {
  double v = 0.0;
  for (int i = 0; i < n; ++i) v += part[i];
  return v;
}
#include "hdg_tracer.cuh"

template <int K>
static int project_cg(int nc, int ncg, const double* xy, const int* cellmap, const int* inc_ptr, const int* inc_idx,
                      const double* dinv, const double* Q, double* Qcg, double rtol, int maxit, int* iters) {
  constexpr int NLOC = Dims<K>::NQ1;
  const int G = 1, own = ncg;
  const size_t cs = (size_t)NLOC * nc;
  std::vector<double> yK(2 * cs), x(2 * (size_t)ncg), r(x.size()), z(x.size()), p(x.size()), Ap(x.size()), part(2 * G);
  TracerScalars scal;
  k_cgp_load<K>(xy, nc, Q, yK.data());
  k_cgp_gather<0>(ncg, own, cs, inc_ptr, inc_idx, yK.data(), dinv, x.data(), r.data(), z.data(), p.data(), Ap.data(),
                  part.data());
  k_cgp_finish(part.data(), G, &scal, 0, 0, 1);
  int par = 0, it = 0;
  const double tol2 = rtol * rtol;
  bool done = false;
  while (!done && it < maxit) {
    ++it;
    k_cgp_cellop<K>(xy, nc, ncg, cellmap, p.data(), yK.data());
    k_cgp_gather<1>(ncg, own, cs, inc_ptr, inc_idx, yK.data(), dinv, x.data(), r.data(), z.data(), p.data(), Ap.data(),
                    part.data());
    k_cgp_finish(part.data(), G, &scal, 1, par, 0);
    k_cgp_update(ncg, own, &scal, par, dinv, p.data(), Ap.data(), x.data(), r.data(), z.data(), part.data());
    k_cgp_finish(part.data(), G, &scal, 0, par ^ 1, 0);
    k_cgp_dir(ncg, &scal, par, z.data(), p.data());
    par ^= 1;
    if (it % 4 == 0 || it == maxit)
      done = scal.rz[par][0] <= tol2 * scal.rz0[0] && scal.rz[par][1] <= tol2 * scal.rz0[1];
  }
  k_cgp_tocell<K>(nc, ncg, cellmap, x.data(), Qcg);
  *iters = it;
  return done ? 0 : 2;
}

#define BY_K(k, ...)                \
  switch (k) {                      \
    case 1: { constexpr int K = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int K = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int K = 3; __VA_ARGS__; } break; \
    case 4: { constexpr int K = 4; __VA_ARGS__; } break; \
    default: return 1;              \
  }

extern "C" {

int trh_project_cg(int k, int nc, int ncg, const double* xy, const int* cellmap, const int* inc_ptr,
                   const int* inc_idx, const double* dinv, const double* Q, double* Qcg, double rtol, int maxit,
                   int* iters) {
  BY_K(k, return project_cg<K>(nc, ncg, xy, cellmap, inc_ptr, inc_idx, dinv, Q, Qcg, rtol, maxit, iters))
  return 1;
}

// compile-time tables (default facet rule)
int trh_advect_t(int k, int nc, const double* xy, const int* nbr, const int* nbr_e, const double* U, const double* q,
                 double c0, const double* acc, double c1, double* out) {
  BY_K(k, k_tracer_adv_t<K>(xy, nbr, nbr_e, nc, U, q, c0, acc, c1, out))
  return 0;
}

// runtime tables (any facet rule)
int trh_advect(int k, int nc, const double* xy, const int* nbr, const int* nbr_e, int nq_cell, const double* tab_cell,
               int nq_facet, const double* tab_facet, const double* U, const double* q, double c0, const double* acc,
               double c1, double* out) {
  BY_K(k, k_tracer_adv<K>(xy, nbr, nbr_e, nc, nq_cell, tab_cell, nq_facet, tab_facet, U, q, c0, acc, c1, out))
  return 0;
}
}
