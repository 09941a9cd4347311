// Velocity-side kernels of the HDG timesteppers: BDM projection, the implicit advection/penalty
// operator f_impl, weak divergence, pressure gradient, trace reconstruction.  One thread per cell,
// SoA fields, neighbour data gathered through L2 (one-deep facet halo).
//
// Reference forms (src/timesteppers/):
//   project_bdm          common.py:59-70,91-108
//   _f_impl              hdg_imex.py:313-331   (Chorin operator hdg_implicit.py:103-125)
//   _weak_divergence     hdg_imex.py:353-365   (Chorin rhs hdg_implicit.py:145)
//   _pressure_gradient   hdg_imex.py:333-340
//   _reconstruct_trace   hdg_imex.py:450-469
//
// All velocity "dual vectors" are produced in Riesz form (multiplied by M^-1 = 1/detJ), which the
// orthonormal basis makes free: int_K w (.) dx = detJ * sum_q WQ[q] (.) phi(q).
#pragma once
#include "hdg_local.cuh"

// neighbour of `cell` across local facet e (or -1) and the neighbour's local facet index
__global__ void k_build_nbr(const int* __restrict__ cell_facet, const int* __restrict__ facet_cell,
                            const int* __restrict__ facet_local, int nc, int nf, int* __restrict__ nbr,
                            int* __restrict__ nbr_e) {
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    for (int e = 0; e < 3; ++e) {
      int f = cell_facet[(size_t)e * nc + cell];
      int c0 = facet_cell[f], c1 = facet_cell[(size_t)nf + f];
      int other = (c0 == cell) ? 1 : 0;
      int oc = other ? c1 : c0;
      nbr[(size_t)e * nc + cell] = oc;
      nbr_e[(size_t)e * nc + cell] = oc >= 0 ? facet_local[(size_t)other * nf + f] : -1;
    }
  }
}

template <int K, typename P, typename R>  // P: storage type of the field, R: register type (both deduced)
__device__ __forceinline__ void load_Q(const P* __restrict__ Q, int nc, int cell, R (&q)[2][Dims<K>::NQ1]) {
  constexpr int NQ1 = Dims<K>::NQ1;
  HDG_UNROLL
  for (int c = 0; c < 2; ++c)
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i) q[c][i] = Q[(size_t)(c * NQ1 + i) * nc + cell];
}

// ------------------------------------------------------------------------------------------------
// BDM projection, pass 1: flux moments mu[e][j] = |e| int_0^1 (Q.n_e) l_j ds  (cell-local
// parametrisation), written to the facet-side buffer fm[(side*(K+2)+j)*nf + f]
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(128) k_bdm_moments(const double* __restrict__ xy,
                                                     const int* __restrict__ cell_facet,
                                                     const int* __restrict__ facet_cell, int nc, int nf,
                                                     const double* __restrict__ Q, double* __restrict__ fm) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NJ = K + 2;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    double q[2][NQ1];
    load_Q<K>(Q, nc, cell, q);
    HDG_UNROLL
    for (int e = 0; e < 3; ++e) {
      int f = cell_facet[(size_t)e * nc + cell];
      int side = (facet_cell[f] == cell) ? 0 : 1;
      double cx = g.le[e] * g.n[e][0], cy = g.le[e] * g.n[e][1];
      double mu[NJ];
      HDG_UNROLL
      for (int j = 0; j < NJ; ++j) mu[j] = 0.0;
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) {
        double qn = cx * q[0][i] + cy * q[1][i];
        HDG_UNROLL
        for (int j = 0; j < NJ; ++j)
          if (T::BF(e, j, i) != 0.0) mu[j] = fma(T::BF(e, j, i), qn, mu[j]);
      }
      HDG_UNROLL
      for (int j = 0; j < NJ; ++j) fm[(size_t)(side * NJ + j) * nf + f] = mu[j];
    }
  }
}

// pass 2: average the moments (zero on the boundary, DirichletBC common.py:106-107), lift the
// change back with the constant reference table LIFT through the contravariant Piola map
template <int K>
__global__ void __launch_bounds__(128) k_bdm_lift(const double* __restrict__ xy, const int* __restrict__ cell_facet,
                                                  const int* __restrict__ facet_cell, int nc, int nf,
                                                  const double* __restrict__ Q, const double* __restrict__ fm,
                                                  double* __restrict__ Qs) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NJ = K + 2;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    double delta[3][NJ];
    HDG_UNROLL
    for (int e = 0; e < 3; ++e) {
      int f = cell_facet[(size_t)e * nc + cell];
      int c0 = facet_cell[f], c1 = facet_cell[(size_t)nf + f];
      int side = (c0 == cell) ? 0 : 1;
      bool interior = c1 >= 0;
      HDG_UNROLL
      for (int j = 0; j < NJ; ++j) {
        double own = fm[(size_t)(side * NJ + j) * nf + f];
        double d = -own;
        if (interior) {
          double other = fm[(size_t)((1 - side) * NJ + j) * nf + f];
          // the neighbour's functional has the opposite normal and reversed parametrisation
          d = 0.5 * (((j & 1) ? other : -other) - own);
        }
        delta[e][j] = d;
      }
    }
    // J = inverse of Ji
    double J00 = g.Ji[1][1] * g.detJ, J01 = -g.Ji[0][1] * g.detJ, J10 = -g.Ji[1][0] * g.detJ,
           J11 = g.Ji[0][0] * g.detJ;
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i) {
      double h0 = 0.0, h1 = 0.0;
      HDG_UNROLL
      for (int e = 0; e < 3; ++e)
        HDG_UNROLL
        for (int j = 0; j < NJ; ++j) {
          if (T::LIFT(0, i, e, j) != 0.0) h0 = fma(T::LIFT(0, i, e, j), delta[e][j], h0);
          if (T::LIFT(1, i, e, j) != 0.0) h1 = fma(T::LIFT(1, i, e, j), delta[e][j], h1);
        }
      Qs[(size_t)i * nc + cell] = Q[(size_t)i * nc + cell] + g.idetJ * (J00 * h0 + J01 * h1);
      Qs[(size_t)(NQ1 + i) * nc + cell] = Q[(size_t)(NQ1 + i) * nc + cell] + g.idetJ * (J10 * h0 + J11 * h1);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// traces of a velocity field on local facet E at the NQF Gauss points (cell-local parametrisation)
// ------------------------------------------------------------------------------------------------
template <int K, int E, typename R>
__device__ __forceinline__ void trace_at_points(const R (&x)[2][Dims<K>::NQ1], bool reversed,
                                                R (&out)[RefTables<K>::NQF][2]) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NQF = T::NQF;
  HDG_UNROLL
  for (int q = 0; q < NQF; ++q) {
    R v0 = 0, v1 = 0;
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i) {
      v0 = fma((R)T::PHIF(E, q, i), x[0][i], v0);
      v1 = fma((R)T::PHIF(E, q, i), x[1][i], v1);
    }
    out[q][0] = v0;
    out[q][1] = v1;
  }
  if (reversed) {  // Gauss points are symmetric: s_q -> 1 - s_q is q -> NQF-1-q
    HDG_UNROLL
    for (int q = 0; q < NQF / 2; ++q) {
      HDG_UNROLL
      for (int c = 0; c < 2; ++c) {
        R t = out[q][c];
        out[q][c] = out[NQF - 1 - q][c];
        out[NQF - 1 - q][c] = t;
      }
    }
  }
}

template <int K, typename P, typename R>
__device__ __forceinline__ void nbr_trace(const P* __restrict__ X, int nc, int nbr, int e2,
                                          R (&out)[RefTables<K>::NQF][2]) {
  R xn[2][Dims<K>::NQ1];
  load_Q<K>(X, nc, nbr, xn);
  switch (e2) {
    case 0: trace_at_points<K, 0>(xn, true, out); break;
    case 1: trace_at_points<K, 1>(xn, true, out); break;
    default: trace_at_points<K, 2>(xn, true, out); break;
  }
}

// ------------------------------------------------------------------------------------------------
// y = c0 * z + c1 * M^-1 f_impl(., x; Q*)     (hdg_imex.py:313-331), z = x unless Z is given
// Per cell K (outward normal n, s = Q*.n which is single valued for BDM Q*):
//   - int_K w_c (Q*.grad) x_c
//   + int_{dK int} (s/2 - [upwind]|s|) (x_K - x_nbr).w - alpha/h_F ((x_K - x_nbr).n)(w.n)
//   - int_{dK bnd} alpha/h_F (x.n)(w.n)
// R = arithmetic and storage type of the fields: double on every FP64 path; float inside the mixed-precision
// tentative-velocity solver (hdg_engine.cu, run_tentative_mixed), where the operator only acts on corrections.
// ------------------------------------------------------------------------------------------------
// s[q] = n_E . Q* at the facet points, from the pulled-back field Qh: Q* = J Qh  =>  n.Q*_i = (J^T n)_d Qh[d][i]
template <int K, int E, typename R>
__device__ __forceinline__ void facet_flux(const Geo& g, const R (&Qh)[2][Dims<K>::NQ1], R (&s)[RefTables<K>::NQF]) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NQF = T::NQF;
  double J00 = g.Ji[1][1] * g.detJ, J01 = -g.Ji[0][1] * g.detJ, J10 = -g.Ji[1][0] * g.detJ, J11 = g.Ji[0][0] * g.detJ;
  const R m0 = (R)(g.n[E][0] * J00 + g.n[E][1] * J10), m1 = (R)(g.n[E][0] * J01 + g.n[E][1] * J11);
  HDG_UNROLL
  for (int q = 0; q < NQF; ++q) {
    R v = 0;
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i) v = fma((R)T::PHIF(E, q, i), m0 * Qh[0][i] + m1 * Qh[1][i], v);
    s[q] = v;
  }
}

template <int K, bool UPWIND, int E, typename P, typename R>
__device__ __forceinline__ void fimpl_facet(const Geo& g, double alpha, int nc, int nbr, int nbr_e,
                                            const P* __restrict__ X, const R (&x)[2][Dims<K>::NQ1],
                                            const R (&sflux)[RefTables<K>::NQF], R (&acc)[2][Dims<K>::NQ1]) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NQF = T::NQF;
  const R nx = (R)g.n[E][0], ny = (R)g.n[E][1];
  const R scale = (R)(g.le[E] * g.idetJ);
  const R hf = (R)(alpha / g.le[E]);
  R xo[NQF][2];
  trace_at_points<K, E>(x, false, xo);
  R vec[NQF][2];
  if (nbr >= 0) {
    R xnb[NQF][2];
    nbr_trace<K>(X, nc, nbr, nbr_e, xnb);
    HDG_UNROLL
    for (int q = 0; q < NQF; ++q) {
      const R s = sflux[q];
      R j0 = xo[q][0] - xnb[q][0], j1 = xo[q][1] - xnb[q][1];
      R coef = (R)0.5 * s - (UPWIND ? fabs(s) : (R)0);
      R pen = hf * (j0 * nx + j1 * ny);
      R w = (R)T::WF(q) * scale;
      vec[q][0] = w * (coef * j0 - pen * nx);
      vec[q][1] = w * (coef * j1 - pen * ny);
    }
  } else {
    HDG_UNROLL
    for (int q = 0; q < NQF; ++q) {
      R pen = hf * (xo[q][0] * nx + xo[q][1] * ny);
      R w = (R)T::WF(q) * scale;
      vec[q][0] = -w * pen * nx;
      vec[q][1] = -w * pen * ny;
    }
  }
  HDG_UNROLL
  for (int q = 0; q < NQF; ++q)
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i) {
      acc[0][i] = fma((R)T::PHIF(E, q, i), vec[q][0], acc[0][i]);
      acc[1][i] = fma((R)T::PHIF(E, q, i), vec[q][1], acc[1][i]);
    }
}

template <int K, bool UPWIND, typename R = double>
__global__ void __launch_bounds__(128, (K <= 2 ? 3 : 1)) k_fimpl(const double* __restrict__ xy, const int* __restrict__ nbr,
                                               const int* __restrict__ nbr_e, int nc, double alpha,
                                               const R* __restrict__ Qstar, const R* __restrict__ X,
                                               const R* __restrict__ Z, R c0, R c1, R* __restrict__ Y) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NQ = T::NQ;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    R x[2][NQ1], Qh[2][NQ1], acc[2][NQ1];
    load_Q<K>(X, nc, cell, x);
    {
      R qs[2][NQ1];
      load_Q<K>(Qstar, nc, cell, qs);
      const R j00 = (R)g.Ji[0][0], j01 = (R)g.Ji[0][1], j10 = (R)g.Ji[1][0], j11 = (R)g.Ji[1][1];
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) {
        Qh[0][i] = j00 * qs[0][i] + j01 * qs[1][i];
        Qh[1][i] = j10 * qs[0][i] + j11 * qs[1][i];
      }
    }
    HDG_UNROLL
    for (int c = 0; c < 2; ++c)
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) acc[c][i] = 0;
    // volume term: -(1/detJ) int_K w_c (Q*.grad x_c) = -sum_q WQ[q] phi_i(q) (Qh . grad^ x_c)(q)
    HDG_UNROLL
    for (int q = 0; q < NQ; ++q) {
      R a0 = 0, a1 = 0, g00 = 0, g01 = 0, g10 = 0, g11 = 0;
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) {
        a0 = fma((R)T::PHI(q, i), Qh[0][i], a0);
        a1 = fma((R)T::PHI(q, i), Qh[1][i], a1);
        if (T::DPHI(0, q, i) != 0.0) {
          g00 = fma((R)T::DPHI(0, q, i), x[0][i], g00);
          g10 = fma((R)T::DPHI(0, q, i), x[1][i], g10);
        }
        if (T::DPHI(1, q, i) != 0.0) {
          g01 = fma((R)T::DPHI(1, q, i), x[0][i], g01);
          g11 = fma((R)T::DPHI(1, q, i), x[1][i], g11);
        }
      }
      R w = -(R)T::WQ(q);
      R v0 = w * (a0 * g00 + a1 * g01), v1 = w * (a0 * g10 + a1 * g11);
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) {
        acc[0][i] = fma((R)T::PHI(q, i), v0, acc[0][i]);
        acc[1][i] = fma((R)T::PHI(q, i), v1, acc[1][i]);
      }
    }
    {
      R sf[T::NQF];
      const int n0 = nbr[cell], n1 = nbr[(size_t)nc + cell], n2 = nbr[2 * (size_t)nc + cell];
      if (n0 >= 0) facet_flux<K, 0>(g, Qh, sf);
      fimpl_facet<K, UPWIND, 0>(g, alpha, nc, n0, nbr_e[cell], X, x, sf, acc);
      if (n1 >= 0) facet_flux<K, 1>(g, Qh, sf);
      fimpl_facet<K, UPWIND, 1>(g, alpha, nc, n1, nbr_e[(size_t)nc + cell], X, x, sf, acc);
      if (n2 >= 0) facet_flux<K, 2>(g, Qh, sf);
      fimpl_facet<K, UPWIND, 2>(g, alpha, nc, n2, nbr_e[2 * (size_t)nc + cell], X, x, sf, acc);
    }
    HDG_UNROLL
    for (int c = 0; c < 2; ++c)
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) {
        size_t idx = (size_t)(c * NQ1 + i) * nc + cell;
        Y[idx] = c0 * (Z ? Z[idx] : x[c][i]) + c1 * acc[c][i];
      }
  }
}

// ------------------------------------------------------------------------------------------------
// The advecting velocity Q* is fixed during a tentative-velocity solve (32+ operator applications), so everything
// k_fimpl derives from it is tabulated once per solve:
//   pre[(2 q + d) nc + cell]              = pulled-back Q* at volume point q, component d        (2 NQ values)
//   pre[(2 NQ + e NQF + q) nc + cell]     = n_e . Q* at point q of local facet e                 (3 NQF values)
// k_fimpl_q is k_fimpl reading that table instead of Q*: 470 of the 2 330 FMAs per cell (k = 2) and the 2 NQ1
// registers of the pulled-back field go away, at the price of 2 NQ + 3 NQF - 2 NQ1 more doubles read per cell.
// ------------------------------------------------------------------------------------------------
template <int K>
struct FimplPre {
  static constexpr int N = 2 * RefTables<K>::NQ + 3 * RefTables<K>::NQF;
};

template <int K>
__global__ void __launch_bounds__(128) k_fimpl_pre(const double* __restrict__ xy, int nc,
                                                   const double* __restrict__ Qstar, double* __restrict__ pre) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NQ = T::NQ, NQF = T::NQF;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    double qs[2][NQ1], Qh[2][NQ1];
    load_Q<K>(Qstar, nc, cell, qs);
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i) {
      Qh[0][i] = g.Ji[0][0] * qs[0][i] + g.Ji[0][1] * qs[1][i];
      Qh[1][i] = g.Ji[1][0] * qs[0][i] + g.Ji[1][1] * qs[1][i];
    }
    HDG_UNROLL
    for (int q = 0; q < NQ; ++q) {
      double a0 = 0.0, a1 = 0.0;
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) {
        a0 = fma(T::PHI(q, i), Qh[0][i], a0);
        a1 = fma(T::PHI(q, i), Qh[1][i], a1);
      }
      pre[(size_t)(2 * q) * nc + cell] = a0;
      pre[(size_t)(2 * q + 1) * nc + cell] = a1;
    }
    double sf[NQF];
    facet_flux<K, 0>(g, Qh, sf);
    HDG_UNROLL
    for (int q = 0; q < NQF; ++q) pre[(size_t)(2 * NQ + q) * nc + cell] = sf[q];
    facet_flux<K, 1>(g, Qh, sf);
    HDG_UNROLL
    for (int q = 0; q < NQF; ++q) pre[(size_t)(2 * NQ + NQF + q) * nc + cell] = sf[q];
    facet_flux<K, 2>(g, Qh, sf);
    HDG_UNROLL
    for (int q = 0; q < NQF; ++q) pre[(size_t)(2 * NQ + 2 * NQF + q) * nc + cell] = sf[q];
  }
}

template <int K, bool UPWIND>
__global__ void __launch_bounds__(128, (K <= 2 ? 3 : 1)) k_fimpl_q(const double* __restrict__ xy, const int* __restrict__ nbr,
                                                 const int* __restrict__ nbr_e, int nc, double alpha,
                                                 const double* __restrict__ pre, const double* __restrict__ X,
                                                 const double* __restrict__ Z, double c0, double c1,
                                                 double* __restrict__ Y) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NQ = T::NQ, NQF = T::NQF;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    double x[2][NQ1], acc[2][NQ1];
    load_Q<K>(X, nc, cell, x);
    HDG_UNROLL
    for (int c = 0; c < 2; ++c)
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) acc[c][i] = 0.0;
    HDG_UNROLL
    for (int q = 0; q < NQ; ++q) {
      const double a0 = pre[(size_t)(2 * q) * nc + cell], a1 = pre[(size_t)(2 * q + 1) * nc + cell];
      double g00 = 0.0, g01 = 0.0, g10 = 0.0, g11 = 0.0;
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) {
        if (T::DPHI(0, q, i) != 0.0) {
          g00 = fma(T::DPHI(0, q, i), x[0][i], g00);
          g10 = fma(T::DPHI(0, q, i), x[1][i], g10);
        }
        if (T::DPHI(1, q, i) != 0.0) {
          g01 = fma(T::DPHI(1, q, i), x[0][i], g01);
          g11 = fma(T::DPHI(1, q, i), x[1][i], g11);
        }
      }
      double w = -T::WQ(q);
      double v0 = w * (a0 * g00 + a1 * g01), v1 = w * (a0 * g10 + a1 * g11);
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) {
        acc[0][i] = fma(T::PHI(q, i), v0, acc[0][i]);
        acc[1][i] = fma(T::PHI(q, i), v1, acc[1][i]);
      }
    }
    {
      double sf[NQF];
      const int n0 = nbr[cell], n1 = nbr[(size_t)nc + cell], n2 = nbr[2 * (size_t)nc + cell];
      HDG_UNROLL
      for (int q = 0; q < NQF; ++q) sf[q] = pre[(size_t)(2 * NQ + q) * nc + cell];
      fimpl_facet<K, UPWIND, 0>(g, alpha, nc, n0, nbr_e[cell], X, x, sf, acc);
      HDG_UNROLL
      for (int q = 0; q < NQF; ++q) sf[q] = pre[(size_t)(2 * NQ + NQF + q) * nc + cell];
      fimpl_facet<K, UPWIND, 1>(g, alpha, nc, n1, nbr_e[(size_t)nc + cell], X, x, sf, acc);
      HDG_UNROLL
      for (int q = 0; q < NQF; ++q) sf[q] = pre[(size_t)(2 * NQ + 2 * NQF + q) * nc + cell];
      fimpl_facet<K, UPWIND, 2>(g, alpha, nc, n2, nbr_e[2 * (size_t)nc + cell], X, x, sf, acc);
    }
    HDG_UNROLL
    for (int c = 0; c < 2; ++c)
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) {
        size_t idx = (size_t)(c * NQ1 + i) * nc + cell;
        Y[idx] = c0 * (Z ? Z[idx] : x[c][i]) + c1 * acc[c][i];
      }
  }
}

// ------------------------------------------------------------------------------------------------
// The operator of the augmented tentative-velocity iteration is f_impl WITHOUT the normal-jump penalty (alpha = 0: the
// penalty lives in the multiplier rows, hdg_tent.cuh).  Without it the two velocity components do not couple: the
// advection form acts on x_0 and x_1 separately with the same scalar weights (Q* at the volume points, n.Q* at the facet
// points, both in the table of k_fimpl_pre).  k_fimpl_c therefore gives every (cell, component) its own thread:
//   * per thread 10 + 10 (+ 10 neighbour) doubles of state instead of 20 + 20 + 20 (k = 2): no spills, 4 instead of 3
//     resident CTAs, twice the threads (k_fimpl: 168 registers + 208 B of spills, 12 % occupancy, FP64 pipe 45 % active
//     in ncu r1h);
//   * 2 000 instead of 2 330 FMAs per cell at k = 2 (no Q* evaluation, no penalty).
// Lanes 0-15 of a warp take component 0 of 16 consecutive cells, lanes 16-31 component 1 of the same cells, so every
// half-warp reads 128 contiguous bytes of the SoA fields and both halves share the table lines.
//   Y = c0 Z + c1 M^-1 f_impl(., X; Q*)|_{alpha = 0}        (Z = X if null)
// Measured on a B200 at nx = 1024, k = 2 (in situ, profiles/r2/bench_r2t_*.json, bench_r2u_*.json): 547 us per launch
// against 627 us of k_fimpl in the same loop; ncu (profiles/r2/ncu_r2s_fimpl_c_*): DRAM traffic 1.93 GB = the algorithmic
// bytes, DRAM at 42 %, FP64 pipe at 44 %, issue slots 59 % busy, the chip at its power cap.  Three attempts to hide the
// load latency the stall sampling points at were all SLOWER and are not kept, except the last as an opt-in:
//   * reference tables from the constant bank (LDCU.128, 24 % fewer instructions than the UMOV pairs sm_100a needs for
//     every 64-bit immediate): 98.6 vs 94.8 ms per solve;
//   * prefetch.global.L1 of the neighbour coefficients (and of the facet rows of the table) as soon as the indices are
//     known: 604 vs 546 us per launch;
//   * the cell's own rows staged in shared memory by TMA bulk copies (k_fimpl_t below): 598 vs 547 us.
// ------------------------------------------------------------------------------------------------
template <int K, int E>
__device__ __forceinline__ void trace1_at_points(const double (&x)[Dims<K>::NQ1], bool reversed,
                                                 double (&out)[RefTables<K>::NQF]) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NQF = T::NQF;
  HDG_UNROLL
  for (int q = 0; q < NQF; ++q) {
    double v = 0.0;
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i)
      if (T::PHIF(E, q, i) != 0.0) v = fma(T::PHIF(E, q, i), x[i], v);
    out[reversed ? NQF - 1 - q : q] = v;  // Gauss points are symmetric: s_q -> 1 - s_q is q -> NQF-1-q
  }
}

// Src: where the cell's own rows come from -- global memory (FimplSrcGlobal) or the shared-memory tile that the bulk
// copies of k_fimpl_t filled (FimplSrcTile); pre(r) = table row r, x(i) / z(i) = coefficient i of this thread's component
struct FimplSrcGlobal {
  const double* __restrict__ pre;
  const double* __restrict__ Xc;
  const double* __restrict__ Zc;  // may be null: z = x
  size_t nc, cell;
  __device__ __forceinline__ double P(int r) const { return pre[(size_t)r * nc + cell]; }
  __device__ __forceinline__ double X(int i) const { return Xc[(size_t)i * nc + cell]; }
  __device__ __forceinline__ bool has_z() const { return Zc != nullptr; }
  __device__ __forceinline__ double Zv(int i) const { return Zc[(size_t)i * nc + cell]; }
};

template <int K, bool UPWIND, int E, class Src>
__device__ __forceinline__ void fimpl_c_facet(double scale, int nc, int nbr, int nbr_e, const double* __restrict__ Xc,
                                              const Src& src, int sfrow, const double (&x)[Dims<K>::NQ1],
                                              double (&acc)[Dims<K>::NQ1]) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NQF = T::NQF;
  if (nbr < 0) return;  // boundary facet: only the penalty acts there
  double xn[NQ1], sf[NQF];
  HDG_UNROLL
  for (int i = 0; i < NQ1; ++i) xn[i] = Xc[(size_t)i * nc + nbr];
  HDG_UNROLL
  for (int q = 0; q < NQF; ++q) sf[q] = src.P(sfrow + q);
  double xo[NQF], xnb[NQF];
  trace1_at_points<K, E>(x, false, xo);
  switch (nbr_e) {
    case 0: trace1_at_points<K, 0>(xn, true, xnb); break;
    case 1: trace1_at_points<K, 1>(xn, true, xnb); break;
    default: trace1_at_points<K, 2>(xn, true, xnb); break;
  }
  HDG_UNROLL
  for (int q = 0; q < NQF; ++q) {
    const double s = sf[q];
    const double coef = 0.5 * s - (UPWIND ? fabs(s) : 0.0);
    xo[q] = (T::WF(q) * scale) * (coef * (xo[q] - xnb[q]));
  }
  HDG_UNROLL
  for (int q = 0; q < NQF; ++q)
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i)
      if (T::PHIF(E, q, i) != 0.0) acc[i] = fma(T::PHIF(E, q, i), xo[q], acc[i]);
}

// one (cell, component): Yc[i] = c0 z_i + c1 [M^-1 f_impl(., x; Q*)]_i  for the NQ1 coefficients of the component
template <int K, bool UPWIND, class Src>
__device__ __forceinline__ void fimpl_c_item(const double* __restrict__ xy, const int* __restrict__ nbr,
                                             const int* __restrict__ nbr_e, int nc, size_t cell,
                                             const double* __restrict__ Xc, const Src& src, double c0, double c1,
                                             double* __restrict__ Yc) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NQ = T::NQ, NQF = T::NQF;
  // everything the facet terms need from global memory is requested first (indices, geometry)
  const int n0 = nbr[cell], n1 = nbr[(size_t)nc + cell], n2 = nbr[2 * (size_t)nc + cell];
  const int e0 = nbr_e[cell], e1 = nbr_e[(size_t)nc + cell], e2 = nbr_e[2 * (size_t)nc + cell];
  const double x0 = xy[cell], y0 = xy[(size_t)nc + cell], x1 = xy[2 * (size_t)nc + cell], y1 = xy[3 * (size_t)nc + cell],
               x2 = xy[4 * (size_t)nc + cell], y2 = xy[5 * (size_t)nc + cell];
  double x[NQ1], acc[NQ1];
  HDG_UNROLL
  for (int i = 0; i < NQ1; ++i) {
    x[i] = src.X(i);
    acc[i] = 0.0;
  }
  // volume term: -sum_q WQ[q] phi_i(q) (Qh . grad^ x_c)(q), Qh(q) from the table
  HDG_UNROLL
  for (int q = 0; q < NQ; ++q) {
    const double a0 = src.P(2 * q), a1 = src.P(2 * q + 1);
    double g0 = 0.0, g1 = 0.0;
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i) {
      if (T::DPHI(0, q, i) != 0.0) g0 = fma(T::DPHI(0, q, i), x[i], g0);
      if (T::DPHI(1, q, i) != 0.0) g1 = fma(T::DPHI(1, q, i), x[i], g1);
    }
    const double v = -T::WQ(q) * (a0 * g0 + a1 * g1);
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i)
      if (T::PHI(q, i) != 0.0) acc[i] = fma(T::PHI(q, i), v, acc[i]);
  }
  // facet terms: int_{dK int} (s/2 - [upwind]|s|) (x_K - x_nbr) w,  weight |e| / detJ
  {
    const double idetJ = 1.0 / ((x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0));
    // facet e runs from vertex (e+1)%3 to (e+2)%3 (make_geo)
    const double l0 = sqrt((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1));
    const double l1 = sqrt((x0 - x2) * (x0 - x2) + (y0 - y2) * (y0 - y2));
    const double l2 = sqrt((x1 - x0) * (x1 - x0) + (y1 - y0) * (y1 - y0));
    fimpl_c_facet<K, UPWIND, 0>(l0 * idetJ, nc, n0, e0, Xc, src, 2 * NQ, x, acc);
    fimpl_c_facet<K, UPWIND, 1>(l1 * idetJ, nc, n1, e1, Xc, src, 2 * NQ + NQF, x, acc);
    fimpl_c_facet<K, UPWIND, 2>(l2 * idetJ, nc, n2, e2, Xc, src, 2 * NQ + 2 * NQF, x, acc);
  }
  const bool hz = src.has_z();
  HDG_UNROLL
  for (int i = 0; i < NQ1; ++i) Yc[(size_t)i * nc + cell] = c0 * (hz ? src.Zv(i) : x[i]) + c1 * acc[i];
}

template <int K, bool UPWIND>
__global__ void __launch_bounds__(128, (K <= 2 ? 4 : (K == 3 ? 2 : 1)))
    k_fimpl_c(const double* __restrict__ xy, const int* __restrict__ nbr, const int* __restrict__ nbr_e, int nc,
              const double* __restrict__ pre, const double* __restrict__ X, const double* __restrict__ Z, double c0,
              double c1, double* __restrict__ Y) {
  constexpr int NQ1 = Dims<K>::NQ1;
  const long long nitems = 32LL * ((nc + 15) / 16);
  for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < nitems;
       item += (long long)gridDim.x * blockDim.x) {
    const int lane = (int)(item & 31);
    const int c = lane >> 4;
    const long long cell_ll = (item >> 5) * 16 + (lane & 15);
    if (cell_ll >= nc) continue;
    const size_t cell = (size_t)cell_ll;
    const double* __restrict__ Xc = X + (size_t)c * NQ1 * nc;
    const FimplSrcGlobal src{pre, Xc, Z ? Z + (size_t)c * NQ1 * nc : nullptr, (size_t)nc, cell};
    fimpl_c_item<K, UPWIND>(xy, nbr, nbr_e, nc, cell, Xc, src, c0, c1, Y + (size_t)c * NQ1 * nc);
  }
}

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// k_fimpl_t (opt-in, hdg_set_tuning("fimpl_split", 3); k <= 2): k_fimpl_c with the cell's own rows staged by the TMA.
// One CTA = 64 consecutive cells x 2 components; its rows of the table, of x and of z are 512 contiguous bytes each in
// the SoA layout, so warp 0 hands all of them to the copy engine at once (cp.async.bulk global -> shared, completion
// counted in bytes on an mbarrier) and the threads read them from shared memory: every byte of the tile is in flight
// from the first cycle and none of it occupies a register.  Only the neighbour cells' coefficients (gathers) remain
// ordinary loads.  Needs nc even (16-byte aligned rows) and full tiles; the last, partial tile takes the global path.
// Parity-green on the GPU (profiles/r2/pytest_tma_r2t.log) but 9 % slower than k_fimpl_c (598 vs 547 us in situ): all
// threads of a CTA wait for the whole tile before the first FMA, whereas the row-by-row loads of k_fimpl_c let the warps
// of a CTA drift apart and overlap their load and FMA phases; hence not the default.
// ------------------------------------------------------------------------------------------------
struct FimplSrcTile {
  const double* tile;  // [rows][64] in shared memory: table rows, then x (2 NQ1 rows), then z (2 NQ1 rows, optional)
  int j;               // cell within the tile
  int xrow, zrow;      // first row of this thread's component of x / z (zrow < 0: z = x)
  __device__ __forceinline__ double P(int r) const { return tile[r * 64 + j]; }
  __device__ __forceinline__ double X(int i) const { return tile[(xrow + i) * 64 + j]; }
  __device__ __forceinline__ bool has_z() const { return zrow >= 0; }
  __device__ __forceinline__ double Zv(int i) const { return tile[(zrow + i) * 64 + j]; }
};

template <int K>
struct FimplTile {
  static constexpr int NPRE = 2 * RefTables<K>::NQ + 3 * RefTables<K>::NQF, NX = 2 * Dims<K>::NQ1;
  static constexpr int ROWS = NPRE + 2 * NX;
  static constexpr size_t SMEM = (size_t)ROWS * 64 * sizeof(double) + 16;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int K, bool UPWIND>
__global__ void __launch_bounds__(128, (K <= 2 ? 4 : 1))
    k_fimpl_t(const double* __restrict__ xy, const int* __restrict__ nbr, const int* __restrict__ nbr_e, int nc,
              const double* __restrict__ pre, const double* __restrict__ X, const double* __restrict__ Z, double c0,
              double c1, double* __restrict__ Y) {
  constexpr int NQ1 = Dims<K>::NQ1, NPRE = FimplTile<K>::NPRE, NX = FimplTile<K>::NX;
  extern __shared__ __align__(128) unsigned char fimpl_smem[];
  double* tile = reinterpret_cast<double*>(fimpl_smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(fimpl_smem + (size_t)FimplTile<K>::ROWS * 64 * sizeof(double));
  const int cell0 = blockIdx.x * 64;
  const bool staged = cell0 + 64 <= nc;  // block-uniform
  const int nrows = NPRE + NX + (Z ? NX : 0);
  if (staged) {
    const uint32_t bar_a = smem_u32(bar);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      if (threadIdx.x == 0)
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(nrows * 512) : "memory");
      __syncwarp();
      for (int r = threadIdx.x; r < nrows; r += 32) {
        const double* src = r < NPRE        ? pre + (size_t)r * nc + cell0
                            : r < NPRE + NX ? X + (size_t)(r - NPRE) * nc + cell0
                                            : Z + (size_t)(r - NPRE - NX) * nc + cell0;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(tile + (size_t)r * 64)),
                     "l"(src), "r"(512), "r"(bar_a)
                     : "memory");
      }
    }
  }
  const int lane = threadIdx.x & 31, c = lane >> 4, j = (threadIdx.x >> 5) * 16 + (lane & 15);
  const long long cell_ll = (long long)cell0 + j;
  const double* __restrict__ Xc = X + (size_t)c * NQ1 * nc;
  double* __restrict__ Yc = Y + (size_t)c * NQ1 * nc;
  if (staged) {
    // wait for the bytes (phase 0 of the barrier); a copy that never completes traps instead of hanging the device
    const uint32_t bar_a = smem_u32(bar);
    const long long t0 = clock64();
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(bar_a), "r"(0)
          : "memory");
      if (!done && clock64() - t0 > 4000000000LL) __trap();
    }
    const FimplSrcTile src{tile, j, NPRE + c * NQ1, Z ? NPRE + NX + c * NQ1 : -1};
    fimpl_c_item<K, UPWIND>(xy, nbr, nbr_e, nc, (size_t)cell_ll, Xc, src, c0, c1, Yc);
  } else if (cell_ll < nc) {
    const FimplSrcGlobal src{pre, Xc, Z ? Z + (size_t)c * NQ1 * nc : nullptr, (size_t)nc, (size_t)cell_ll};
    fimpl_c_item<K, UPWIND>(xy, nbr, nbr_e, nc, (size_t)cell_ll, Xc, src, c0, c1, Yc);
  }
}
#endif  // __CUDACC__

// ------------------------------------------------------------------------------------------------
// weak divergence as a dual vector on the pressure space:
//   mode 0:  Rp = scale * int_K psi div Q                                  (hdg_implicit.py:145)
//   mode 1:  Rp = scale * _weak_divergence(psi, Q)                          (hdg_imex.py:353-365)
//            = int_K psi div Q - 1/2 int_{dK int} psi n.(Q_K - Q_nbr) - int_{dK bnd} psi n.Q
// ------------------------------------------------------------------------------------------------
template <int K, int E>
__device__ __forceinline__ void wdiv_facet(const Geo& g, int nc, int nbr, int nbr_e, const double* __restrict__ Q,
                                           const double (&q)[2][Dims<K>::NQ1], double scale,
                                           double (&out)[Dims<K>::NP]) {
  using T = RefTables<K>;
  constexpr int NP = Dims<K>::NP, NQF = T::NQF;
  double xo[NQF][2];
  trace_at_points<K, E>(q, false, xo);
  double fac = 1.0;
  if (nbr >= 0) {
    double xnb[NQF][2];
    nbr_trace<K>(Q, nc, nbr, nbr_e, xnb);
    HDG_UNROLL
    for (int p = 0; p < NQF; ++p) {
      xo[p][0] -= xnb[p][0];
      xo[p][1] -= xnb[p][1];
    }
    fac = 0.5;
  }
  HDG_UNROLL
  for (int p = 0; p < NQF; ++p) {
    double v = -scale * fac * g.le[E] * T::WF(p) * (g.n[E][0] * xo[p][0] + g.n[E][1] * xo[p][1]);
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) out[a] = fma(T::PSIF(E, p, a), v, out[a]);
  }
}

template <int K>
__global__ void __launch_bounds__(128) k_weak_div(const double* __restrict__ xy, const int* __restrict__ nbr,
                                                  const int* __restrict__ nbr_e, int nc,
                                                  const double* __restrict__ Q, double scale, int mode,
                                                  double* __restrict__ Rp) {
  constexpr int NQ1 = Dims<K>::NQ1, NP = Dims<K>::NP;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    double q[2][NQ1], out[NP];
    load_Q<K>(Q, nc, cell, q);
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) out[a] = 0.0;
    apply_B_over_detJ<K>(g, q, scale * g.detJ, out);
    if (mode == 1) {
      wdiv_facet<K, 0>(g, nc, nbr[cell], nbr_e[cell], Q, q, scale, out);
      wdiv_facet<K, 1>(g, nc, nbr[(size_t)nc + cell], nbr_e[(size_t)nc + cell], Q, q, scale, out);
      wdiv_facet<K, 2>(g, nc, nbr[2 * (size_t)nc + cell], nbr_e[2 * (size_t)nc + cell], Q, q, scale, out);
    }
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) Rp[(size_t)a * nc + cell] = out[a];
  }
}

// ------------------------------------------------------------------------------------------------
// Y = c0 * Y + c1 * M^-1 g(w, p, lambda),   g = B^T p - E^T lambda   (hdg_imex.py:333-340)
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(128) k_pgrad(const double* __restrict__ xy, const int* __restrict__ flip,
                                               const int* __restrict__ cell_facet, int nc, int nf,
                                               const double* __restrict__ p, const double* __restrict__ lamg,
                                               double c0, double c1, double* __restrict__ Y) {
  using D = Dims<K>;
  constexpr int NQ1 = D::NQ1, NP = D::NP, NL1 = D::NL1;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    double u[2][NQ1], phi[NP], lam[3][NL1];
    HDG_UNROLL
    for (int e = 0; e < 3; ++e) {
      int f = cell_facet[(size_t)e * nc + cell];
      int fl = flip[(size_t)e * nc + cell];
      HDG_UNROLL
      for (int m = 0; m < NL1; ++m) lam[e][m] = flip_sign(fl, m) * lamg[(size_t)m * nf + f];
    }
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) phi[a] = p[(size_t)a * nc + cell];
    HDG_UNROLL
    for (int c = 0; c < 2; ++c)
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) u[c][i] = 0.0;
    apply_Bt_over_detJ<K>(g, phi, 1.0, u);
    apply_Et<K>(g, lam, -g.idetJ, u);
    HDG_UNROLL
    for (int c = 0; c < 2; ++c)
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) {
        size_t idx = (size_t)(c * NQ1 + i) * nc + cell;
        Y[idx] = (c0 != 0.0 ? c0 * Y[idx] : 0.0) + c1 * u[c][i];
      }
  }
}

// ------------------------------------------------------------------------------------------------
// _reconstruct_trace (hdg_imex.py:450-469), pass 1 per cell: gK[e][m] = (E Q + tau F p)[e][m] in
// the global facet orientation; pass 2 (k_trace_avg) divides by tau * multiplicity * |F|.
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(128) k_trace_moments(const double* __restrict__ xy, const int* __restrict__ flip,
                                                       int nc, double tau, const double* __restrict__ Q,
                                                       const double* __restrict__ p, double* __restrict__ gK) {
  using D = Dims<K>;
  constexpr int NQ1 = D::NQ1, NP = D::NP, NL1 = D::NL1;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    double u[2][NQ1], phi[NP], lam[3][NL1];
    load_Q<K>(Q, nc, cell, u);
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) phi[a] = p[(size_t)a * nc + cell];
    HDG_UNROLL
    for (int e = 0; e < 3; ++e)
      HDG_UNROLL
      for (int m = 0; m < NL1; ++m) lam[e][m] = 0.0;
    apply_E<K>(g, u, 1.0, lam);
    apply_F<K>(g, phi, tau, lam);
    HDG_UNROLL
    for (int e = 0; e < 3; ++e) {
      int fl = flip[(size_t)e * nc + cell];
      double s = 1.0 / (tau * g.le[e]);
      HDG_UNROLL
      for (int m = 0; m < NL1; ++m) gK[(size_t)(e * NL1 + m) * nc + cell] = flip_sign(fl, m) * lam[e][m] * s;
    }
  }
}

template <int K>
__global__ void __launch_bounds__(256) k_trace_avg(const double* __restrict__ gK, const int* __restrict__ facet_cell,
                                                   const int* __restrict__ facet_local, int nc, int nf,
                                                   double* __restrict__ lam) {
  constexpr int NL1 = Dims<K>::NL1;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    int c0 = facet_cell[f], c1 = facet_cell[(size_t)nf + f];
    int e0 = facet_local[f], e1 = facet_local[(size_t)nf + f];
    HDG_UNROLL
    for (int m = 0; m < NL1; ++m) {
      double v = gK[(size_t)(e0 * NL1 + m) * nc + c0];
      if (c1 >= 0) v = 0.5 * (v + gK[(size_t)(e1 * NL1 + m) * nc + c1]);
      lam[(size_t)m * nf + f] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Constraint rows of the mixed operator applied to a state (the monolithic residual of the fully
// implicit stage, hdg_imex.py:602-610 / hdg_implicit.py:172-183):  Gamma(psi, mu; u, phi, lambda)
// (hdg_imex.py:342-351) as dual vectors
//   Rp = B u + T phi - tau F^T lambda                      (per cell)
//   Rl = sum_{K in f} (E u + tau F phi - tau |f| lambda)    (per facet; pass 2 = k_facet_sum)
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(128) k_gamma_cell(const double* __restrict__ xy, const int* __restrict__ flip,
                                                    const int* __restrict__ cell_facet, int nc, int nf, double tau,
                                                    const double* __restrict__ Q, const double* __restrict__ p,
                                                    const double* __restrict__ lamg, double* __restrict__ Rp,
                                                    double* __restrict__ gK) {
  using T = RefTables<K>;
  using D = Dims<K>;
  constexpr int NQ1 = D::NQ1, NP = D::NP, NL1 = D::NL1;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    double u[2][NQ1], phi[NP], lam[3][NL1], out[3][NL1], rp[NP];
    load_Q<K>(Q, nc, cell, u);
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) {
      phi[a] = p[(size_t)a * nc + cell];
      rp[a] = 0.0;
    }
    int fl[3];
    HDG_UNROLL
    for (int e = 0; e < 3; ++e) {
      int f = cell_facet[(size_t)e * nc + cell];
      fl[e] = flip[(size_t)e * nc + cell];
      HDG_UNROLL
      for (int m = 0; m < NL1; ++m) {
        lam[e][m] = flip_sign(fl[e], m) * lamg[(size_t)m * nf + f];
        out[e][m] = 0.0;
      }
    }
    apply_B_over_detJ<K>(g, u, g.detJ, rp);
    HDG_UNROLL
    for (int e = 0; e < 3; ++e) {
      double c = tau * g.le[e];
      HDG_UNROLL
      for (int a = 0; a < NP; ++a) {
        double v = 0.0;
        HDG_UNROLL
        for (int b = 0; b < NP; ++b) {
          const double t = (b <= a) ? T::TT(e, a, b) : T::TT(e, b, a);
          if (t != 0.0) v = fma(t, phi[b], v);
        }
        rp[a] = fma(c, v, rp[a]);
      }
    }
    apply_Ft<K>(g, lam, -tau, rp);
    apply_E<K>(g, u, 1.0, out);
    apply_F<K>(g, phi, tau, out);
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) Rp[(size_t)a * nc + cell] = rp[a];
    HDG_UNROLL
    for (int e = 0; e < 3; ++e)
      HDG_UNROLL
      for (int m = 0; m < NL1; ++m)
        gK[(size_t)(e * NL1 + m) * nc + cell] = flip_sign(fl[e], m) * (out[e][m] - tau * g.le[e] * lam[e][m]);
  }
}

template <int K>
__global__ void __launch_bounds__(256) k_facet_sum(const double* __restrict__ gK, const int* __restrict__ facet_cell,
                                                   const int* __restrict__ facet_local, int nc, int nf,
                                                   double* __restrict__ out) {
  constexpr int NL1 = Dims<K>::NL1;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    int c0 = facet_cell[f], c1 = facet_cell[(size_t)nf + f];
    int e0 = facet_local[f], e1 = facet_local[(size_t)nf + f];
    HDG_UNROLL
    for (int m = 0; m < NL1; ++m) {
      double v = gK[(size_t)(e0 * NL1 + m) * nc + c0];
      if (c1 >= 0) v += gK[(size_t)(e1 * NL1 + m) * nc + c1];
      out[(size_t)m * nf + f] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// pressure-reconstruction right-hand side (hdg_imex.py:204-207):
//   Rp = _weak_divergence(psi, X),  X = -b + (grad Q) Q;      Rl = - mu n.b ds
// integrated by parts (identity for piecewise polynomials):
//   Rp = - int_K grad psi . X + int_{dK int} psi n . avg(X)
// ------------------------------------------------------------------------------------------------
template <int K, int E>
__device__ __forceinline__ void recon_X_facet(const Geo& g, const double (&q)[2][Dims<K>::NQ1],
                                              const double (&b)[2][Dims<K>::NQ1], bool reversed,
                                              double (&X)[RefTables<K>::NQF][2]) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NQF = T::NQF;
  HDG_UNROLL
  for (int p = 0; p < NQF; ++p) {
    double q0 = 0, q1 = 0, b0 = 0, b1 = 0, g00 = 0, g01 = 0, g10 = 0, g11 = 0;
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i) {
      q0 = fma(T::PHIF(E, p, i), q[0][i], q0);
      q1 = fma(T::PHIF(E, p, i), q[1][i], q1);
      b0 = fma(T::PHIF(E, p, i), b[0][i], b0);
      b1 = fma(T::PHIF(E, p, i), b[1][i], b1);
      if (T::DPHIF(E, 0, p, i) != 0.0) {
        g00 = fma(T::DPHIF(E, 0, p, i), q[0][i], g00);
        g10 = fma(T::DPHIF(E, 0, p, i), q[1][i], g10);
      }
      if (T::DPHIF(E, 1, p, i) != 0.0) {
        g01 = fma(T::DPHIF(E, 1, p, i), q[0][i], g01);
        g11 = fma(T::DPHIF(E, 1, p, i), q[1][i], g11);
      }
    }
    double h0 = g.Ji[0][0] * q0 + g.Ji[0][1] * q1, h1 = g.Ji[1][0] * q0 + g.Ji[1][1] * q1;
    X[p][0] = -b0 + h0 * g00 + h1 * g01;
    X[p][1] = -b1 + h0 * g10 + h1 * g11;
  }
  if (reversed) {
    HDG_UNROLL
    for (int p = 0; p < NQF / 2; ++p)
      HDG_UNROLL
      for (int c = 0; c < 2; ++c) {
        double t = X[p][c];
        X[p][c] = X[NQF - 1 - p][c];
        X[NQF - 1 - p][c] = t;
      }
  }
}

template <int K, int E>
__device__ __forceinline__ void recon_facet(const Geo& g, const double* __restrict__ xy, int nc, int nf, int cell,
                                            int nbr, int nbr_e, int f, int fl, const double* __restrict__ Q,
                                            const double* __restrict__ B, const double (&q)[2][Dims<K>::NQ1],
                                            const double (&b)[2][Dims<K>::NQ1], double (&out)[Dims<K>::NP],
                                            double* __restrict__ Rl) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NP = Dims<K>::NP, NQF = T::NQF, NL1 = Dims<K>::NL1;
  if (nbr >= 0) {
    double X[NQF][2], Xn[NQF][2];
    recon_X_facet<K, E>(g, q, b, false, X);
    {
      Geo gn = make_geo(xy, nc, nbr);
      double qn[2][NQ1], bn[2][NQ1];
      load_Q<K>(Q, nc, nbr, qn);
      load_Q<K>(B, nc, nbr, bn);
      switch (nbr_e) {
        case 0: recon_X_facet<K, 0>(gn, qn, bn, true, Xn); break;
        case 1: recon_X_facet<K, 1>(gn, qn, bn, true, Xn); break;
        default: recon_X_facet<K, 2>(gn, qn, bn, true, Xn); break;
      }
    }
    HDG_UNROLL
    for (int p = 0; p < NQF; ++p) {
      double v = 0.5 * g.le[E] * T::WF(p) *
                 (g.n[E][0] * (X[p][0] + Xn[p][0]) + g.n[E][1] * (X[p][1] + Xn[p][1]));
      HDG_UNROLL
      for (int a = 0; a < NP; ++a) out[a] = fma(T::PSIF(E, p, a), v, out[a]);
    }
  } else {
    // boundary facet: Rl = - |e| int mu n.b ds   (this cell is the only writer)
    double bt[NQF][2];
    trace_at_points<K, E>(b, false, bt);
    double r[NL1];
    HDG_UNROLL
    for (int m = 0; m < NL1; ++m) r[m] = 0.0;
    HDG_UNROLL
    for (int p = 0; p < NQF; ++p) {
      double v = -g.le[E] * T::WF(p) * (g.n[E][0] * bt[p][0] + g.n[E][1] * bt[p][1]);
      HDG_UNROLL
      for (int m = 0; m < NL1; ++m) r[m] = fma(T::LEG(m, p), v, r[m]);
    }
    HDG_UNROLL
    for (int m = 0; m < NL1; ++m) Rl[(size_t)m * nf + f] = flip_sign(fl, m) * r[m];
  }
}

template <int K>
__global__ void __launch_bounds__(128) k_recon_rhs(const double* __restrict__ xy, const int* __restrict__ nbr,
                                                   const int* __restrict__ nbr_e, const int* __restrict__ cell_facet,
                                                   const int* __restrict__ flip, int nc, int nf,
                                                   const double* __restrict__ Q, const double* __restrict__ B,
                                                   double* __restrict__ Rp, double* __restrict__ Rl) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NP = Dims<K>::NP, NQ = T::NQ;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    double q[2][NQ1], b[2][NQ1], out[NP];
    load_Q<K>(Q, nc, cell, q);
    load_Q<K>(B, nc, cell, b);
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) out[a] = 0.0;
    HDG_UNROLL
    for (int p = 0; p < NQ; ++p) {
      double q0 = 0, q1 = 0, b0 = 0, b1 = 0, g00 = 0, g01 = 0, g10 = 0, g11 = 0;
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) {
        q0 = fma(T::PHI(p, i), q[0][i], q0);
        q1 = fma(T::PHI(p, i), q[1][i], q1);
        b0 = fma(T::PHI(p, i), b[0][i], b0);
        b1 = fma(T::PHI(p, i), b[1][i], b1);
        if (T::DPHI(0, p, i) != 0.0) {
          g00 = fma(T::DPHI(0, p, i), q[0][i], g00);
          g10 = fma(T::DPHI(0, p, i), q[1][i], g10);
        }
        if (T::DPHI(1, p, i) != 0.0) {
          g01 = fma(T::DPHI(1, p, i), q[0][i], g01);
          g11 = fma(T::DPHI(1, p, i), q[1][i], g11);
        }
      }
      double h0 = g.Ji[0][0] * q0 + g.Ji[0][1] * q1, h1 = g.Ji[1][0] * q0 + g.Ji[1][1] * q1;
      double X0 = -b0 + h0 * g00 + h1 * g01, X1 = -b1 + h0 * g10 + h1 * g11;
      // pulled-back X and the reference gradient of psi
      double w = -g.detJ * T::WQ(p);
      double Xh0 = w * (g.Ji[0][0] * X0 + g.Ji[0][1] * X1), Xh1 = w * (g.Ji[1][0] * X0 + g.Ji[1][1] * X1);
      HDG_UNROLL
      for (int a = 0; a < NP; ++a) {
        if (T::DPSI(0, p, a) != 0.0) out[a] = fma(T::DPSI(0, p, a), Xh0, out[a]);
        if (T::DPSI(1, p, a) != 0.0) out[a] = fma(T::DPSI(1, p, a), Xh1, out[a]);
      }
    }
    recon_facet<K, 0>(g, xy, nc, nf, cell, nbr[cell], nbr_e[cell], cell_facet[cell], flip[cell], Q, B, q, b, out, Rl);
    recon_facet<K, 1>(g, xy, nc, nf, cell, nbr[(size_t)nc + cell], nbr_e[(size_t)nc + cell],
                      cell_facet[(size_t)nc + cell], flip[(size_t)nc + cell], Q, B, q, b, out, Rl);
    recon_facet<K, 2>(g, xy, nc, nf, cell, nbr[2 * (size_t)nc + cell], nbr_e[2 * (size_t)nc + cell],
                      cell_facet[2 * (size_t)nc + cell], flip[2 * (size_t)nc + cell], Q, B, q, b, out, Rl);
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) Rp[(size_t)a * nc + cell] = out[a];
  }
}

// mass-weighted inner product of two cell fields: partial sums of detJ * x . y
__global__ void __launch_bounds__(256) k_l2_inner(const double* __restrict__ xy, int nc, int nc_own, int ndof,
                                                  const double* __restrict__ x, const double* __restrict__ y,
                                                  double* __restrict__ partial) {
  double acc = 0.0;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc_own; cell += gridDim.x * blockDim.x) {
    double x0 = xy[cell], y0 = xy[(size_t)nc + cell];
    double x1 = xy[2 * (size_t)nc + cell], y1 = xy[3 * (size_t)nc + cell];
    double x2 = xy[4 * (size_t)nc + cell], y2 = xy[5 * (size_t)nc + cell];
    double d = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
    double s = 0.0;
    for (int i = 0; i < ndof; ++i) s = fma(x[(size_t)i * nc + cell], y[(size_t)i * nc + cell], s);
    acc = fma(d, s, acc);
  }
  __shared__ double sm[8];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < 8 ? sm[threadIdx.x] : 0.0;
    for (int o = 4; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// small fused vector kernels
// ------------------------------------------------------------------------------------------------
struct LinComb {
  int n;
  double c[8];
  const double* x[8];
};
__global__ void k_lincomb(size_t len, LinComb lc, double* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
    double v = 0.0;
    for (int t = 0; t < lc.n; ++t) v = fma(lc.c[t], lc.x[t][i], v);
    out[i] = v;
  }
}

// y = detJ^(+-1) * x for a cell field with ndof dofs per cell
__global__ void k_mass(const double* __restrict__ xy, int nc, int ndof, int inverse, const double* __restrict__ x,
                       double* __restrict__ y) {
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    double x0 = xy[cell], y0 = xy[(size_t)nc + cell];
    double x1 = xy[2 * (size_t)nc + cell], y1 = xy[3 * (size_t)nc + cell];
    double x2 = xy[4 * (size_t)nc + cell], y2 = xy[5 * (size_t)nc + cell];
    double d = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
    if (inverse) d = 1.0 / d;
    for (int i = 0; i < ndof; ++i) y[(size_t)i * nc + cell] = d * x[(size_t)i * nc + cell];
  }
}
