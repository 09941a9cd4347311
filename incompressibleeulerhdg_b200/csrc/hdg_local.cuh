// Per-cell local algebra of the HDG mixed-Poisson operator, closed form for affine triangles.
//
// Reference forms: a_mixed_poisson  src/timesteppers/hdg_imex.py:123-127 with
// _pressure_gradient :333-340 and _Gamma :342-351 (written out again at hdg_implicit.py:133-143).
// Local block structure, rows (w,psi,mu) x columns (u,phi,lambda), SURVEY.md §8 a1:
//
//        [  M   -B^T    E^T   ]        M = detJ * I            (orthonormal Dubiner basis)
//        [  B     T   -tau F^T ]       B = detJ * sum_d Jinv[d][c] D_d
//        [  E   tau F  -tau G  ]       E_e = |e| n_e (x) Ehat_e,  F_e = |e| Fhat_e,  G_e = |e| I
//
// Where Slate factorises the dense nA x nA block A_K = [[M,-B^T],[B,T]] with Eigen PartialPivLU
// for every cell, the orthonormal basis gives M = detJ*I, so A_K^-1 reduces to one SPD NP x NP
// Cholesky of  H = T + B B^T / detJ  and every other product is a contraction of compile-time
// reference tensors (hdg_tables.inc) with a handful of geometric scalars.
#pragma once
#include "hdg_tables.inc"

#define HDG_UNROLL _Pragma("unroll")

template <int K>
struct Dims {
  using T = RefTables<K>;
  static constexpr int NQ1 = T::NQ1, NP = T::NP, NL1 = T::NL1;
  static constexpr int NQ = 2 * NQ1, NL = 3 * NL1, NA = NQ + NP;
  static constexpr int NH = NP * (NP + 1) / 2;
};

struct Geo {
  double detJ, idetJ;
  double Ji[2][2];  // Ji[d][c] = d xi_d / d x_c
  double le[3];     // facet lengths
  double n[3][2];   // outward unit normals
};

__device__ __forceinline__ Geo make_geo(const double* __restrict__ xy, int nc, int cell) {
  // xy is SoA: xy[(v*2+c)*nc + cell]
  double x0 = xy[0 * nc + cell], y0 = xy[1 * (size_t)nc + cell];
  double x1 = xy[2 * (size_t)nc + cell], y1 = xy[3 * (size_t)nc + cell];
  double x2 = xy[4 * (size_t)nc + cell], y2 = xy[5 * (size_t)nc + cell];
  Geo g;
  double J00 = x1 - x0, J01 = x2 - x0, J10 = y1 - y0, J11 = y2 - y0;
  g.detJ = J00 * J11 - J01 * J10;
  g.idetJ = 1.0 / g.detJ;
  g.Ji[0][0] = J11 * g.idetJ;
  g.Ji[0][1] = -J01 * g.idetJ;
  g.Ji[1][0] = -J10 * g.idetJ;
  g.Ji[1][1] = J00 * g.idetJ;
  // facet e runs from vertex (e+1)%3 to (e+2)%3
  double tx[3] = {x2 - x1, x0 - x2, x1 - x0};
  double ty[3] = {y2 - y1, y0 - y2, y1 - y0};
  HDG_UNROLL
  for (int e = 0; e < 3; ++e) {
    double l = sqrt(tx[e] * tx[e] + ty[e] * ty[e]);
    g.le[e] = l;
    double il = 1.0 / l;
    g.n[e][0] = ty[e] * il;
    g.n[e][1] = -tx[e] * il;
  }
  return g;
}

// sign of Legendre mode m under reversal of the facet parametrisation
__device__ __forceinline__ double flip_sign(int flip, int m) { return (flip && (m & 1)) ? -1.0 : 1.0; }

// ---- packed symmetric NP x NP helpers (lower triangle, row major: idx(a,b) = a(a+1)/2 + b, b <= a)
__host__ __device__ constexpr int tri(int a, int b) { return a * (a + 1) / 2 + b; }

template <int K>
__device__ __forceinline__ void build_H(const Geo& g, double tau, double (&H)[Dims<K>::NH]) {
  using T = RefTables<K>;
  constexpr int NP = Dims<K>::NP;
  // metric of the inverse map: gm[d][d'] = sum_c Ji[d][c] Ji[d'][c]
  double g00 = g.Ji[0][0] * g.Ji[0][0] + g.Ji[0][1] * g.Ji[0][1];
  double g01 = g.Ji[0][0] * g.Ji[1][0] + g.Ji[0][1] * g.Ji[1][1];
  double g11 = g.Ji[1][0] * g.Ji[1][0] + g.Ji[1][1] * g.Ji[1][1];
  double c0 = g.detJ * g00, c1 = g.detJ * g01, c2 = g.detJ * g11;
  double t0 = tau * g.le[0], t1 = tau * g.le[1], t2 = tau * g.le[2];
  HDG_UNROLL
  for (int a = 0; a < NP; ++a) {
    HDG_UNROLL
    for (int b = 0; b <= a; ++b) {
      double h = 0.0;
      if (T::KK(0, a, b) != 0.0) h = fma(c0, T::KK(0, a, b), h);
      if (T::KK(1, a, b) != 0.0) h = fma(c1, T::KK(1, a, b), h);
      if (T::KK(2, a, b) != 0.0) h = fma(c2, T::KK(2, a, b), h);
      if (T::TT(0, a, b) != 0.0) h = fma(t0, T::TT(0, a, b), h);
      if (T::TT(1, a, b) != 0.0) h = fma(t1, T::TT(1, a, b), h);
      if (T::TT(2, a, b) != 0.0) h = fma(t2, T::TT(2, a, b), h);
      H[tri(a, b)] = h;
    }
  }
}

// in-place Cholesky H = L L^T (lower, packed); the diagonal is stored as 1/L_aa
template <int N>
__device__ __forceinline__ void cholesky(double (&H)[N * (N + 1) / 2]) {
  HDG_UNROLL
  for (int j = 0; j < N; ++j) {
    double d = H[tri(j, j)];
    HDG_UNROLL
    for (int k = 0; k < j; ++k) d = fma(-H[tri(j, k)], H[tri(j, k)], d);
    double inv = rsqrt(d);
    H[tri(j, j)] = inv;
    HDG_UNROLL
    for (int i = j + 1; i < N; ++i) {
      double s = H[tri(i, j)];
      HDG_UNROLL
      for (int k = 0; k < j; ++k) s = fma(-H[tri(i, k)], H[tri(j, k)], s);
      H[tri(i, j)] = s * inv;
    }
  }
}

// x <- (L L^T)^-1 x
template <int N>
__device__ __forceinline__ void chol_solve(const double (&L)[N * (N + 1) / 2], double (&x)[N]) {
  HDG_UNROLL
  for (int i = 0; i < N; ++i) {
    double s = x[i];
    HDG_UNROLL
    for (int k = 0; k < i; ++k) s = fma(-L[tri(i, k)], x[k], s);
    x[i] = s * L[tri(i, i)];
  }
  HDG_UNROLL
  for (int i = N - 1; i >= 0; --i) {
    double s = x[i];
    HDG_UNROLL
    for (int k = i + 1; k < N; ++k) s = fma(-L[tri(k, i)], x[k], s);
    x[i] = s * L[tri(i, i)];
  }
}

// out[a] += scale * (B t)[a] / detJ = scale * sum_d sum_i D[d][a][i] (sum_c Ji[d][c] t[c][i])
template <int K>
__device__ __forceinline__ void apply_B_over_detJ(const Geo& g, const double (&t)[2][Dims<K>::NQ1], double scale,
                                                  double (&out)[Dims<K>::NP]) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NP = Dims<K>::NP;
  HDG_UNROLL
  for (int d = 0; d < 2; ++d) {
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i) {
      double v = scale * (g.Ji[d][0] * t[0][i] + g.Ji[d][1] * t[1][i]);
      HDG_UNROLL
      for (int a = 0; a < NP; ++a)
        if (T::D(d, a, i) != 0.0) out[a] = fma(T::D(d, a, i), v, out[a]);
    }
  }
}

// out[c][i] += scale * (B^T phi)[c][i] / detJ
template <int K>
__device__ __forceinline__ void apply_Bt_over_detJ(const Geo& g, const double (&phi)[Dims<K>::NP], double scale,
                                                   double (&out)[2][Dims<K>::NQ1]) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NP = Dims<K>::NP;
  HDG_UNROLL
  for (int i = 0; i < NQ1; ++i) {
    double s0 = 0.0, s1 = 0.0;
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) {
      if (T::D(0, a, i) != 0.0) s0 = fma(T::D(0, a, i), phi[a], s0);
      if (T::D(1, a, i) != 0.0) s1 = fma(T::D(1, a, i), phi[a], s1);
    }
    s0 *= scale;
    s1 *= scale;
    out[0][i] += g.Ji[0][0] * s0 + g.Ji[1][0] * s1;
    out[1][i] += g.Ji[0][1] * s0 + g.Ji[1][1] * s1;
  }
}

// out[c][i] += scale * (E^T lam)[c][i];  lam holds sigma-corrected trace coefficients lam[e][m]
template <int K>
__device__ __forceinline__ void apply_Et(const Geo& g, const double (&lam)[3][Dims<K>::NL1], double scale,
                                         double (&out)[2][Dims<K>::NQ1]) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NL1 = Dims<K>::NL1;
  HDG_UNROLL
  for (int e = 0; e < 3; ++e) {
    double cx = scale * g.le[e] * g.n[e][0], cy = scale * g.le[e] * g.n[e][1];
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i) {
      double v = 0.0;
      HDG_UNROLL
      for (int m = 0; m < NL1; ++m)
        if (T::E(e, m, i) != 0.0) v = fma(T::E(e, m, i), lam[e][m], v);
      out[0][i] = fma(cx, v, out[0][i]);
      out[1][i] = fma(cy, v, out[1][i]);
    }
  }
}

// out[e][m] += scale * (E u)[e][m]   (sigma NOT applied; caller multiplies by flip signs)
template <int K>
__device__ __forceinline__ void apply_E(const Geo& g, const double (&u)[2][Dims<K>::NQ1], double scale,
                                        double (&out)[3][Dims<K>::NL1]) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NL1 = Dims<K>::NL1;
  HDG_UNROLL
  for (int e = 0; e < 3; ++e) {
    double cx = scale * g.le[e] * g.n[e][0], cy = scale * g.le[e] * g.n[e][1];
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i) {
      double un = cx * u[0][i] + cy * u[1][i];
      HDG_UNROLL
      for (int m = 0; m < NL1; ++m)
        if (T::E(e, m, i) != 0.0) out[e][m] = fma(T::E(e, m, i), un, out[e][m]);
    }
  }
}

// out[e][m] += scale * (F phi)[e][m]   (sigma NOT applied)
template <int K>
__device__ __forceinline__ void apply_F(const Geo& g, const double (&phi)[Dims<K>::NP], double scale,
                                        double (&out)[3][Dims<K>::NL1]) {
  using T = RefTables<K>;
  constexpr int NP = Dims<K>::NP, NL1 = Dims<K>::NL1;
  HDG_UNROLL
  for (int e = 0; e < 3; ++e) {
    double c = scale * g.le[e];
    HDG_UNROLL
    for (int m = 0; m < NL1; ++m) {
      double v = 0.0;
      HDG_UNROLL
      for (int a = 0; a < NP; ++a)
        if (T::F(e, m, a) != 0.0) v = fma(T::F(e, m, a), phi[a], v);
      out[e][m] = fma(c, v, out[e][m]);
    }
  }
}

// out[a] += scale * (F^T lam)[a]   (lam sigma-corrected)
template <int K>
__device__ __forceinline__ void apply_Ft(const Geo& g, const double (&lam)[3][Dims<K>::NL1], double scale,
                                         double (&out)[Dims<K>::NP]) {
  using T = RefTables<K>;
  constexpr int NP = Dims<K>::NP, NL1 = Dims<K>::NL1;
  HDG_UNROLL
  for (int e = 0; e < 3; ++e) {
    double c = scale * g.le[e];
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) {
      double v = 0.0;
      HDG_UNROLL
      for (int m = 0; m < NL1; ++m)
        if (T::F(e, m, a) != 0.0) v = fma(T::F(e, m, a), lam[e][m], v);
      out[a] = fma(c, v, out[a]);
    }
  }
}

// Solve the eliminated block: given local residuals (Ru, Rp) and (optionally) sigma-corrected
// trace values lam, compute  (u, phi) = A_K^-1 ((Ru, Rp) - B_K lam):
//     t    = Ru - E^T lam
//     phi  = H^-1 (Rp - B t / detJ + tau F^T lam)
//     u    = (t + B^T phi) / detJ
// On entry u holds Ru and phi holds Rp; on exit they hold the solution.  L is the Cholesky of H.
template <int K, bool HAS_LAM>
__device__ __forceinline__ void local_solve(const Geo& g, double tau, const double (&L)[Dims<K>::NH],
                                            const double (&lam)[3][Dims<K>::NL1], double (&u)[2][Dims<K>::NQ1],
                                            double (&phi)[Dims<K>::NP]) {
  constexpr int NQ1 = Dims<K>::NQ1;
  if (HAS_LAM) {
    apply_Et<K>(g, lam, -1.0, u);
    apply_Ft<K>(g, lam, tau, phi);
  }
  apply_B_over_detJ<K>(g, u, -1.0, phi);
  chol_solve<Dims<K>::NP>(L, phi);
  HDG_UNROLL
  for (int c = 0; c < 2; ++c) {
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i) u[c][i] *= g.idetJ;
  }
  apply_Bt_over_detJ<K>(g, phi, 1.0, u);
}
