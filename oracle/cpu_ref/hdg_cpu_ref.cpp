// Compiled CPU restatement of the reference's Chorin-projection HDG step  --  TEST / MEASUREMENT INFRASTRUCTURE ONLY.
//
// This is the `cpu_baseline` / `--impl reference` arm of bench.py: the algorithm of
// `IncompressibleEulerHDGImplicit.solve` (src/timesteppers/hdg_implicit.py:92-190) in plain C++ with OpenMP, written
// independently of the product (nothing under incompressibleeulerhdg_b200/csrc is included or linked).  It follows the
// numpy oracle (oracle/hdg_oracle.py, oracle/timesteppers.py), which is pinned against the reference's own form-building
// code (tests/test_forms_golden.py); tests/test_cpu_ref.py checks every stage of this file against that oracle to 1e-10.
// Only tests/, bench.py's cpu_baseline / reference legs and __graft_entry__.build() may build or load it.
//
// What is restated (reference file:line):
//   project_bdm                       src/timesteppers/common.py:59-70,91-108
//   a_tentative / b_rhs_tentative     src/timesteppers/hdg_implicit.py:103-129   (f_impl: hdg_imex.py:313-331)
//   a_poisson / b_rhs_poisson         src/timesteppers/hdg_implicit.py:133-146   (local operator hdg_imex.py:123-127,
//                                     static condensation = firedrake.SCPC hdg_imex.py:128-137: dense LU per cell,
//                                     S_K = D - C A^-1 B, forward elimination, trace solve, back-substitution)
//   velocity / pressure update        src/timesteppers/hdg_implicit.py:150,188-190
// Timer labels are the reference's (src/auxilliary/logging.py:11-31 with hdg_imex.py:257,274,551,564): "timestep",
// "bdm_projection", "tentative_velocity_solve", "pressure_solve".
//
// Solvers.  The reference hands both systems to PETSc (default LU for `solve(a == L)`); a sparse direct solver is not
// restated here.  Tentative velocity: right-preconditioned BiCGStab, operator applied matrix-free by quadrature.  The
// normal-jump penalty of f_impl is alpha N^T N with N = the Legendre moments of the normal jump (h_F = |F|), whose weight
// against the mass matrix grows like dt/h^2, so the preconditioner is the advection-free operator M + dt alpha N^T N,
// inverted with the Woodbury identity: a facet system X = 1/(dt alpha) + N M^-1 N^T (sparse, well conditioned) solved by
// a fixed number of Chebyshev / facet-block-Jacobi sweeps.  (A cell-block-Jacobi preconditioner needs 350-500 iterations
// at nx = 64-128; selectable with precond = 0 for comparison.)
// Pressure correction: conjugate gradients on -S with the facet-diagonal blocks as preconditioner, the constant mode
// projected out of the right-hand side (hdg_imex.py:471-489).  Both run to the relative tolerance the caller passes.
//
// Conventions: FP64, AoS fields Q[nc][2][nQ1], p[nc][np], lam[nf][k+1] in the modal bases the tables were tabulated in
// (the caller passes HDGOracle's tabulation, so the fields are directly comparable with the oracle's).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using vec = std::vector<double>;
using ivec = std::vector<int>;

inline double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// dense LU with partial pivoting, row-major n x n, in place; returns false on a zero pivot
bool lu_factor(double* a, int* piv, int n) {
  for (int c = 0; c < n; ++c) {
    int p = c;
    double best = std::fabs(a[c * n + c]);
    for (int r = c + 1; r < n; ++r)
      if (std::fabs(a[r * n + c]) > best) best = std::fabs(a[r * n + c]), p = r;
    piv[c] = p;
    if (best == 0.0) return false;
    if (p != c)
      for (int j = 0; j < n; ++j) std::swap(a[c * n + j], a[p * n + j]);
    const double inv = 1.0 / a[c * n + c];
    for (int r = c + 1; r < n; ++r) {
      const double l = a[r * n + c] * inv;
      a[r * n + c] = l;
      if (l != 0.0)
        for (int j = c + 1; j < n; ++j) a[r * n + j] -= l * a[c * n + j];
    }
  }
  return true;
}

// solve with nrhs right-hand sides stored as columns of b (row-major n x nrhs), in place
void lu_solve(const double* a, const int* piv, int n, double* b, int nrhs) {
  for (int c = 0; c < n; ++c)  // whole rows were swapped during the factorisation: permute first
    if (piv[c] != c)
      for (int j = 0; j < nrhs; ++j) std::swap(b[c * nrhs + j], b[piv[c] * nrhs + j]);
  for (int c = 0; c < n; ++c) {
    for (int r = c + 1; r < n; ++r) {
      const double l = a[r * n + c];
      if (l != 0.0)
        for (int j = 0; j < nrhs; ++j) b[r * nrhs + j] -= l * b[c * nrhs + j];
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    for (int c = r + 1; c < n; ++c) {
      const double u = a[r * n + c];
      if (u != 0.0)
        for (int j = 0; j < nrhs; ++j) b[r * nrhs + j] -= u * b[c * nrhs + j];
    }
    const double inv = 1.0 / a[r * n + r];
    for (int j = 0; j < nrhs; ++j) b[r * nrhs + j] *= inv;
  }
}

struct Timers {
  double t[4] = {0, 0, 0, 0};  // timestep, bdm_projection, tentative_velocity_solve, pressure_solve
  long n[4] = {0, 0, 0, 0};
};

struct Ref {
  int k, nQ1, nQ, np, nl1, nl, nA, nq, nqf, nint, nfm, nc, nf, upwind;
  double tau, alpha, volume;
  // reference-element tabulation (HDGOracle._tabulate / _bdm_setup)
  vec wq, phiQ, dphiQ, phiP, wf, phiQf, phiPf, ell, bdmF, bdmI;
  // geometry-free tensors built from the tabulation
  vec M1, Bref, Eref, Fref, Tref, Gref, phiQt, d0t, d1t, phiQft, phiQfr;
  // mesh
  vec xy;
  ivec cell_facet, cell_flip, facet_cell, facet_local;
  // geometry
  vec detJ, Jinv, normal, elen, hFinv;
  ivec nbr, nbre, plus;
  // trace system: block-ELL, 5 block columns per facet row
  vec ell_val, dinv, SK;
  ivec ell_col;
  // coarse space of the trace CG: facet mode 0 summed over square aggregates of facets, dense Cholesky factor
  int nagg = 0;
  ivec agg;
  vec coarseL;
  // tentative-velocity preconditioner: cell blocks (precond 0, per step) or facet-multiplier / Woodbury (precond 1)
  vec blk;
  ivec blkpiv;
  int precond = 1, NM = 0, cheb_sweeps = 8;
  vec legN, BF, xval, xdinv;
  double x_adt = -1.0, x_lmax = 0.0;
  Timers tm;
  char err[256] = {0};
};

// ---------------------------------------------------------------------------------------------- geometry
void geometry(Ref& R) {
  const int nc = R.nc, nf = R.nf;
  R.detJ.resize(nc);
  R.Jinv.resize((size_t)nc * 4);
  R.normal.resize((size_t)nc * 6);
  R.elen.resize((size_t)nc * 3);
  R.hFinv.resize(nf);
  R.nbr.assign((size_t)nc * 3, -1);
  R.nbre.assign((size_t)nc * 3, -1);
  R.plus.assign((size_t)nc * 3, 1);
  double vol = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : vol)
  for (int c = 0; c < nc; ++c) {
    const double* x = &R.xy[(size_t)c * 6];
    const double j00 = x[2] - x[0], j01 = x[4] - x[0], j10 = x[3] - x[1], j11 = x[5] - x[1];
    const double det = j00 * j11 - j01 * j10;
    R.detJ[c] = det;
    vol += 0.5 * det;
    double* ji = &R.Jinv[(size_t)c * 4];  // ji[d*2+c] = d xi_d / d x_c
    ji[0] = j11 / det;
    ji[1] = -j01 / det;
    ji[2] = -j10 / det;
    ji[3] = j00 / det;
    for (int e = 0; e < 3; ++e) {
      const double* va = x + 2 * ((e + 1) % 3);
      const double* vb = x + 2 * ((e + 2) % 3);
      const double tx = vb[0] - va[0], ty = vb[1] - va[1];
      const double le = std::hypot(tx, ty);
      R.elen[(size_t)c * 3 + e] = le;
      R.normal[(size_t)c * 6 + 2 * e] = ty / le;
      R.normal[(size_t)c * 6 + 2 * e + 1] = -tx / le;
    }
  }
  R.volume = vol;
#pragma omp parallel for schedule(static)
  for (int f = 0; f < nf; ++f) {
    const int c0 = R.facet_cell[2 * f], c1 = R.facet_cell[2 * f + 1];
    const int e0 = R.facet_local[2 * f], e1 = R.facet_local[2 * f + 1];
    R.hFinv[f] = 1.0 / R.elen[(size_t)c0 * 3 + e0];  // common.py:36-57, measured in the first adjacent cell
    if (c1 >= 0) {
      R.nbr[(size_t)c0 * 3 + e0] = c1;
      R.nbre[(size_t)c0 * 3 + e0] = e1;
      R.nbr[(size_t)c1 * 3 + e1] = c0;
      R.nbre[(size_t)c1 * 3 + e1] = e0;
      R.plus[(size_t)c1 * 3 + e1] = 0;  // the '+' side of a facet is its first cell
    }
  }
}

// geometry-free tensors of the local mixed-Poisson operator (hdg_imex.py:123-127)
void reference_tensors(Ref& R) {
  const int nQ1 = R.nQ1, np = R.np, nl1 = R.nl1, nq = R.nq, nqf = R.nqf;
  R.M1.assign((size_t)nQ1 * nQ1, 0.0);
  for (int i = 0; i < nQ1; ++i)
    for (int j = 0; j < nQ1; ++j) {
      double s = 0;
      for (int q = 0; q < nq; ++q) s += R.wq[q] * R.phiQ[i * nq + q] * R.phiQ[j * nq + q];
      R.M1[i * nQ1 + j] = s;
    }
  R.Bref.assign((size_t)2 * np * nQ1, 0.0);  // [d][a][i] = sum_q w phiP_a d_d phiQ_i
  for (int d = 0; d < 2; ++d)
    for (int a = 0; a < np; ++a)
      for (int i = 0; i < nQ1; ++i) {
        double s = 0;
        for (int q = 0; q < nq; ++q) s += R.wq[q] * R.phiP[a * nq + q] * R.dphiQ[((size_t)i * nq + q) * 2 + d];
        R.Bref[((size_t)d * np + a) * nQ1 + i] = s;
      }
  R.Eref.assign((size_t)3 * 2 * nl1 * nQ1, 0.0);  // [e][flip][m][i]
  R.Fref.assign((size_t)3 * 2 * nl1 * np, 0.0);   // [e][flip][m][a]
  R.Tref.assign((size_t)3 * np * np, 0.0);        // [e][a][b]
  R.Gref.assign((size_t)2 * nl1 * nl1, 0.0);      // [flip][m][l]
  for (int e = 0; e < 3; ++e) {
    for (int fl = 0; fl < 2; ++fl)
      for (int m = 0; m < nl1; ++m) {
        const double* lm = &R.ell[((size_t)fl * nl1 + m) * nqf];
        for (int i = 0; i < nQ1; ++i) {
          double s = 0;
          for (int q = 0; q < nqf; ++q) s += R.wf[q] * lm[q] * R.phiQf[((size_t)e * nQ1 + i) * nqf + q];
          R.Eref[(((size_t)e * 2 + fl) * nl1 + m) * nQ1 + i] = s;
        }
        for (int a = 0; a < np; ++a) {
          double s = 0;
          for (int q = 0; q < nqf; ++q) s += R.wf[q] * lm[q] * R.phiPf[((size_t)e * np + a) * nqf + q];
          R.Fref[(((size_t)e * 2 + fl) * nl1 + m) * np + a] = s;
        }
      }
    for (int a = 0; a < np; ++a)
      for (int b = 0; b < np; ++b) {
        double s = 0;
        for (int q = 0; q < nqf; ++q)
          s += R.wf[q] * R.phiPf[((size_t)e * np + a) * nqf + q] * R.phiPf[((size_t)e * np + b) * nqf + q];
        R.Tref[((size_t)e * np + a) * np + b] = s;
      }
  }
  for (int fl = 0; fl < 2; ++fl)
    for (int m = 0; m < nl1; ++m)
      for (int l = 0; l < nl1; ++l) {
        double s = 0;
        for (int q = 0; q < nqf; ++q)
          s += R.wf[q] * R.ell[((size_t)fl * nl1 + m) * nqf + q] * R.ell[((size_t)fl * nl1 + l) * nqf + q];
        R.Gref[((size_t)fl * nl1 + m) * nl1 + l] = s;
      }
}

// local system of one cell (HDGOracle.local_system): A [nA x nA], Bk [nA x nl], Ck [nl x nA], Dk [nl x nl]
void local_system(const Ref& R, int c, double* A, double* Bk, double* Ck, double* Dk) {
  const int nQ1 = R.nQ1, nQ = R.nQ, np = R.np, nl1 = R.nl1, nl = R.nl, nA = R.nA;
  const double det = R.detJ[c];
  const double* ji = &R.Jinv[(size_t)c * 4];
  std::fill(A, A + (size_t)nA * nA, 0.0);
  std::fill(Bk, Bk + (size_t)nA * nl, 0.0);
  std::fill(Ck, Ck + (size_t)nl * nA, 0.0);
  std::fill(Dk, Dk + (size_t)nl * nl, 0.0);
  for (int cc = 0; cc < 2; ++cc)
    for (int i = 0; i < nQ1; ++i)
      for (int j = 0; j < nQ1; ++j) A[(size_t)(cc * nQ1 + i) * nA + cc * nQ1 + j] = det * R.M1[i * nQ1 + j];
  // B[a][(cc,i)] = detJ sum_d Jinv[d][cc] Bref[d][a][i];  A = [[M, -B^T], [B, T]]
  for (int a = 0; a < np; ++a)
    for (int cc = 0; cc < 2; ++cc)
      for (int i = 0; i < nQ1; ++i) {
        const double b = det * (ji[0 * 2 + cc] * R.Bref[((size_t)0 * np + a) * nQ1 + i] +
                                ji[1 * 2 + cc] * R.Bref[((size_t)1 * np + a) * nQ1 + i]);
        A[(size_t)(nQ + a) * nA + cc * nQ1 + i] = b;
        A[(size_t)(cc * nQ1 + i) * nA + nQ + a] = -b;
      }
  for (int e = 0; e < 3; ++e) {
    const double le = R.elen[(size_t)c * 3 + e];
    const int fl = R.cell_flip[(size_t)c * 3 + e];
    const double* n = &R.normal[(size_t)c * 6 + 2 * e];
    for (int a = 0; a < np; ++a)
      for (int b = 0; b < np; ++b) A[(size_t)(nQ + a) * nA + nQ + b] += R.tau * le * R.Tref[((size_t)e * np + a) * np + b];
    for (int m = 0; m < nl1; ++m) {
      const int row = e * nl1 + m;
      for (int cc = 0; cc < 2; ++cc)
        for (int i = 0; i < nQ1; ++i) {
          const double v = le * n[cc] * R.Eref[(((size_t)e * 2 + fl) * nl1 + m) * nQ1 + i];
          Ck[(size_t)row * nA + cc * nQ1 + i] = v;
          Bk[(size_t)(cc * nQ1 + i) * nl + row] = v;
        }
      for (int a = 0; a < np; ++a) {
        const double v = R.tau * le * R.Fref[(((size_t)e * 2 + fl) * nl1 + m) * np + a];
        Ck[(size_t)row * nA + nQ + a] = v;
        Bk[(size_t)(nQ + a) * nl + row] = -v;
      }
      for (int l = 0; l < nl1; ++l)
        Dk[(size_t)row * nl + e * nl1 + l] = -R.tau * le * R.Gref[((size_t)fl * nl1 + m) * nl1 + l];
    }
  }
}

void coarse_setup(Ref& R);

// S_K = D - C A^-1 B for every cell, gathered facet-wise into the block-ELL trace matrix (hdg_imex.py:128-135)
bool condense(Ref& R) {
  const int nl = R.nl, nA = R.nA, nl1 = R.nl1, nc = R.nc, nf = R.nf, bb = nl1 * nl1;
  R.SK.resize((size_t)nc * nl * nl);
  bool ok = true;
#pragma omp parallel
  {
    vec A((size_t)nA * nA), Bk((size_t)nA * nl), Ck((size_t)nl * nA), Dk((size_t)nl * nl);
    ivec piv(nA);
#pragma omp for schedule(static)
    for (int c = 0; c < nc; ++c) {
      local_system(R, c, A.data(), Bk.data(), Ck.data(), Dk.data());
      if (!lu_factor(A.data(), piv.data(), nA)) ok = false;
      lu_solve(A.data(), piv.data(), nA, Bk.data(), nl);  // Bk <- A^-1 B
      double* S = &R.SK[(size_t)c * nl * nl];
      for (int r = 0; r < nl; ++r)
        for (int s = 0; s < nl; ++s) {
          double v = Dk[(size_t)r * nl + s];
          for (int a = 0; a < nA; ++a) v -= Ck[(size_t)r * nA + a] * Bk[(size_t)a * nl + s];
          S[(size_t)r * nl + s] = v;
        }
    }
  }
  if (!ok) return false;
  R.ell_col.assign((size_t)nf * 5, 0);
  R.ell_val.assign((size_t)nf * 5 * bb, 0.0);
  R.dinv.assign((size_t)nf * bb, 0.0);
#pragma omp parallel for schedule(static)
  for (int f = 0; f < nf; ++f) {
    int* col = &R.ell_col[(size_t)f * 5];
    double* val = &R.ell_val[(size_t)f * 5 * bb];
    for (int j = 0; j < 5; ++j) col[j] = f;
    for (int s = 0; s < 2; ++s) {
      const int c = R.facet_cell[2 * f + s];
      if (c < 0) continue;
      const int e = R.facet_local[2 * f + s];
      const double* S = &R.SK[(size_t)c * nl * nl];
      for (int j = 0; j < 3; ++j) {
        const int e2 = (e + j) % 3;
        const int slot = j == 0 ? 0 : 1 + 2 * s + (j - 1);
        if (j) col[slot] = R.cell_facet[(size_t)c * 3 + e2];
        for (int m = 0; m < nl1; ++m)
          for (int l = 0; l < nl1; ++l) val[(size_t)slot * bb + m * nl1 + l] += S[(size_t)(e * nl1 + m) * nl + e2 * nl1 + l];
      }
    }
    // block-Jacobi preconditioner of P = -S: inverse of the (negated) diagonal block
    double D[64], I[64];
    int piv[8];
    for (int i = 0; i < bb; ++i) D[i] = -val[i];
    for (int m = 0; m < nl1; ++m)
      for (int l = 0; l < nl1; ++l) I[m * nl1 + l] = m == l ? 1.0 : 0.0;
    lu_factor(D, piv, nl1);
    lu_solve(D, piv, nl1, I, nl1);
    for (int i = 0; i < bb; ++i) R.dinv[(size_t)f * bb + i] = I[i];
  }
  coarse_setup(R);
  return true;
}

// Two-level additive preconditioner of the trace CG: block-Jacobi + Z (Z^T P Z)^+ Z^T with Z = the facet mode 0 summed
// over nb x nb square aggregates of facets (about 8 x 8 mesh squares each).  P = -S has the null vector Z 1, so the
// coarse matrix is regularised with the rank-one term (tr/nagg^2) 1 1^T, which acts on that null vector only.
void coarse_setup(Ref& R) {
  const int nf = R.nf, nl1 = R.nl1, bb = nl1 * nl1;
  double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
  for (size_t i = 0; i < R.xy.size(); i += 2)
    for (int d = 0; d < 2; ++d) lo[d] = std::min(lo[d], R.xy[i + d]), hi[d] = std::max(hi[d], R.xy[i + d]);
  const int nb = std::max(1, std::min(48, (int)std::lround(std::sqrt(0.5 * R.nc) / 8.0)));
  R.nagg = nb > 1 ? nb * nb : 0;  // a single aggregate is the null vector itself: no coarse space
  R.agg.assign(nf, 0);
  if (!R.nagg) return;
  for (int f = 0; f < nf; ++f) {
    const int c = R.facet_cell[2 * f], e = R.facet_local[2 * f];
    const double* x = &R.xy[(size_t)c * 6];
    const double mx = 0.5 * (x[2 * ((e + 1) % 3)] + x[2 * ((e + 2) % 3)]), my = 0.5 * (x[2 * ((e + 1) % 3) + 1] + x[2 * ((e + 2) % 3) + 1]);
    const int ix = std::min(nb - 1, std::max(0, (int)((mx - lo[0]) / (hi[0] - lo[0]) * nb)));
    const int iy = std::min(nb - 1, std::max(0, (int)((my - lo[1]) / (hi[1] - lo[1]) * nb)));
    R.agg[f] = iy * nb + ix;
  }
  const int n = R.nagg;
  vec A((size_t)n * n, 0.0);
  for (int f = 0; f < nf; ++f)
    for (int j = 0; j < 5; ++j) {
      const int col = R.ell_col[(size_t)f * 5 + j];
      if (j && col == f) continue;
      A[(size_t)R.agg[f] * n + R.agg[col]] -= R.ell_val[((size_t)f * 5 + j) * bb];  // (mode 0, mode 0) entry of P = -S
    }
  double tr = 0;
  for (int i = 0; i < n; ++i) tr += A[(size_t)i * n + i];
  const double reg = tr / ((double)n * n);
  for (size_t i = 0; i < A.size(); ++i) A[i] += reg;
  // dense Cholesky A = L L^T (lower, row-major)
  for (int j = 0; j < n; ++j) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
    d = std::sqrt(std::max(d, 1e-300));
    A[(size_t)j * n + j] = d;
#pragma omp parallel for schedule(static)
    for (int i = j + 1; i < n; ++i) {
      double v = A[(size_t)i * n + j];
      for (int k = 0; k < j; ++k) v -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      A[(size_t)i * n + j] = v / d;
    }
  }
  R.coarseL.swap(A);
}

// y = -S x  (P = -S is symmetric positive semi-definite)
void spmv_neg(const Ref& R, const double* x, double* y) {
  const int nl1 = R.nl1, nf = R.nf, bb = nl1 * nl1;
#pragma omp parallel for schedule(static)
  for (int f = 0; f < nf; ++f) {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const double* val = &R.ell_val[(size_t)f * 5 * bb];
    const int* col = &R.ell_col[(size_t)f * 5];
    for (int j = 0; j < 5; ++j) {
      if (j && col[j] == f) continue;  // empty slot
      const double* xv = &x[(size_t)col[j] * nl1];
      for (int m = 0; m < nl1; ++m)
        for (int l = 0; l < nl1; ++l) acc[m] -= val[(size_t)j * bb + m * nl1 + l] * xv[l];
    }
    for (int m = 0; m < nl1; ++m) y[(size_t)f * nl1 + m] = acc[m];
  }
}

double dot(const double* a, const double* b, size_t n) {
  double s = 0;
#pragma omp parallel for schedule(static) reduction(+ : s)
  for (size_t i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

// condensed mixed-Poisson solve (HDGOracle.solve_condensed): Ru [nc][nQ], Rp [nc][np], Rl [nf][nl1] (any may be null = 0)
int poisson_solve(Ref& R, const double* Ru, const double* Rp, const double* Rl, double rtol, int maxit, double* u,
                  double* p, double* lam, int* iters) {
  const int nl = R.nl, nA = R.nA, nl1 = R.nl1, nc = R.nc, nf = R.nf, nQ = R.nQ, np = R.np, bb = nl1 * nl1;
  const size_t n = (size_t)nf * nl1;
  // forward elimination: r = Rl - sum_K P_K^T C_K A_K^-1 R_K   (local LU recomputed like Slate does)
  vec contrib((size_t)nc * nl), r(n);
#pragma omp parallel
  {
    vec A((size_t)nA * nA), Bk((size_t)nA * nl), Ck((size_t)nl * nA), Dk((size_t)nl * nl), x0(nA);
    ivec piv(nA);
#pragma omp for schedule(static)
    for (int c = 0; c < nc; ++c) {
      local_system(R, c, A.data(), Bk.data(), Ck.data(), Dk.data());
      lu_factor(A.data(), piv.data(), nA);
      for (int i = 0; i < nQ; ++i) x0[i] = Ru ? Ru[(size_t)c * nQ + i] : 0.0;
      for (int a = 0; a < np; ++a) x0[nQ + a] = Rp ? Rp[(size_t)c * np + a] : 0.0;
      lu_solve(A.data(), piv.data(), nA, x0.data(), 1);
      for (int rI = 0; rI < nl; ++rI) {
        double v = 0;
        for (int a = 0; a < nA; ++a) v += Ck[(size_t)rI * nA + a] * x0[a];
        contrib[(size_t)c * nl + rI] = v;
      }
    }
  }
  double zr = 0;
#pragma omp parallel for schedule(static) reduction(+ : zr)
  for (int f = 0; f < nf; ++f) {
    for (int m = 0; m < nl1; ++m) {
      double v = Rl ? Rl[(size_t)f * nl1 + m] : 0.0;
      for (int s = 0; s < 2; ++s) {
        const int c = R.facet_cell[2 * f + s];
        if (c >= 0) v -= contrib[(size_t)c * nl + R.facet_local[2 * f + s] * nl1 + m];
      }
      r[(size_t)f * nl1 + m] = v;
      if (m == 0) zr += v;
    }
  }
  // remove the component along the null vector z (mode 0 == 1 on every facet): r -= z (z.r)/(z.z)
  const double shift = zr / nf;
#pragma omp parallel for schedule(static)
  for (int f = 0; f < nf; ++f) r[(size_t)f * nl1] -= shift;
  // CG on P lam = -r  (P = -S), block-Jacobi preconditioner
  vec x(n, 0.0), z(n), pd(n), Ap(n), res(n);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) res[i] = -r[i];
  vec cr(R.nagg);
  auto precond = [&](const double* in, double* out) {
    // coarse correction: cr = (Z^T P Z)^-1 Z^T in
    const int na = R.nagg;
    cr.assign(std::max(na, 1), 0.0);
    for (int f = 0; f < nf; ++f) cr[R.agg[f]] += in[(size_t)f * nl1];
    const double* L = R.coarseL.data();
    for (int i = 0; i < na; ++i) {
      double v = cr[i];
      for (int k2 = 0; k2 < i; ++k2) v -= L[(size_t)i * na + k2] * cr[k2];
      cr[i] = v / L[(size_t)i * na + i];
    }
    for (int i = na - 1; i >= 0; --i) {
      double v = cr[i] / L[(size_t)i * na + i];
      cr[i] = v;
      for (int k2 = 0; k2 < i; ++k2) cr[k2] -= L[(size_t)i * na + k2] * v;
    }
#pragma omp parallel for schedule(static)
    for (int f = 0; f < nf; ++f)
      for (int m = 0; m < nl1; ++m) {
        double v = m == 0 ? cr[R.agg[f]] : 0.0;
        for (int l = 0; l < nl1; ++l) v += R.dinv[(size_t)f * bb + m * nl1 + l] * in[(size_t)f * nl1 + l];
        out[(size_t)f * nl1 + m] = v;
      }
  };
  precond(res.data(), z.data());
  pd = z;
  double rz = dot(res.data(), z.data(), n);
  const double rz0 = rz;
  int it = 0;
  while (it < maxit && rz > rtol * rtol * rz0 && rz0 > 0) {
    spmv_neg(R, pd.data(), Ap.data());
    const double alpha = rz / dot(pd.data(), Ap.data(), n);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
      x[i] += alpha * pd[i];
      res[i] -= alpha * Ap[i];
    }
    precond(res.data(), z.data());
    const double rz1 = dot(res.data(), z.data(), n);
    const double beta = rz1 / rz;
    rz = rz1;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) pd[i] = z[i] + beta * pd[i];
    ++it;
  }
  if (iters) *iters = it;
  const bool conv = !(rz > rtol * rtol * rz0);
  // back-substitution x_K = A_K^-1 (R_K - B_K lam_K)
  double pint = 0;
#pragma omp parallel
  {
    vec A((size_t)nA * nA), Bk((size_t)nA * nl), Ck((size_t)nl * nA), Dk((size_t)nl * nl), xl(nA);
    ivec piv(nA);
#pragma omp for schedule(static) reduction(+ : pint)
    for (int c = 0; c < nc; ++c) {
      local_system(R, c, A.data(), Bk.data(), Ck.data(), Dk.data());
      lu_factor(A.data(), piv.data(), nA);
      for (int i = 0; i < nQ; ++i) xl[i] = Ru ? Ru[(size_t)c * nQ + i] : 0.0;
      for (int a = 0; a < np; ++a) xl[nQ + a] = Rp ? Rp[(size_t)c * np + a] : 0.0;
      for (int e = 0; e < 3; ++e) {
        const int f = R.cell_facet[(size_t)c * 3 + e];
        for (int m = 0; m < nl1; ++m) {
          const double lv = x[(size_t)f * nl1 + m];
          for (int a = 0; a < nA; ++a) xl[a] -= Bk[(size_t)a * nl + e * nl1 + m] * lv;
        }
      }
      lu_solve(A.data(), piv.data(), nA, xl.data(), 1);
      for (int i = 0; i < nQ; ++i) u[(size_t)c * nQ + i] = xl[i];
      for (int a = 0; a < np; ++a) p[(size_t)c * np + a] = xl[nQ + a];
      pint += R.detJ[c] * xl[nQ] / std::sqrt(2.0);  // int_K p dx: Dubiner mode 0 is the constant sqrt(2)
    }
  }
  // _shift_pressure (hdg_imex.py:471-478)
  const double sh = pint / R.volume;
#pragma omp parallel for schedule(static)
  for (int c = 0; c < nc; ++c) p[(size_t)c * np] -= sh / std::sqrt(2.0);  // p == 1 has the coefficient 1/sqrt(2)
#pragma omp parallel for schedule(static)
  for (int f = 0; f < nf; ++f) {
    for (int m = 0; m < nl1; ++m) lam[(size_t)f * nl1 + m] = x[(size_t)f * nl1 + m];
    lam[(size_t)f * nl1] -= sh;
  }
  return conv ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------- BDM projection
int project_bdm(const Ref& R, const double* Q, double* Qstar) {
  const int nQ1 = R.nQ1, nQ = R.nQ, k1 = R.k + 1, nm = k1 + 1, nfm = R.nfm, nint = R.nint, nc = R.nc;
  vec mom((size_t)nc * nQ);
  // moments of every cell: L_K Q_K
  auto build_L = [&](int c, double* L) {
    const double* ji = &R.Jinv[(size_t)c * 4];
    for (int e = 0; e < 3; ++e) {
      const double* n = &R.normal[(size_t)c * 6 + 2 * e];
      for (int j = 0; j < nm; ++j)
        for (int cc = 0; cc < 2; ++cc)
          for (int i = 0; i < nQ1; ++i)
            L[(size_t)(e * nm + j) * nQ + cc * nQ1 + i] = n[cc] * R.bdmF[((size_t)e * nm + j) * nQ1 + i];
    }
    for (int w = 0; w < nint; ++w)
      for (int cc = 0; cc < 2; ++cc)
        for (int i = 0; i < nQ1; ++i)
          L[(size_t)(nfm + w) * nQ + cc * nQ1 + i] = ji[0 * 2 + cc] * R.bdmI[((size_t)w * 2 + 0) * nQ1 + i] +
                                                     ji[1 * 2 + cc] * R.bdmI[((size_t)w * 2 + 1) * nQ1 + i];
  };
#pragma omp parallel
  {
    vec L((size_t)nQ * nQ);
#pragma omp for schedule(static)
    for (int c = 0; c < nc; ++c) {
      build_L(c, L.data());
      for (int r = 0; r < nQ; ++r) {
        double v = 0;
        for (int i = 0; i < nQ; ++i) v += L[(size_t)r * nQ + i] * Q[(size_t)c * nQ + i];
        mom[(size_t)c * nQ + r] = v;
      }
    }
  }
  bool ok = true;
#pragma omp parallel
  {
    vec L((size_t)nQ * nQ), m(nQ);
    ivec piv(nQ);
#pragma omp for schedule(static)
    for (int c = 0; c < nc; ++c) {
      for (int r = 0; r < nQ; ++r) m[r] = mom[(size_t)c * nQ + r];
      for (int e = 0; e < 3; ++e) {
        const int nb = R.nbr[(size_t)c * 3 + e];
        for (int j = 0; j < nm; ++j) {
          if (nb < 0)
            m[e * nm + j] = 0.0;  // DirichletBC zero on the boundary (common.py:106-107)
          else {
            const double sign = (j % 2 == 0) ? -1.0 : 1.0;  // reversed parametrisation and opposite normal
            m[e * nm + j] = 0.5 * (mom[(size_t)c * nQ + e * nm + j] +
                                   sign * mom[(size_t)nb * nQ + R.nbre[(size_t)c * 3 + e] * nm + j]);
          }
        }
      }
      build_L(c, L.data());
      if (!lu_factor(L.data(), piv.data(), nQ)) ok = false;
      lu_solve(L.data(), piv.data(), nQ, m.data(), 1);
      for (int i = 0; i < nQ; ++i) Qstar[(size_t)c * nQ + i] = m[i];
    }
  }
  return ok ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------- f_impl
// per-cell scratch for the quadrature evaluation of f_impl
struct FScratch {
  vec gphi, adv, Qf, Qn, s, vecv;
  explicit FScratch(const Ref& R)
      : gphi((size_t)R.nQ1 * R.nq * 2), adv((size_t)R.nq * 2), Qf((size_t)R.nqf * 2), Qn((size_t)R.nqf * 2), s(R.nqf),
        vecv((size_t)R.nqf * 2) {}
};

// physical gradients of the velocity basis, gphi[(i nq + q) 2 + c] (used by the cell-block preconditioner)
inline void physical_gradients(const Ref& R, int c, double* gphi) {
  const double* ji = &R.Jinv[(size_t)c * 4];
  const size_t n = (size_t)R.nQ1 * R.nq;
  for (size_t iq = 0; iq < n; ++iq) {
    const double d0 = R.dphiQ[iq * 2], d1 = R.dphiQ[iq * 2 + 1];
    gphi[iq * 2] = ji[0] * d0 + ji[2] * d1;
    gphi[iq * 2 + 1] = ji[1] * d0 + ji[3] * d1;
  }
}

// point-major copies of the tabulation (contiguous in the basis index, which is the inner loop of every contraction):
// phiQt[q][i], d0t/d1t[q][i] (reference derivatives), phiQft[e][q][i] and phiQfr[e][q][i] = phiQf[e][i][nqf-1-q] (the
// neighbour traverses the facet in the opposite direction)
void transpose_tables(Ref& R) {
  const int nQ1 = R.nQ1, nq = R.nq, nqf = R.nqf;
  R.phiQt.resize((size_t)nq * nQ1), R.d0t.resize((size_t)nq * nQ1), R.d1t.resize((size_t)nq * nQ1);
  for (int q = 0; q < nq; ++q)
    for (int i = 0; i < nQ1; ++i) {
      R.phiQt[(size_t)q * nQ1 + i] = R.phiQ[(size_t)i * nq + q];
      R.d0t[(size_t)q * nQ1 + i] = R.dphiQ[((size_t)i * nq + q) * 2];
      R.d1t[(size_t)q * nQ1 + i] = R.dphiQ[((size_t)i * nq + q) * 2 + 1];
    }
  R.phiQft.resize((size_t)3 * nqf * nQ1), R.phiQfr.resize((size_t)3 * nqf * nQ1);
  for (int e = 0; e < 3; ++e)
    for (int q = 0; q < nqf; ++q)
      for (int i = 0; i < nQ1; ++i) {
        R.phiQft[((size_t)e * nqf + q) * nQ1 + i] = R.phiQf[((size_t)e * nQ1 + i) * nqf + q];
        R.phiQfr[((size_t)e * nqf + q) * nQ1 + i] = R.phiQf[((size_t)e * nQ1 + i) * nqf + (nqf - 1 - q)];
      }
}

inline void eval2(const double* __restrict tab, const double* __restrict q, int nQ1, double& v0, double& v1) {
  double a = 0, b = 0;
  for (int i = 0; i < nQ1; ++i) {
    a += q[i] * tab[i];
    b += q[nQ1 + i] * tab[i];
  }
  v0 = a;
  v1 = b;
}

// Q*('+').n_K at the facet points of facet e of cell c (HDGOracle._plus_side_flux)
inline void plus_side_flux(const Ref& R, const double* Qstar, int c, int e, double* s) {
  const int nQ1 = R.nQ1, nQ = R.nQ, nqf = R.nqf;
  const double* n = &R.normal[(size_t)c * 6 + 2 * e];
  const bool own = R.plus[(size_t)c * 3 + e] != 0;
  const int cell = own ? c : R.nbr[(size_t)c * 3 + e];
  const int ee = own ? e : R.nbre[(size_t)c * 3 + e];
  const double* tab = own ? &R.phiQft[(size_t)ee * nqf * nQ1] : &R.phiQfr[(size_t)ee * nqf * nQ1];
  const double* q = &Qstar[(size_t)cell * nQ];
  for (int qq = 0; qq < nqf; ++qq) {
    double v0, v1;
    eval2(tab + (size_t)qq * nQ1, q, nQ1, v0, v1);
    s[qq] = n[0] * v0 + n[1] * v1;
  }
}

// out = f_impl(w, Q; Q*) as a dual vector (HDGOracle.f_impl_apply, hdg_imex.py:313-331)
void fimpl_cell(const Ref& R, const double* Q, const double* Qstar, int c, FScratch& S, double* __restrict out) {
  const int nQ1 = R.nQ1, nQ = R.nQ, nq = R.nq, nqf = R.nqf;
  const double det = R.detJ[c];
  const double* ji = &R.Jinv[(size_t)c * 4];
  const double* qk = &Q[(size_t)c * nQ];
  const double* qs = &Qstar[(size_t)c * nQ];
  for (int i = 0; i < nQ; ++i) out[i] = 0.0;
  for (int q = 0; q < nq; ++q) {
    const double* ph = &R.phiQt[(size_t)q * nQ1];
    const double* d0 = &R.d0t[(size_t)q * nQ1];
    const double* d1 = &R.d1t[(size_t)q * nQ1];
    double s0 = 0, s1 = 0, a0 = 0, a1 = 0, b0 = 0, b1 = 0;
    for (int i = 0; i < nQ1; ++i) {
      s0 += qs[i] * ph[i];
      s1 += qs[nQ1 + i] * ph[i];
      a0 += qk[i] * d0[i];
      a1 += qk[i] * d1[i];
      b0 += qk[nQ1 + i] * d0[i];
      b1 += qk[nQ1 + i] * d1[i];
    }
    // physical gradients d_x = Jinv[0][0] d_xi + Jinv[1][0] d_eta, d_y = Jinv[0][1] d_xi + Jinv[1][1] d_eta
    const double gxx = ji[0] * a0 + ji[2] * a1, gxy = ji[1] * a0 + ji[3] * a1;
    const double gyx = ji[0] * b0 + ji[2] * b1, gyy = ji[1] * b0 + ji[3] * b1;
    const double w = -det * R.wq[q];
    const double adv0 = w * (s0 * gxx + s1 * gxy), adv1 = w * (s0 * gyx + s1 * gyy);
    for (int i = 0; i < nQ1; ++i) {
      out[i] += adv0 * ph[i];
      out[nQ1 + i] += adv1 * ph[i];
    }
  }
  for (int e = 0; e < 3; ++e) {
    const int nb = R.nbr[(size_t)c * 3 + e];
    const double le = R.elen[(size_t)c * 3 + e];
    const double* n = &R.normal[(size_t)c * 6 + 2 * e];
    const double pen = R.alpha * R.hFinv[R.cell_facet[(size_t)c * 3 + e]];
    const double* tab = &R.phiQft[(size_t)e * nqf * nQ1];
    if (nb >= 0) {
      const double* tabn = &R.phiQfr[(size_t)R.nbre[(size_t)c * 3 + e] * nqf * nQ1];
      const double* qn = &Q[(size_t)nb * nQ];
      plus_side_flux(R, Qstar, c, e, S.s.data());
      for (int qq = 0; qq < nqf; ++qq) {
        double f0, f1, n0, n1;
        eval2(tab + (size_t)qq * nQ1, qk, nQ1, f0, f1);
        eval2(tabn + (size_t)qq * nQ1, qn, nQ1, n0, n1);
        const double j0 = f0 - n0, j1 = f1 - n1;
        double coef = 0.5 * S.s[qq];
        if (R.upwind) coef -= std::fabs(S.s[qq]);
        const double nj = pen * (n[0] * j0 + n[1] * j1);
        const double w = le * R.wf[qq];
        const double v0 = w * (coef * j0 - nj * n[0]), v1 = w * (coef * j1 - nj * n[1]);
        const double* ph = tab + (size_t)qq * nQ1;
        for (int i = 0; i < nQ1; ++i) {
          out[i] += v0 * ph[i];
          out[nQ1 + i] += v1 * ph[i];
        }
      }
    } else {
      for (int qq = 0; qq < nqf; ++qq) {
        double f0, f1;
        eval2(tab + (size_t)qq * nQ1, qk, nQ1, f0, f1);
        const double nq_ = -le * R.wf[qq] * pen * (n[0] * f0 + n[1] * f1);
        const double* ph = tab + (size_t)qq * nQ1;
        for (int i = 0; i < nQ1; ++i) {
          out[i] += nq_ * n[0] * ph[i];
          out[nQ1 + i] += nq_ * n[1] * ph[i];
        }
      }
    }
  }
}

// (w, Q) dx by quadrature (HDGOracle.mass_Q)
inline void mass_cell(const Ref& R, const double* q, int c, double* out) {
  const int nQ1 = R.nQ1;
  const double det = R.detJ[c];
  for (int cc = 0; cc < 2; ++cc)
    for (int i = 0; i < nQ1; ++i) {
      double v = 0;
      for (int j = 0; j < nQ1; ++j) v += R.M1[i * nQ1 + j] * q[cc * nQ1 + j];
      out[cc * nQ1 + i] = det * v;
    }
}

// y = mass(x) - adt f_impl(x; Q*)    (a_tentative, hdg_implicit.py:103-125)
void tentative_apply(const Ref& R, const double* Qstar, double adt, const double* x, double* y) {
  const int nQ = R.nQ, nc = R.nc;
#pragma omp parallel
  {
    FScratch S(R);
    vec f(nQ), m(nQ);
#pragma omp for schedule(static)
    for (int c = 0; c < nc; ++c) {
      fimpl_cell(R, x, Qstar, c, S, f.data());
      mass_cell(R, &x[(size_t)c * nQ], c, m.data());
      for (int i = 0; i < nQ; ++i) y[(size_t)c * nQ + i] = m[i] - adt * f[i];
    }
  }
}

// cell-diagonal blocks of a_tentative, LU-factorised (the BiCGStab preconditioner)
bool tentative_blocks(Ref& R, const double* Qstar, double adt) {
  const int nQ1 = R.nQ1, nQ = R.nQ, nq = R.nq, nqf = R.nqf, nc = R.nc;
  R.blk.resize((size_t)nc * nQ * nQ);
  R.blkpiv.resize((size_t)nc * nQ);
  bool ok = true;
#pragma omp parallel
  {
    vec gphi((size_t)nQ1 * nq * 2), sc((size_t)nQ1 * nQ1), s(nqf), wphi((size_t)nQ1 * nqf);
#pragma omp for schedule(static)
    for (int c = 0; c < nc; ++c) {
      const double det = R.detJ[c];
      const double* qs = &Qstar[(size_t)c * nQ];
      physical_gradients(R, c, gphi.data());
      // scalar part shared by both components: detJ M1 - adt ( -detJ int phi_i (Q*.grad phi_j) + facet flux )
      for (int i = 0; i < nQ1 * nQ1; ++i) sc[i] = det * R.M1[i];
      for (int q = 0; q < nq; ++q) {
        double s0 = 0, s1 = 0;
        for (int i = 0; i < nQ1; ++i) {
          const double ph = R.phiQ[(size_t)i * nq + q];
          s0 += qs[i] * ph;
          s1 += qs[nQ1 + i] * ph;
        }
        const double w = adt * det * R.wq[q];
        for (int i = 0; i < nQ1; ++i) {
          const double wi = w * R.phiQ[(size_t)i * nq + q];
          for (int j = 0; j < nQ1; ++j)
            sc[i * nQ1 + j] += wi * (s0 * gphi[((size_t)j * nq + q) * 2] + s1 * gphi[((size_t)j * nq + q) * 2 + 1]);
        }
      }
      double* D = &R.blk[(size_t)c * nQ * nQ];
      std::fill(D, D + (size_t)nQ * nQ, 0.0);
      for (int e = 0; e < 3; ++e) {
        const int nb = R.nbr[(size_t)c * 3 + e];
        const double le = R.elen[(size_t)c * 3 + e];
        const double* n = &R.normal[(size_t)c * 6 + 2 * e];
        const double pen = R.alpha * R.hFinv[R.cell_facet[(size_t)c * 3 + e]];
        if (nb >= 0) {
          plus_side_flux(R, Qstar, c, e, s.data());
          for (int qq = 0; qq < nqf; ++qq) {
            double coef = 0.5 * s[qq];
            if (R.upwind) coef -= std::fabs(s[qq]);
            const double w = adt * le * R.wf[qq] * coef;
            for (int i = 0; i < nQ1; ++i) {
              const double wi = w * R.phiQf[((size_t)e * nQ1 + i) * nqf + qq];
              for (int j = 0; j < nQ1; ++j) sc[i * nQ1 + j] -= wi * R.phiQf[((size_t)e * nQ1 + j) * nqf + qq];
            }
          }
        }
        // penalty (own part, interior and boundary): + adt alpha/h_F |e| n_c n_c' int phi_i phi_j
        for (int i = 0; i < nQ1; ++i)
          for (int j = 0; j < nQ1; ++j) {
            double v = 0;
            for (int qq = 0; qq < nqf; ++qq)
              v += R.wf[qq] * R.phiQf[((size_t)e * nQ1 + i) * nqf + qq] * R.phiQf[((size_t)e * nQ1 + j) * nqf + qq];
            v *= adt * pen * le;
            for (int c1 = 0; c1 < 2; ++c1)
              for (int c2 = 0; c2 < 2; ++c2) D[(size_t)(c1 * nQ1 + i) * nQ + c2 * nQ1 + j] += n[c1] * n[c2] * v;
          }
      }
      for (int cc = 0; cc < 2; ++cc)
        for (int i = 0; i < nQ1; ++i)
          for (int j = 0; j < nQ1; ++j) D[(size_t)(cc * nQ1 + i) * nQ + cc * nQ1 + j] += sc[i * nQ1 + j];
      if (!lu_factor(D, &R.blkpiv[(size_t)c * nQ], nQ)) ok = false;
    }
  }
  return ok;
}

void block_precond(const Ref& R, const double* in, double* out) {
  const int nQ = R.nQ, nc = R.nc;
#pragma omp parallel for schedule(static)
  for (int c = 0; c < nc; ++c) {
    double* o = &out[(size_t)c * nQ];
    for (int i = 0; i < nQ; ++i) o[i] = in[(size_t)c * nQ + i];
    lu_solve(&R.blk[(size_t)c * nQ * nQ], &R.blkpiv[(size_t)c * nQ], nQ, o, 1);
  }
}

// ---- facet-multiplier preconditioner --------------------------------------------------------------------------------
// BF[e][fl][j][i] = int_0^1 l_j(s_glob) phi_i|_e ds : moments of the trace of basis function i on local facet e
void moment_tables(Ref& R) {
  const int nQ1 = R.nQ1, nqf = R.nqf, NM = R.NM;
  R.BF.assign((size_t)3 * 2 * NM * nQ1, 0.0);
  for (int e = 0; e < 3; ++e)
    for (int fl = 0; fl < 2; ++fl)
      for (int j = 0; j < NM; ++j)
        for (int i = 0; i < nQ1; ++i) {
          double v = 0;
          for (int q = 0; q < nqf; ++q)
            v += R.wf[q] * R.legN[((size_t)fl * NM + j) * nqf + q] * R.phiQf[((size_t)e * nQ1 + i) * nqf + q];
          R.BF[(((size_t)e * 2 + fl) * NM + j) * nQ1 + i] = v;
        }
}

// y = X x,  X = 1/(adt alpha) + N M^-1 N^T  (block-ELL with the column pattern of the trace matrix)
void x_spmv(const Ref& R, const double* x, double* y) {
  const int NM = R.NM, nf = R.nf, bb = NM * NM;
#pragma omp parallel for schedule(static)
  for (int f = 0; f < nf; ++f) {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const double* val = &R.xval[(size_t)f * 5 * bb];
    const int* col = &R.ell_col[(size_t)f * 5];
    for (int j = 0; j < 5; ++j) {
      if (j && col[j] == f) continue;
      const double* xv = &x[(size_t)col[j] * NM];
      for (int m = 0; m < NM; ++m)
        for (int l = 0; l < NM; ++l) acc[m] += val[(size_t)j * bb + m * NM + l] * xv[l];
    }
    for (int m = 0; m < NM; ++m) y[(size_t)f * NM + m] = acc[m];
  }
}

void x_dinv(const Ref& R, const double* in, double* out) {
  const int NM = R.NM, nf = R.nf, bb = NM * NM;
#pragma omp parallel for schedule(static)
  for (int f = 0; f < nf; ++f)
    for (int m = 0; m < NM; ++m) {
      double v = 0;
      for (int l = 0; l < NM; ++l) v += R.xdinv[(size_t)f * bb + m * NM + l] * in[(size_t)f * NM + l];
      out[(size_t)f * NM + m] = v;
    }
}

bool x_setup(Ref& R, double adt) {
  if (R.x_adt == adt) return true;
  const int NM = R.NM, nQ1 = R.nQ1, nc = R.nc, nf = R.nf, bb = NM * NM, ng = 3 * NM;
  if (R.BF.empty()) moment_tables(R);
  vec GK((size_t)nc * ng * ng);
#pragma omp parallel for schedule(static)
  for (int c = 0; c < nc; ++c) {
    const double idet = 1.0 / R.detJ[c];
    for (int e = 0; e < 3; ++e)
      for (int e2 = 0; e2 < 3; ++e2) {
        const double* n = &R.normal[(size_t)c * 6 + 2 * e];
        const double* n2 = &R.normal[(size_t)c * 6 + 2 * e2];
        const double nn = (n[0] * n2[0] + n[1] * n2[1]) * idet;
        const double* b1 = &R.BF[((size_t)e * 2 + R.cell_flip[(size_t)c * 3 + e]) * NM * nQ1];
        const double* b2 = &R.BF[((size_t)e2 * 2 + R.cell_flip[(size_t)c * 3 + e2]) * NM * nQ1];
        for (int j = 0; j < NM; ++j)
          for (int l = 0; l < NM; ++l) {
            double v = 0;
            for (int i = 0; i < nQ1; ++i) v += b1[j * nQ1 + i] * b2[l * nQ1 + i];
            GK[((size_t)c * ng + e * NM + j) * ng + e2 * NM + l] = nn * v;
          }
      }
  }
  R.xval.assign((size_t)nf * 5 * bb, 0.0);
  R.xdinv.assign((size_t)nf * bb, 0.0);
  const double ia = 1.0 / (adt * R.alpha);
  bool ok = true;
#pragma omp parallel for schedule(static)
  for (int f = 0; f < nf; ++f) {
    double* val = &R.xval[(size_t)f * 5 * bb];
    for (int m = 0; m < NM; ++m) val[m * NM + m] = ia;
    for (int s = 0; s < 2; ++s) {
      const int c = R.facet_cell[2 * f + s];
      if (c < 0) continue;
      const int e = R.facet_local[2 * f + s];
      for (int j = 0; j < 3; ++j) {
        const int e2 = (e + j) % 3;
        const int slot = j == 0 ? 0 : 1 + 2 * s + (j - 1);
        for (int m = 0; m < NM; ++m)
          for (int l = 0; l < NM; ++l)
            val[(size_t)slot * bb + m * NM + l] += GK[((size_t)c * ng + e * NM + m) * ng + e2 * NM + l];
      }
    }
    double D[64], I[64];
    int piv[8];
    for (int i = 0; i < bb; ++i) D[i] = val[i];
    for (int m = 0; m < NM; ++m)
      for (int l = 0; l < NM; ++l) I[m * NM + l] = m == l ? 1.0 : 0.0;
    if (!lu_factor(D, piv, NM)) ok = false;
    lu_solve(D, piv, NM, I, NM);
    for (int i = 0; i < bb; ++i) R.xdinv[(size_t)f * bb + i] = I[i];
  }
  if (!ok) return false;
  // lambda_max(D^-1 X) by power iteration (from below; the Chebyshev interval ends 10 % above it)
  const size_t n = (size_t)nf * NM;
  vec x(n), y(n), z(n);
  uint64_t st = 0x9E3779B97F4A7C15ull;
  for (size_t i = 0; i < n; ++i) {
    st ^= st << 13, st ^= st >> 7, st ^= st << 17;
    x[i] = (double)(st >> 11) / 9007199254740992.0 - 0.5;
  }
  double lam = 1.0;
  for (int it = 0; it < 40; ++it) {
    x_spmv(R, x.data(), y.data());
    x_dinv(R, y.data(), z.data());
    lam = std::sqrt(dot(z.data(), z.data(), n) / dot(x.data(), x.data(), n));
    const double inv = 1.0 / std::sqrt(dot(z.data(), z.data(), n));
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) x[i] = z[i] * inv;
  }
  R.x_lmax = lam;
  R.x_adt = adt;
  return true;
}

// mu ~= X^-1 t by a fixed number of Chebyshev / block-Jacobi sweeps (a fixed linear operator)
void x_solve(const Ref& R, const double* t, double* mu, vec& r, vec& d, vec& w) {
  const size_t n = (size_t)R.nf * R.NM;
  const double lmax = 1.1 * R.x_lmax, lmin = lmax / 8.0;
  const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
  double rho = 1.0 / sigma;
  x_dinv(R, t, d.data());
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) {
    r[i] = t[i];
    d[i] /= theta;
    mu[i] = 0.0;
  }
  for (int j = 0; j < R.cheb_sweeps; ++j) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) mu[i] += d[i];
    if (j == R.cheb_sweeps - 1) break;
    x_spmv(R, d.data(), w.data());
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) r[i] -= w[i];
    x_dinv(R, r.data(), w.data());
    const double rho1 = 1.0 / (2.0 * sigma - rho);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) d[i] = rho1 * rho * d[i] + 2.0 * rho1 / delta * w[i];
    rho = rho1;
  }
}

struct WoodburyScratch {
  vec t, mu, r, d, w;
  explicit WoodburyScratch(const Ref& R) {
    const size_t n = (size_t)R.nf * R.NM;
    t.resize(n), mu.resize(n), r.resize(n), d.resize(n), w.resize(n);
  }
};

// out = (M + adt alpha N^T N)^-1 in  =  M^-1 in - M^-1 N^T X^-1 N M^-1 in
void woodbury_precond(const Ref& R, const double* in, double* out, WoodburyScratch& W) {
  const int NM = R.NM, nQ1 = R.nQ1, nQ = R.nQ, nc = R.nc, nf = R.nf;
  // cell moments of M^-1 in, summed facet-wise: t = N M^-1 in
  vec cm((size_t)nc * 3 * NM);
#pragma omp parallel for schedule(static)
  for (int c = 0; c < nc; ++c) {
    const double idet = 1.0 / R.detJ[c];
    for (int e = 0; e < 3; ++e) {
      const double* n = &R.normal[(size_t)c * 6 + 2 * e];
      const double* bf = &R.BF[((size_t)e * 2 + R.cell_flip[(size_t)c * 3 + e]) * NM * nQ1];
      for (int j = 0; j < NM; ++j) {
        double v = 0;
        for (int i = 0; i < nQ1; ++i)
          v += bf[j * nQ1 + i] * (n[0] * in[(size_t)c * nQ + i] + n[1] * in[(size_t)c * nQ + nQ1 + i]);
        cm[((size_t)c * 3 + e) * NM + j] = v * idet;
      }
    }
  }
#pragma omp parallel for schedule(static)
  for (int f = 0; f < nf; ++f)
    for (int j = 0; j < NM; ++j) {
      double v = 0;
      for (int s = 0; s < 2; ++s) {
        const int c = R.facet_cell[2 * f + s];
        if (c >= 0) v += cm[((size_t)c * 3 + R.facet_local[2 * f + s]) * NM + j];
      }
      W.t[(size_t)f * NM + j] = v;
    }
  x_solve(R, W.t.data(), W.mu.data(), W.r, W.d, W.w);
#pragma omp parallel for schedule(static)
  for (int c = 0; c < nc; ++c) {
    const double idet = 1.0 / R.detJ[c];
    double* o = &out[(size_t)c * nQ];
    for (int i = 0; i < nQ; ++i) o[i] = in[(size_t)c * nQ + i] * idet;
    for (int e = 0; e < 3; ++e) {
      const double* n = &R.normal[(size_t)c * 6 + 2 * e];
      const double* bf = &R.BF[((size_t)e * 2 + R.cell_flip[(size_t)c * 3 + e]) * NM * nQ1];
      const double* m = &W.mu[(size_t)R.cell_facet[(size_t)c * 3 + e] * NM];
      for (int i = 0; i < nQ1; ++i) {
        double v = 0;
        for (int j = 0; j < NM; ++j) v += bf[j * nQ1 + i] * m[j];
        o[i] -= idet * n[0] * v;
        o[nQ1 + i] -= idet * n[1] * v;
      }
    }
  }
}

// right-preconditioned BiCGStab for a_tentative x = b; converged when ||b - A x|| <= rtol ||b||
int tentative_solve(Ref& R, const double* Qstar, double adt, const double* b, double* x, double rtol, int maxit,
                    int* iters) {
  const size_t n = (size_t)R.nc * R.nQ;
  if (R.precond == 0 ? !tentative_blocks(R, Qstar, adt) : !x_setup(R, adt)) return 2;
  WoodburyScratch W(R);
  auto block_precond = [&](const Ref& R_, const double* in, double* out) {
    if (R_.precond == 0)
      ::block_precond(R_, in, out);
    else
      woodbury_precond(R_, in, out, W);
  };
  vec r(n), rh(n), p(n), v(n), s(n), t(n), ph(n), sh(n);
  tentative_apply(R, Qstar, adt, x, r.data());
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) r[i] = b[i] - r[i];
  rh = r;
  p = r;
  const double bb = dot(b, b, n);
  double rho = dot(rh.data(), r.data(), n), rr = rho;
  int it = 0;
  while (it < maxit && rr > rtol * rtol * bb) {
    block_precond(R, p.data(), ph.data());
    tentative_apply(R, Qstar, adt, ph.data(), v.data());
    const double alpha = rho / dot(rh.data(), v.data(), n);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) s[i] = r[i] - alpha * v[i];
    block_precond(R, s.data(), sh.data());
    tentative_apply(R, Qstar, adt, sh.data(), t.data());
    const double tt = dot(t.data(), t.data(), n);
    const double omega = tt > 0 ? dot(t.data(), s.data(), n) / tt : 0.0;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
      x[i] += alpha * ph[i] + omega * sh[i];
      r[i] = s[i] - omega * t[i];
    }
    const double rho1 = dot(rh.data(), r.data(), n);
    rr = dot(r.data(), r.data(), n);
    const double beta = (rho1 / rho) * (alpha / omega);
    rho = rho1;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) p[i] = r[i] + beta * (p[i] - omega * v[i]);
    ++it;
    if (!std::isfinite(rr) || omega == 0.0) break;
  }
  if (iters) *iters = it;
  return rr <= rtol * rtol * bb ? 0 : 1;
}

// int_K psi div Q dx (HDGOracle.cell_divergence, the Chorin right-hand side hdg_implicit.py:145)
void cell_divergence(const Ref& R, const double* Q, double scale, double* out) {
  const int nQ1 = R.nQ1, nQ = R.nQ, np = R.np, nc = R.nc;
#pragma omp parallel for schedule(static)
  for (int c = 0; c < nc; ++c) {
    const double det = R.detJ[c];
    const double* ji = &R.Jinv[(size_t)c * 4];
    for (int a = 0; a < np; ++a) {
      double v = 0;
      for (int cc = 0; cc < 2; ++cc)
        for (int i = 0; i < nQ1; ++i)
          v += (ji[0 * 2 + cc] * R.Bref[((size_t)0 * np + a) * nQ1 + i] + ji[1 * 2 + cc] * R.Bref[((size_t)1 * np + a) * nQ1 + i]) *
               Q[(size_t)c * nQ + cc * nQ1 + i];
      out[(size_t)c * np + a] = scale * det * v;
    }
  }
}

}  // namespace

extern "C" {

void* hdgcpu_create(int k, int nc, int nf, const double* cell_xy, const int32_t* cell_facet, const int32_t* cell_flip,
                    const int32_t* facet_cell, const int32_t* facet_local, int nq, int nqf, int nint, const double* wq,
                    const double* phiQ, const double* dphiQ, const double* phiP, const double* wf, const double* phiQf,
                    const double* phiPf, const double* ell, const double* bdmF, const double* bdmI, const double* legN,
                    double tau, double alpha, int upwind, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  Ref* R = new Ref;
  R->k = k;
  R->nQ1 = (k + 2) * (k + 3) / 2;
  R->nQ = 2 * R->nQ1;
  R->np = (k + 1) * (k + 2) / 2;
  R->nl1 = k + 1;
  R->nl = 3 * (k + 1);
  R->nA = R->nQ + R->np;
  R->nq = nq;
  R->nqf = nqf;
  R->nint = nint;
  R->nfm = 3 * (k + 2);
  R->nc = nc;
  R->nf = nf;
  R->tau = tau;
  R->alpha = alpha;
  R->upwind = upwind;
  if (R->nl1 > 8 || R->nfm + nint != R->nQ) {
    delete R;
    return nullptr;
  }
  const int nQ1 = R->nQ1, np = R->np, nl1 = R->nl1;
  R->wq.assign(wq, wq + nq);
  R->phiQ.assign(phiQ, phiQ + (size_t)nQ1 * nq);
  R->dphiQ.assign(dphiQ, dphiQ + (size_t)nQ1 * nq * 2);
  R->phiP.assign(phiP, phiP + (size_t)np * nq);
  R->wf.assign(wf, wf + nqf);
  R->phiQf.assign(phiQf, phiQf + (size_t)3 * nQ1 * nqf);
  R->phiPf.assign(phiPf, phiPf + (size_t)3 * np * nqf);
  R->ell.assign(ell, ell + (size_t)2 * nl1 * nqf);
  R->bdmF.assign(bdmF, bdmF + (size_t)3 * (k + 2) * nQ1);
  R->bdmI.assign(bdmI, bdmI + (size_t)nint * 2 * nQ1);
  R->NM = k + 2;
  R->legN.assign(legN, legN + (size_t)2 * (k + 2) * nqf);
  R->xy.assign(cell_xy, cell_xy + (size_t)nc * 6);
  R->cell_facet.assign(cell_facet, cell_facet + (size_t)nc * 3);
  R->cell_flip.assign(cell_flip, cell_flip + (size_t)nc * 3);
  R->facet_cell.assign(facet_cell, facet_cell + (size_t)nf * 2);
  R->facet_local.assign(facet_local, facet_local + (size_t)nf * 2);
  geometry(*R);
  reference_tensors(*R);
  transpose_tables(*R);
  if (!condense(*R)) {
    delete R;
    return nullptr;
  }
  return R;
}

void hdgcpu_destroy(void* h) { delete static_cast<Ref*>(h); }

int hdgcpu_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// precond: 0 = cell-block-Jacobi, 1 = facet-multiplier (Woodbury) with `sweeps` Chebyshev sweeps
void hdgcpu_set_tentative_precond(void* h, int precond, int sweeps) {
  Ref& R = *static_cast<Ref*>(h);
  R.precond = precond;
  if (sweeps > 0) R.cheb_sweeps = sweeps;
}

// out = (M + adt alpha N^T N)^-1 in with the Chebyshev-approximated facet solve (test hook)
int hdgcpu_precond_apply(void* h, double adt, const double* in, double* out) {
  Ref& R = *static_cast<Ref*>(h);
  if (!x_setup(R, adt)) return 1;
  WoodburyScratch W(R);
  woodbury_precond(R, in, out, W);
  return 0;
}

int hdgcpu_project_bdm(void* h, const double* Q, double* Qstar) { return project_bdm(*static_cast<Ref*>(h), Q, Qstar); }

int hdgcpu_fimpl_apply(void* h, const double* Q, const double* Qstar, double* out) {
  Ref& R = *static_cast<Ref*>(h);
#pragma omp parallel
  {
    FScratch S(R);
#pragma omp for schedule(static)
    for (int c = 0; c < R.nc; ++c) fimpl_cell(R, Q, Qstar, c, S, &out[(size_t)c * R.nQ]);
  }
  return 0;
}

int hdgcpu_local_schur(void* h, double* SK) {
  Ref& R = *static_cast<Ref*>(h);
  std::memcpy(SK, R.SK.data(), R.SK.size() * sizeof(double));
  return 0;
}

int hdgcpu_tentative_solve(void* h, const double* Qstar, double adt, const double* b, double* x, double rtol, int maxit,
                           int* iters) {
  return tentative_solve(*static_cast<Ref*>(h), Qstar, adt, b, x, rtol, maxit, iters);
}

int hdgcpu_poisson_solve(void* h, const double* Ru, const double* Rp, const double* Rl, double rtol, int maxit, double* u,
                         double* p, double* lam, int* iters) {
  return poisson_solve(*static_cast<Ref*>(h), Ru, Rp, Rl, rtol, maxit, u, p, lam, iters);
}

// one Chorin timestep (hdg_implicit.py:92-190, use_projection_method=True).  f = the interpolated forcing of this
// step (hdg_implicit.py:100).  Q, p are updated in place; iters = {BiCGStab, CG} iteration counts.
int hdgcpu_chorin_step(void* h, double* Q, double* p, const double* f, double dt, double rtol, int maxit, int* iters) {
  Ref& R = *static_cast<Ref*>(h);
  const int nQ = R.nQ, np = R.np, nc = R.nc, nf = R.nf;
  const size_t n = (size_t)nc * nQ;
  const double t0 = now();
  vec Qstar(n), rhs(n), Qt(n), Rp((size_t)nc * np), u(n), phi((size_t)nc * np), lam((size_t)nf * R.nl1);
  int rc = project_bdm(R, Q, Qstar.data());  // :98
  const double t1 = now();
  if (rc) return 10 + rc;
  // b_rhs_tentative = (Q, w) + dt (f, w)   :126
#pragma omp parallel
  {
    vec m(nQ), tmp(nQ);
#pragma omp for schedule(static)
    for (int c = 0; c < nc; ++c) {
      for (int i = 0; i < nQ; ++i) tmp[i] = Q[(size_t)c * nQ + i] + dt * f[(size_t)c * nQ + i];
      mass_cell(R, tmp.data(), c, m.data());
      for (int i = 0; i < nQ; ++i) rhs[(size_t)c * nQ + i] = m[i];
    }
  }
  Qt.assign(Q, Q + n);  // initial guess: the velocity of the previous step
  rc = tentative_solve(R, Qstar.data(), dt, rhs.data(), Qt.data(), rtol, maxit, iters ? &iters[0] : nullptr);  // :129
  const double t2 = now();
  if (rc) return 20 + rc;
  cell_divergence(R, Qt.data(), -1.0 / dt, Rp.data());  // :145
  rc = poisson_solve(R, nullptr, Rp.data(), nullptr, rtol, 100 * maxit, u.data(), phi.data(), lam.data(),
                     iters ? &iters[1] : nullptr);  // :146
  const double t3 = now();
  if (rc) return 30 + rc;
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) Q[i] = Qt[i] + dt * u[i];  // :150
  std::memcpy(p, phi.data(), phi.size() * sizeof(double));  // :188-190 (poisson_solve already removed the mean)
  const double t4 = now();
  R.tm.t[0] += t4 - t0;
  R.tm.t[1] += t1 - t0;
  R.tm.t[2] += t2 - t1;
  R.tm.t[3] += t3 - t2;
  for (int i = 0; i < 4; ++i) R.tm.n[i]++;
  return 0;
}

// accumulated seconds / calls of the reference's timer labels: timestep, bdm_projection, tentative_velocity_solve,
// pressure_solve
void hdgcpu_timers(void* h, double* seconds4, int64_t* calls4) {
  Ref& R = *static_cast<Ref*>(h);
  for (int i = 0; i < 4; ++i) {
    seconds4[i] = R.tm.t[i];
    calls4[i] = R.tm.n[i];
  }
}

}  // extern "C"
