// Penalty-robust solver for the tentative-velocity system
//     (I - a M^-1 f_impl(.;Q*)) x = b ,   a = a_ii dt        (hdg_imex.py:233-255, hdg_implicit.py:103-129)
// which the reference hands to GMRES+ILU / a direct LU.
//
// f_impl contains the normal-jump penalty  -alpha/h_F int_F [[x.n]][[w.n]]  (hdg_imex.py:319-323).
// With L2-orthonormal Legendre moments on every facet,  N x |_F,j = int_0^1 [[x.n]] l_j ds
// (j <= k+1, exact for the degree-(k+1) normal trace), the penalty is exactly  alpha N^T N  because
// h_F = |F| (common.py:36-57).  Its weight relative to the mass matrix grows like a/h^2, which is
// what makes an unpreconditioned Krylov method stall on fine meshes.  The stiff term is removed by
// introducing the facet multiplier  mu = a alpha N x:
//
//     [ I - a F0     M^-1 N^T      ] [ x  ]   [ b ]        F0 = M^-1 f_impl with alpha = 0
//     [ N           -1/(a alpha) I ] [ mu ] = [ 0 ]
//
// and right-preconditioning BiCGStab with the same matrix without the advection part, whose Schur
// complement  X = 1/(a alpha) I + N M^-1 N^T  is a facet "mass" matrix: symmetric positive
// definite, 5 blocks of (k+2)^2 per row, spectrum of the facet-block-Jacobi preconditioned X inside
// [0.25, 1.8] independently of h, k, a (checked in tests).  X^-1 is replaced by a fixed number of
// Chebyshev/facet-block-Jacobi sweeps, applied matrix-free: every block is a compile-time reference
// table GG scaled by (n_e . n_e') / detJ.  In this formulation an inexact X only perturbs the
// preconditioner by O(eps); the stiffness never multiplies the error.
#pragma once
#include "hdg_local.cuh"

template <int K>
struct TentDims {
  static constexpr int NM = K + 2;                 // normal-moment modes per facet
  static constexpr int NMH = NM * (NM + 1) / 2;
};

// per-facet geometric coefficients of X (setup, geometry only):
//   tc[(3 s + j) nf + f] = (n_e . n_{(e+j)%3}) / detJ  of the cell on side s (0 on a missing side)
//   tcol[(2 s + j - 1) nf + f] = facet (e+j)%3 of that cell, j = 1, 2        (f itself if missing)
//   tbits[f] bit (3 s + j) = orientation flip of facet (e+j)%3 in that cell
__global__ void k_tent_setup(const double* __restrict__ xy, const int* __restrict__ cell_facet,
                             const int* __restrict__ cell_flip, const int* __restrict__ facet_cell,
                             const int* __restrict__ facet_local, int nc, int nf, double* __restrict__ tc,
                             int* __restrict__ tcol, int* __restrict__ tbits) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    int bits = 0;
    for (int s = 0; s < 2; ++s) {
      int cell = facet_cell[(size_t)s * nf + f];
      int e = facet_local[(size_t)s * nf + f];
      if (cell < 0) {
        for (int j = 0; j < 3; ++j) tc[(size_t)(3 * s + j) * nf + f] = 0.0;
        for (int j = 1; j < 3; ++j) tcol[(size_t)(2 * s + j - 1) * nf + f] = f;
        continue;
      }
      Geo g = make_geo(xy, nc, cell);
      for (int j = 0; j < 3; ++j) {
        int e2 = (e + j) % 3;
        tc[(size_t)(3 * s + j) * nf + f] = (g.n[e][0] * g.n[e2][0] + g.n[e][1] * g.n[e2][1]) * g.idetJ;
        if (cell_flip[(size_t)e2 * nc + cell]) bits |= 1 << (3 * s + j);
        if (j > 0) tcol[(size_t)(2 * s + j - 1) * nf + f] = cell_facet[(size_t)e2 * nc + cell];
      }
    }
    tbits[f] = bits;
  }
}

// cm[(e NM + j) nc + cell] = sigma int_0^1 (y . n_e) l_j ds  in the global facet orientation
// S (here and below) = storage type of the solver vectors: double, or float inside the mixed-precision solver
// (run_tentative_mixed in hdg_engine.cu); the arithmetic of these bandwidth-bound kernels stays FP64
template <int K, typename S = double>
__global__ void __launch_bounds__(128) k_tent_moments(const double* __restrict__ xy, const int* __restrict__ flip,
                                                      int nc, const S* __restrict__ Y, S* __restrict__ cm) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NM = TentDims<K>::NM;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    double y[2][NQ1];
    HDG_UNROLL
    for (int c = 0; c < 2; ++c)
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) y[c][i] = Y[(size_t)(c * NQ1 + i) * nc + cell];
    HDG_UNROLL
    for (int e = 0; e < 3; ++e) {
      int fl = flip[(size_t)e * nc + cell];
      double m[NM];
      HDG_UNROLL
      for (int j = 0; j < NM; ++j) m[j] = 0.0;
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) {
        double yn = g.n[e][0] * y[0][i] + g.n[e][1] * y[1][i];
        HDG_UNROLL
        for (int j = 0; j < NM; ++j)
          if (T::BF(e, j, i) != 0.0) m[j] = fma(T::BF(e, j, i), yn, m[j]);
      }
      HDG_UNROLL
      for (int j = 0; j < NM; ++j) cm[(size_t)(e * NM + j) * nc + cell] = (S)(flip_sign(fl, j) * m[j]);
    }
  }
}

// t = N y_x - y_mu  (y_mu may be null); optionally also sum = N y_x
template <int K, typename S = double>
__global__ void __launch_bounds__(256) k_tent_trhs(const S* __restrict__ cm, const int* __restrict__ facet_cell,
                                                   const int* __restrict__ facet_local, int nc, int nf,
                                                   const S* __restrict__ ymu, S* __restrict__ t, S* __restrict__ nyx) {
  constexpr int NM = TentDims<K>::NM;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    int c0 = facet_cell[f], c1 = facet_cell[(size_t)nf + f];
    int e0 = facet_local[f], e1 = facet_local[(size_t)nf + f];
    HDG_UNROLL
    for (int j = 0; j < NM; ++j) {
      double v = cm[(size_t)(e0 * NM + j) * nc + c0];
      if (c1 >= 0) v += cm[(size_t)(e1 * NM + j) * nc + c1];
      if (nyx) nyx[(size_t)j * nf + f] = (S)v;
      if (ymu) v -= ymu[(size_t)j * nf + f];
      t[(size_t)j * nf + f] = (S)v;
    }
  }
}

// The reference blocks GG(e, e') = BF_e BF_e'^T are Gram matrices of the facet functionals in an L2-orthonormal basis of
// the full space P_{k+1}, hence independent of that basis and invariant under the affine maps of the reference triangle
// onto itself.  The cyclic vertex permutation maps facet e to e + 1 with its parametrisation (from vertex e + 1 to e + 2)
// and has unit Jacobian determinant, so
//     GG(e, (e + j) % 3) = GG(0, j)   for every e,   and   GG(e, e) is diagonal (Legendre moments of one facet)
// (checked on the tables by tests/test_tent_host.py::test_gram_blocks_do_not_depend_on_the_local_facet).  The sweeps use
// this: no switch over the local facet index.  Facets of one warp carry 2-3 different local indices on every mesh
// numbering, so the switch executed its body 2-3 times per side (ncu r2l: 820 instructions per facet, issue slots 52 %
// busy in a kernel that should only wait for memory).
template <int K>
__host__ __device__ constexpr bool tent_gram_diagonal() {
  for (int j = 0; j < TentDims<K>::NM; ++j)
    for (int l = 0; l < TentDims<K>::NM; ++l)
      if (j != l && RefTables<K>::GG(0, 0, j, l) != 0.0) return false;
  return true;
}

// contribution of the cell on one side of facet f to the off-diagonal part of (G mu)_f.  All global loads (geometric
// coefficients c[1..2], neighbour values v[0..1][.]) are issued by the caller before the arithmetic so that the
// independent loads of a facet are in flight together; a missing side has zero coefficients and contributes nothing.
template <int K, typename V>  // V: register type of the neighbour values (float in k_tent_sweep32)
__device__ __forceinline__ void tent_side(int fl0, const int (&fl)[2], const double (&c)[3],
                                          const V (&v)[2][TentDims<K>::NM], double (&acc)[TentDims<K>::NM]) {
  using T = RefTables<K>;
  constexpr int NM = TentDims<K>::NM;
  HDG_UNROLL
  for (int jj = 1; jj < 3; ++jj) {
    double w[NM];
    HDG_UNROLL
    for (int l = 0; l < NM; ++l) w[l] = flip_sign(fl[jj - 1], l) * (double)v[jj - 1][l];
    HDG_UNROLL
    for (int j = 0; j < NM; ++j) {
      double sum = 0.0;
      HDG_UNROLL
      for (int l = 0; l < NM; ++l)
        if (T::GG(0, jj, j, l) != 0.0) sum = fma(T::GG(0, jj, j, l), w[l], sum);
      acc[j] = fma(c[jj] * flip_sign(fl0, j), sum, acc[j]);
    }
  }
}

// diagonal of X = inv_aalpha I + G on facet f: inv_aalpha + (c_0[0] + c_1[0]) GG(0, 0, j, j)   (sign flips squared)
template <int K>
__device__ __forceinline__ void tent_diag(double inv_aalpha, double c00, double c10, double (&D)[TentDims<K>::NM]) {
  static_assert(tent_gram_diagonal<K>(), "GG(e, e) is expected to be diagonal");
  HDG_UNROLL
  for (int j = 0; j < TentDims<K>::NM; ++j) D[j] = fma(c00 + c10, RefTables<K>::GG(0, 0, j, j), inv_aalpha);
}

// One sweep on X = inv_aalpha I + G, matrix-free.
//   mode 0: Chebyshev / facet-block-Jacobi:  r = rhs - X x (x == 0 if zero), d = cd d + cr D^-1 r,
//           xout = x + d                                                  (xout must not alias x)
//   mode 1: residual  xout = rhs + rhs2 - X x
//   mode 2: xout = D^-1 X x   (power iteration for the spectral bound)
// MINB = resident CTAs per SM the register allocation is tuned for (5: 96 registers at
// k = 2; 6: 80 registers; 8: 64 registers with a 120-byte spill) -- selected at run time by
// hdg_set_tuning("sweep_minblocks") so that the variants can be compared on the same box.
template <int K, int MINB, typename S = double>
__global__ void __launch_bounds__(128, MINB) k_tent_sweep(int nf, const int* __restrict__ facet_local,
                                                    const double* __restrict__ tc, const int* __restrict__ tcol,
                                                    const int* __restrict__ tbits, double inv_aalpha,
                                                    const S* __restrict__ rhs, const S* __restrict__ rhs2,
                                                    const S* __restrict__ x, S* __restrict__ d,
                                                    S* __restrict__ xout, double cd, double cr, int zero,
                                                    int mode) {
  constexpr int NM = TentDims<K>::NM, NMH = TentDims<K>::NMH;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    // ---- all loads first ------------------------------------------------------------------------
    const int bits = tbits[f];
    int col[2][2];
    double c[2][3], v[2][2][NM], own[NM], b[NM], dprev[NM];
    HDG_UNROLL
    for (int s = 0; s < 2; ++s) {
      HDG_UNROLL
      for (int j = 0; j < 3; ++j) c[s][j] = tc[(size_t)(3 * s + j) * nf + f];  // zero on a missing side
      HDG_UNROLL
      for (int jj = 0; jj < 2; ++jj) col[s][jj] = tcol[(size_t)(2 * s + jj) * nf + f];  // f itself if missing
    }
    HDG_UNROLL
    for (int j = 0; j < NM; ++j) {
      own[j] = zero ? 0.0 : x[(size_t)j * nf + f];
      b[j] = 0.0;
      if (mode != 2) {
        b[j] = rhs[(size_t)j * nf + f];
        if (mode == 1 && rhs2) b[j] += rhs2[(size_t)j * nf + f];
      }
      dprev[j] = (mode == 0 && cd != 0.0) ? d[(size_t)j * nf + f] : 0.0;
    }
    HDG_UNROLL
    for (int s = 0; s < 2; ++s)
      HDG_UNROLL
      for (int jj = 0; jj < 2; ++jj)
        HDG_UNROLL
        for (int l = 0; l < NM; ++l) v[s][jj][l] = zero ? 0.0 : x[(size_t)l * nf + col[s][jj]];
    // ---- arithmetic -----------------------------------------------------------------------------
    double acc[NM], D[NM];
    HDG_UNROLL
    for (int j = 0; j < NM; ++j) acc[j] = 0.0;
    tent_diag<K>(inv_aalpha, c[0][0], c[1][0], D);
    if (!zero) {
      HDG_UNROLL
      for (int s = 0; s < 2; ++s) {
        const int fl[2] = {(bits >> (3 * s + 1)) & 1, (bits >> (3 * s + 2)) & 1};
        tent_side<K>((bits >> (3 * s)) & 1, fl, c[s], v[s], acc);
      }
    }
    // X x = D own + off-diagonal part
    double r[NM];
    HDG_UNROLL
    for (int j = 0; j < NM; ++j) {
      const double w = zero ? 0.0 : fma(D[j], own[j], acc[j]);
      r[j] = (mode != 2) ? b[j] - w : w;
    }
    if (mode == 1) {
      HDG_UNROLL
      for (int j = 0; j < NM; ++j) xout[(size_t)j * nf + f] = (S)r[j];
      continue;
    }
    HDG_UNROLL
    for (int j = 0; j < NM; ++j) r[j] /= D[j];
    if (mode == 2) {
      HDG_UNROLL
      for (int j = 0; j < NM; ++j) xout[(size_t)j * nf + f] = (S)r[j];
      continue;
    }
    HDG_UNROLL
    for (int j = 0; j < NM; ++j) {
      double di = fma(cd, dprev[j], cr * r[j]);
      d[(size_t)j * nf + f] = (S)di;
      xout[(size_t)j * nf + f] = (S)(own[j] + di);
    }
  }
}

// FP32-*stored* variant of the Chebyshev / facet-block-Jacobi sweep (mode 0 of k_tent_sweep; experimental,
// hdg_set_tuning "tent_fp32", only together with "tent_flex"): the iterate x, the correction d and the output are
// float arrays, the arithmetic stays FP64, the right-hand side stays FP64.  172 instead of 236 bytes per facet.  With
// rounded vectors the preconditioner is no longer an exactly linear operator, which only the flexible solution update
// of BiCGStab tolerates (tests/experiments/tent_fp32_sweeps.py).  The last sweep of a solve writes the multiplier in
// FP64 (xout64) for k_tent_xhat and the residual row of the operator.
// S = storage type of the right-hand side and of the final multiplier (double; float in the mixed-precision solver)
template <int K, int MINB, typename S = double>
__global__ void __launch_bounds__(128, MINB) k_tent_sweep32(int nf, const int* __restrict__ facet_local,
                                                      const double* __restrict__ tc, const int* __restrict__ tcol,
                                                      const int* __restrict__ tbits, double inv_aalpha,
                                                      const S* __restrict__ rhs, const float* __restrict__ x,
                                                      float* __restrict__ d, float* __restrict__ xout32,
                                                      S* __restrict__ xout64, double cd, double cr, int zero) {
  constexpr int NM = TentDims<K>::NM, NMH = TentDims<K>::NMH;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    const int bits = tbits[f];
    int col[2][2];
    // the FP32-stored values stay float in registers until they are used (every register saved while the loads are in
    // flight buys resident warps)
    double c[2][3], b[NM];
    float v[2][2][NM], own[NM], dprev[NM];
    HDG_UNROLL
    for (int s = 0; s < 2; ++s) {
      HDG_UNROLL
      for (int j = 0; j < 3; ++j) c[s][j] = tc[(size_t)(3 * s + j) * nf + f];
      HDG_UNROLL
      for (int jj = 0; jj < 2; ++jj) col[s][jj] = tcol[(size_t)(2 * s + jj) * nf + f];
    }
    HDG_UNROLL
    for (int j = 0; j < NM; ++j) {
      own[j] = zero ? 0.0f : x[(size_t)j * nf + f];
      b[j] = rhs[(size_t)j * nf + f];
      dprev[j] = (cd != 0.0) ? d[(size_t)j * nf + f] : 0.0f;
    }
    HDG_UNROLL
    for (int s = 0; s < 2; ++s)
      HDG_UNROLL
      for (int jj = 0; jj < 2; ++jj)
        HDG_UNROLL
        for (int l = 0; l < NM; ++l) v[s][jj][l] = zero ? 0.0f : x[(size_t)l * nf + col[s][jj]];
    double acc[NM], D[NM], r[NM];
    HDG_UNROLL
    for (int j = 0; j < NM; ++j) acc[j] = 0.0;
    tent_diag<K>(inv_aalpha, c[0][0], c[1][0], D);
    if (!zero) {
      HDG_UNROLL
      for (int s = 0; s < 2; ++s) {
        const int fl[2] = {(bits >> (3 * s + 1)) & 1, (bits >> (3 * s + 2)) & 1};
        tent_side<K>((bits >> (3 * s)) & 1, fl, c[s], v[s], acc);
      }
    }
    // D^-1 (rhs - X x) = D^-1 (rhs - offdiag x) - x
    HDG_UNROLL
    for (int j = 0; j < NM; ++j) r[j] = (b[j] - acc[j]) / D[j] - (double)own[j];
    HDG_UNROLL
    for (int j = 0; j < NM; ++j) {
      const double di = fma(cd, (double)dprev[j], cr * r[j]);
      d[(size_t)j * nf + f] = (float)di;
      if (xout64)
        xout64[(size_t)j * nf + f] = (S)((double)own[j] + di);
      else
        xout32[(size_t)j * nf + f] = (float)((double)own[j] + di);
    }
  }
}

// xh = y - s_K M^-1 N^T mu   (mode 0)   or   xh += y - s_K M^-1 N^T mu   (mode 1, final recovery);
// sK (optional) = the per-cell scalar of the scaled Schur complement (k_tent_scale_tc), 1 if null.
// Zout (optional, mode 0) = xh + M^-1 N^T mu = y + (1 - s_K) M^-1 N^T mu: what the velocity row of the augmented
// operator adds to -a F0(xh)  (with s_K = 1 this is y itself and the caller passes y instead).
template <int K, typename S = double>
__global__ void __launch_bounds__(128) k_tent_xhat(const double* __restrict__ xy, const int* __restrict__ flip,
                                                   const int* __restrict__ cell_facet, int nc, int nf,
                                                   const S* __restrict__ Y, const S* __restrict__ mu,
                                                   S* __restrict__ Xh, int mode,
                                                   const double* __restrict__ sK = nullptr,
                                                   S* __restrict__ Zout = nullptr) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NM = TentDims<K>::NM;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    const double sk = sK ? sK[cell] : 1.0;
    double a0[NQ1], a1[NQ1];
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i) a0[i] = a1[i] = 0.0;
    HDG_UNROLL
    for (int e = 0; e < 3; ++e) {
      int f = cell_facet[(size_t)e * nc + cell];
      int fl = flip[(size_t)e * nc + cell];
      double m[NM];
      HDG_UNROLL
      for (int j = 0; j < NM; ++j) m[j] = flip_sign(fl, j) * mu[(size_t)j * nf + f];
      double cx = g.idetJ * g.n[e][0], cy = g.idetJ * g.n[e][1];
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) {
        double v = 0.0;
        HDG_UNROLL
        for (int j = 0; j < NM; ++j)
          if (T::BF(e, j, i) != 0.0) v = fma(T::BF(e, j, i), m[j], v);
        a0[i] = fma(cx, v, a0[i]);
        a1[i] = fma(cy, v, a1[i]);
      }
    }
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i) {
      size_t i0 = (size_t)i * nc + cell, i1 = (size_t)(NQ1 + i) * nc + cell;
      const double y0 = Y[i0], y1 = Y[i1];
      double v0 = fma(-sk, a0[i], y0), v1 = fma(-sk, a1[i], y1);
      if (Zout) {
        Zout[i0] = (S)(v0 + a0[i]);
        Zout[i1] = (S)(v1 + a1[i]);
      }
      if (mode == 1) {
        v0 += Xh[i0];
        v1 += Xh[i1];
      }
      Xh[i0] = (S)v0;
      Xh[i1] = (S)v1;
    }
  }
}

// Scaled facet Schur complement.  With the cell-block advection preconditioner C = blockdiag(I - a F0)^-1
// (hdg_advblock.cuh) the exact Schur complement of the block preconditioner [[C^-1, M^-1 N^T], [N, -1/(a alpha)]] is
// 1/(a alpha) + N C M^-1 N^T, which is no longer a table times a geometric scalar.  Replacing C by its cell-wise mean
// diagonal  s_K = tr(C_K) / NQ1  inside the Schur complement and in the back-substitution keeps the structure of X --
// the same sweep kernel runs on rescaled coefficients tcs = tc * s_K -- and still carries the bulk of the effect: in a
// dense model of the augmented system (k = 2, 72 cells) FGMRES needs 62 / 110 / 302 / 640 applications at CFL
// 0.32 / 1 / 4 / 10, against 74 / 182 / 724 / > 1500 with the unscaled X and 46 / 77 / 196 / 361 with the exact
// combined operator.  The operator row  out_mu = N xh - mu / (a alpha)  stays exact, because xh is formed with the
// same s_K (k_tent_xhat) and the residual sweep runs on the same tcs.
__global__ void k_tent_scale_tc(int nf, const int* __restrict__ facet_cell, const double* __restrict__ tc,
                                const double* __restrict__ sK, double* __restrict__ tcs) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    for (int s = 0; s < 2; ++s) {
      const int cell = facet_cell[(size_t)s * nf + f];
      const double w = cell >= 0 ? sK[cell] : 0.0;
      for (int j = 0; j < 3; ++j) tcs[(size_t)(3 * s + j) * nf + f] = w * tc[(size_t)(3 * s + j) * nf + f];
    }
  }
}

// lam[cell] = lambda_max( blockdiag(G_K)^-1 G_K )  of the element matrix G_K = N_K M_K^-1 N_K^T (3 x 3 blocks of
// NM x NM, block (e, e') = (n_e . n_e') / detJ GG(e, e')), by power iteration.  For any positive cell weights w_K the
// facet-block-Jacobi preconditioned  1/(a alpha) + sum_K w_K P_K^T G_K P_K  has its spectrum below max_K lam[K]
// (element-wise bound), so this is the upper end of the Chebyshev interval that is safe for the scaled Schur
// complement of every solve.  Geometry only, once per engine.
template <int K>
__global__ void __launch_bounds__(64) k_tent_elem_bound(const double* __restrict__ xy, int nc, double* __restrict__ lam) {
  using T = RefTables<K>;
  constexpr int NM = TentDims<K>::NM, NMH = TentDims<K>::NMH;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    double c[3][3];
    HDG_UNROLL
    for (int e = 0; e < 3; ++e)
      HDG_UNROLL
      for (int e2 = 0; e2 < 3; ++e2) c[e][e2] = (g.n[e][0] * g.n[e2][0] + g.n[e][1] * g.n[e2][1]) * g.idetJ;
    double D[3][NMH];
    HDG_UNROLL
    for (int e = 0; e < 3; ++e) {
      HDG_UNROLL
      for (int j = 0; j < NM; ++j)
        HDG_UNROLL
        for (int l = 0; l <= j; ++l) D[e][tri(j, l)] = c[e][e] * T::GG(e, e, j, l);
      cholesky<NM>(D[e]);
    }
    double x[3][NM];
    HDG_UNROLL
    for (int e = 0; e < 3; ++e)
      HDG_UNROLL
      for (int j = 0; j < NM; ++j) x[e][j] = 1.0 + 0.37 * (e * NM + j) - 0.11 * (e * NM + j) * (e * NM + j);
    double lmax = 0.0;
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
    for (int it = 0; it < 60; ++it) {
      double y[3][NM], nx2 = 0.0, ny2 = 0.0;
      HDG_UNROLL
      for (int e = 0; e < 3; ++e) {
        HDG_UNROLL
        for (int j = 0; j < NM; ++j) {
          double v = 0.0;
          HDG_UNROLL
          for (int e2 = 0; e2 < 3; ++e2)
            HDG_UNROLL
            for (int l = 0; l < NM; ++l) v = fma(c[e][e2] * T::GG(e, e2, j, l), x[e2][l], v);
          y[e][j] = v;
          nx2 = fma(x[e][j], x[e][j], nx2);
        }
        chol_solve<NM>(D[e], y[e]);
        HDG_UNROLL
        for (int j = 0; j < NM; ++j) ny2 = fma(y[e][j], y[e][j], ny2);
      }
      lmax = sqrt(ny2 / nx2);
      const double inv = rsqrt(ny2);
      HDG_UNROLL
      for (int e = 0; e < 3; ++e)
        HDG_UNROLL
        for (int j = 0; j < NM; ++j) x[e][j] = y[e][j] * inv;
    }
    lam[cell] = lmax;
  }
}
