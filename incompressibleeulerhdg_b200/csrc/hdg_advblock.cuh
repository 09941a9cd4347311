// Cell-block advection preconditioner of the tentative-velocity solve (default since round 2;
// hdg_set_tuning("tent_cellblock", 0) switches it off).
//
// The facet-multiplier preconditioner of hdg_tent.cuh removes the stiff normal-jump penalty but leaves the
// advection operator  I - a F0,  F0 = M^-1 f_impl(.;Q*) with alpha = 0  (hdg_imex.py:313-331), untouched;
// at the advective CFL numbers the timesteppers run at, BiCGStab then needs 20-90 iterations.  Composing it
// with the inverse of the cell-diagonal blocks of  I - a F0
//
//     Phat^-1  ->  Phat^-1 diag(C, I) ,    C_K = [ (I - a F0)_KK ]^-1        (one NQ1 x NQ1 block per cell,
//                                                                              the same for both components)
//
// halves the iteration count in the CPU model of the solver (tests/experiments/tent_precond_model.py:
// 88 -> 48 cold, 31 -> 16 warm-started at nx = 12, k = 2, CFL 0.32).  C is a right preconditioner: the
// Krylov residual is still the residual of the unmodified system, so any nonsingular C leaves the converged
// solution unchanged.
//
// The symmetric part of (I - a F0)_KK is  I + a int_dK |s| phi_i phi_j - a/2 int_K div(Q*) phi_i phi_j  (the
// volume term integrates by parts to 1/2 int_dK s phi_i phi_j, the flux term is s/2 - |s|; central flux: without
// the |s| term): positive definite as long as a |div Q*| < 2, which holds with a wide margin for the (weakly
// solenoidal) BDM-projected Q* of the timesteppers.  Gauss-Jordan elimination without pivoting is then stable
// (tests/test_advblock_host.py checks the definiteness on a Taylor-Green step).
//
// The per-cell bodies are plain functions of the cell index so that tests/host_kernels can compile them with
// g++ and check them against the oracle on the CPU; the __global__ wrappers only exist under nvcc.
#pragma once
#include "hdg_local.cuh"

// blk[(i * NQ1 + j) * nc + cell] = (I - a F0)_KK [i][j]   (row i = test function, column j = trial function)
template <int K, bool UPWIND>
__device__ __forceinline__ void advblock_build_cell(const double* __restrict__ xy, const int* __restrict__ nbr,
                                                    int nc, int cell, const double* __restrict__ Qstar, double adt,
                                                    double* __restrict__ blk) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NQ = T::NQ, NQF = T::NQF;
  Geo g = make_geo(xy, nc, cell);
  double Qh[2][NQ1];  // pulled-back advecting velocity  Qh = J^-1 Q*
  HDG_UNROLL
  for (int i = 0; i < NQ1; ++i) {
    double q0 = Qstar[(size_t)i * nc + cell], q1 = Qstar[(size_t)(NQ1 + i) * nc + cell];
    Qh[0][i] = g.Ji[0][0] * q0 + g.Ji[0][1] * q1;
    Qh[1][i] = g.Ji[1][0] * q0 + g.Ji[1][1] * q1;
  }
  // volume term  -sum_q WQ[q] phi_i(q) (Qh . grad^ phi_j)(q):  aq[q][d] = -WQ[q] Qh_d(q)
  double aq[NQ][2];
  HDG_UNROLL
  for (int q = 0; q < NQ; ++q) {
    double a0 = 0.0, a1 = 0.0;
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i) {
      a0 = fma(T::PHI(q, i), Qh[0][i], a0);
      a1 = fma(T::PHI(q, i), Qh[1][i], a1);
    }
    aq[q][0] = -T::WQ(q) * a0;
    aq[q][1] = -T::WQ(q) * a1;
  }
  // own-side part of the flux term on interior facets:  WF[q] |e|/detJ (s/2 - [upwind]|s|),  s = Q*.n
  double cf[3][NQF];
  const double J00 = g.Ji[1][1] * g.detJ, J01 = -g.Ji[0][1] * g.detJ, J10 = -g.Ji[1][0] * g.detJ,
               J11 = g.Ji[0][0] * g.detJ;
  HDG_UNROLL
  for (int e = 0; e < 3; ++e) {
    const bool interior = nbr[(size_t)e * nc + cell] >= 0;
    const double m0 = g.n[e][0] * J00 + g.n[e][1] * J10, m1 = g.n[e][0] * J01 + g.n[e][1] * J11;
    const double scale = g.le[e] * g.idetJ;
    HDG_UNROLL
    for (int q = 0; q < NQF; ++q) {
      double s = 0.0;
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) s = fma(T::PHIF(e, q, i), m0 * Qh[0][i] + m1 * Qh[1][i], s);
      const double coef = 0.5 * s - (UPWIND ? fabs(s) : 0.0);
      cf[e][q] = interior ? T::WF(q) * scale * coef : 0.0;
    }
  }
#ifdef __CUDA_ARCH__
#pragma unroll 1
#endif
  for (int j = 0; j < NQ1; ++j) {  // column j: the operator applied to the j-th basis function (table loads are warp-uniform)
    double col[NQ1];
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i) col[i] = 0.0;
    HDG_UNROLL
    for (int q = 0; q < NQ; ++q) {
      const double v = aq[q][0] * T::DPHI(0, q, j) + aq[q][1] * T::DPHI(1, q, j);
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) col[i] = fma(T::PHI(q, i), v, col[i]);
    }
    HDG_UNROLL
    for (int e = 0; e < 3; ++e)
      HDG_UNROLL
      for (int q = 0; q < NQF; ++q) {
        const double v = cf[e][q] * T::PHIF(e, q, j);
        HDG_UNROLL
        for (int i = 0; i < NQ1; ++i) col[i] = fma(T::PHIF(e, q, i), v, col[i]);
      }
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i)
      blk[(size_t)(i * NQ1 + j) * nc + cell] = ((i == j) ? 1.0 : 0.0) - adt * col[i];
  }
}

// in-place inverse of the N x N block of one cell, Gauss-Jordan without pivoting (see the header comment).
// N <= 10 (k <= 2): fully unrolled, the block lives in registers; larger blocks: rolled loops on a local array.
// The inverse is also stored rounded to FP32 (blk32): the apply kernel reads that copy and does its arithmetic in
// FP64, which halves the bytes of the only HBM-bound kernel of this preconditioner.  Rounding the *entries* of C
// only replaces C by a slightly different fixed matrix -- the preconditioner stays an exactly linear operator, so
// right-preconditioned BiCGStab and the converged solution are unaffected (unlike FP32 *vectors*).
// sK (optional): sK[cell] = tr(C_K) / N, the cell-wise scalar of the scaled Schur complement (hdg_tent.cuh)
template <int N>
__device__ __forceinline__ void advblock_invert_cell(int nc, int cell, double* __restrict__ blk,
                                                     float* __restrict__ blk32, double* __restrict__ sK = nullptr) {
  double A[N * N];
  if constexpr (N <= 10) {
    HDG_UNROLL
    for (int i = 0; i < N * N; ++i) A[i] = blk[(size_t)i * nc + cell];
    HDG_UNROLL
    for (int p = 0; p < N; ++p) {
      const double piv = 1.0 / A[p * N + p];
      A[p * N + p] = 1.0;
      HDG_UNROLL
      for (int c = 0; c < N; ++c) A[p * N + c] *= piv;
      HDG_UNROLL
      for (int r = 0; r < N; ++r) {
        if (r == p) continue;
        const double f = A[r * N + p];
        A[r * N + p] = 0.0;
        HDG_UNROLL
        for (int c = 0; c < N; ++c) A[r * N + c] = fma(-f, A[p * N + c], A[r * N + c]);
      }
    }
    HDG_UNROLL
    for (int i = 0; i < N * N; ++i) {
      blk[(size_t)i * nc + cell] = A[i];
      blk32[(size_t)i * nc + cell] = (float)A[i];
    }
    if (sK) {
      double tr = 0.0;
      HDG_UNROLL
      for (int i = 0; i < N; ++i) tr += A[i * N + i];
      sK[cell] = tr / N;
    }
  } else {
    for (int i = 0; i < N * N; ++i) A[i] = blk[(size_t)i * nc + cell];
    for (int p = 0; p < N; ++p) {
      const double piv = 1.0 / A[p * N + p];
      A[p * N + p] = 1.0;
      for (int c = 0; c < N; ++c) A[p * N + c] *= piv;
      for (int r = 0; r < N; ++r) {
        if (r == p) continue;
        const double f = A[r * N + p];
        A[r * N + p] = 0.0;
        for (int c = 0; c < N; ++c) A[r * N + c] = fma(-f, A[p * N + c], A[r * N + c]);
      }
    }
    for (int i = 0; i < N * N; ++i) {
      blk[(size_t)i * nc + cell] = A[i];
      blk32[(size_t)i * nc + cell] = (float)A[i];
    }
    if (sK) {
      double tr = 0.0;
      for (int i = 0; i < N; ++i) tr += A[i * N + i];
      sK[cell] = tr / N;
    }
  }
}

// Y_c = C_K X_c for both velocity components of one cell (X, Y: SoA velocity fields [2 NQ1][nc]; Y must not alias X);
// blk is the FP32-stored inverse, the arithmetic is FP64
// (S = double; with S = float -- the mixed-precision solver -- vectors and arithmetic are FP32 like the stored block)
template <int K, typename S = double>
__device__ __forceinline__ void advblock_apply_cell(int nc, int cell, const float* __restrict__ blk,
                                                    const S* __restrict__ X, S* __restrict__ Y) {
  constexpr int NQ1 = Dims<K>::NQ1;
  S x[2][NQ1];
  HDG_UNROLL
  for (int c = 0; c < 2; ++c)
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i) x[c][i] = X[(size_t)(c * NQ1 + i) * nc + cell];
  HDG_UNROLL
  for (int i = 0; i < NQ1; ++i) {
    S y0 = 0, y1 = 0;
    HDG_UNROLL
    for (int j = 0; j < NQ1; ++j) {
      const S a = (S)blk[(size_t)(i * NQ1 + j) * nc + cell];
      y0 = fma(a, x[0][j], y0);
      y1 = fma(a, x[1][j], y1);
    }
    Y[(size_t)i * nc + cell] = y0;
    Y[(size_t)(NQ1 + i) * nc + cell] = y1;
  }
}

#ifdef __CUDACC__
template <int K, bool UPWIND>
__global__ void __launch_bounds__(128) k_advblock_build(const double* __restrict__ xy, const int* __restrict__ nbr,
                                                        int nc, const double* __restrict__ Qstar, double adt,
                                                        double* __restrict__ blk) {
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x)
    advblock_build_cell<K, UPWIND>(xy, nbr, nc, cell, Qstar, adt, blk);
}

template <int K>
__global__ void __launch_bounds__(64) k_advblock_invert(int nc, double* __restrict__ blk, float* __restrict__ blk32,
                                                        double* __restrict__ sK) {
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x)
    advblock_invert_cell<Dims<K>::NQ1>(nc, cell, blk, blk32, sK);
}

template <int K, typename S = double>
__global__ void __launch_bounds__(128) k_advblock_apply(int nc, const float* __restrict__ blk,
                                                        const S* __restrict__ X, S* __restrict__ Y) {
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x)
    advblock_apply_cell<K>(nc, cell, blk, X, Y);
}
#endif
