// Per-cell / per-facet kernels of the condensed mixed-Poisson path: static condensation,
// deterministic blocked-ELL gather, forward elimination, back-substitution.  They depend on hdg_local.cuh only, so
// tests/host_kernels can also compile them with g++ and execute them on the CPU against the oracle
// (tests/test_poisson_host.py); the engine launches them from hdg_engine.cu.
//
// Reference: the firedrake.SCPC static condensation of a_mixed_poisson, hdg_imex.py:123-135 (SCPC.initialize /
// SCPC.apply [FD-knowledge]).
#pragma once
#include "hdg_local.cuh"

// ------------------------------------------------------------------------------------------------
// K1+K2: local operator build + static condensation, one thread per cell
//   S_K = -tau G - E E^T / detJ + W H^-1 W^T,   W = E B^T / detJ + tau F,  H = T + B B^T / detJ
// (equal to D - C A^-1 B of hdg_imex.py:128-133 for the blocks of hdg_imex.py:123-127)
// ------------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ double W_entry(const Geo& g, const double (&nu)[3][2], double tau, int e, int m, int a) {
  using T = RefTables<K>;
  double v = 0.0;
  if (T::LL(e, 0, m, a) != 0.0) v = fma(nu[e][0], T::LL(e, 0, m, a), v);
  if (T::LL(e, 1, m, a) != 0.0) v = fma(nu[e][1], T::LL(e, 1, m, a), v);
  if (T::F(e, m, a) != 0.0) v = fma(tau, T::F(e, m, a), v);
  return v * g.le[e];
}

// W_entry is structurally zero where all three reference tables vanish
template <int K>
__device__ __forceinline__ constexpr bool W_nonzero(int e, int m, int a) {
  using T = RefTables<K>;
  return T::LL(e, 0, m, a) != 0.0 || T::LL(e, 1, m, a) != 0.0 || T::F(e, m, a) != 0.0;
}

template <int K>
__global__ void __launch_bounds__(128) k_condense(const double* __restrict__ xy, const int* __restrict__ flip, int nc,
                                                  double tau, double* __restrict__ SK) {
  using T = RefTables<K>;
  using D = Dims<K>;
  constexpr int NP = D::NP, NL1 = D::NL1, NL = D::NL;
  // Work reduction, kept per degree where it measured faster (profiles/condense_bench_r1r.jsonl, 10^6 cells):
  //  * K <= 2: the NL x NP matrix W is built once and kept in registers (54 doubles at k = 2) instead of
  //    being re-derived from the tables inside every dot product;
  //  * K <= 3: S_K = S_K^T is computed on and above the diagonal only and mirrored on store
  //    (k = 3: 0.895 -> 0.525 ms; k = 1: 0.072 -> 0.064 ms; k = 2 unchanged at 0.19 ms: latency bound at
  //    8 warps/SM, not FP64-issue bound).  At k = 4 the triangular loop nest made the register allocation
  //    worse (5.8 -> 12.7 ms), so k = 4 keeps the full loop nest.
  constexpr bool SYM = (K <= 3);
  constexpr bool CACHE_W = (K <= 2);
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    double L[D::NH];
    build_H<K>(g, tau, L);
    cholesky<NP>(L);
    double nu[3][2];
    int fl[3];
    HDG_UNROLL
    for (int e = 0; e < 3; ++e) {
      fl[e] = flip[(size_t)e * nc + cell];
      nu[e][0] = g.Ji[0][0] * g.n[e][0] + g.Ji[0][1] * g.n[e][1];
      nu[e][1] = g.Ji[1][0] * g.n[e][0] + g.Ji[1][1] * g.n[e][1];
    }
    double Wc[CACHE_W ? NL * NP : 1];
    if (CACHE_W) {
      HDG_UNROLL
      for (int e = 0; e < 3; ++e)
        HDG_UNROLL
        for (int m = 0; m < NL1; ++m)
          HDG_UNROLL
          for (int a = 0; a < NP; ++a) Wc[CACHE_W ? (e * NL1 + m) * NP + a : 0] = W_entry<K>(g, nu, tau, e, m, a);
    }
    auto W = [&](int e, int m, int a) -> double {
      return CACHE_W ? Wc[CACHE_W ? (e * NL1 + m) * NP + a : 0] : W_entry<K>(g, nu, tau, e, m, a);
    };
    HDG_UNROLL
    for (int e = 0; e < 3; ++e) {
      HDG_UNROLL
      for (int m = 0; m < NL1; ++m) {
        double v[NP];
        HDG_UNROLL
        for (int a = 0; a < NP; ++a) v[a] = W(e, m, a);
        chol_solve<NP>(L, v);
        double sg = flip_sign(fl[e], m);
        HDG_UNROLL
        for (int e2 = (SYM ? e : 0); e2 < 3; ++e2) {
          double nn = (g.n[e][0] * g.n[e2][0] + g.n[e][1] * g.n[e2][1]) * g.le[e] * g.le[e2] * g.idetJ;
          HDG_UNROLL
          for (int m2 = (SYM && e2 == e ? m : 0); m2 < NL1; ++m2) {
            double s = 0.0;
            HDG_UNROLL
            for (int a = 0; a < NP; ++a)
              if (W_nonzero<K>(e2, m2, a)) s = fma(v[a], W(e2, m2, a), s);
            if (T::NN(e, e2, m, m2) != 0.0) s = fma(-nn, T::NN(e, e2, m, m2), s);
            if (e == e2 && m == m2) s -= tau * g.le[e];
            s *= sg * flip_sign(fl[e2], m2);
            const int r = e * NL1 + m, c = e2 * NL1 + m2;
            SK[(size_t)(r * NL + c) * nc + cell] = s;
            if (SYM && c != r) SK[(size_t)(c * NL + r) * nc + cell] = s;
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K3: deterministic gather of S_K into the blocked-ELL trace matrix P = -S, one thread per facet.
// Row block f: slot 0 = (f,f) summed cell 0 then cell 1; slots 1,2 = the other facets of cell 0 in
// local order (e0+1)%3,(e0+2)%3; slots 3,4 = those of cell 1 (zero blocks pointing at f on the
// boundary).  No atomics: every entry has exactly one writer and a fixed summation order.
// Also inverts the diagonal block (facet-block-Jacobi, the ASMStarPC patches hdg_imex.py:143-152).
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(128) k_assemble(const double* __restrict__ SK, const int* __restrict__ cell_facet,
                                                  const int* __restrict__ facet_cell,
                                                  const int* __restrict__ facet_local, int nc, int nf,
                                                  double* __restrict__ val, int* __restrict__ col,
                                                  double* __restrict__ dinv) {
  using D = Dims<K>;
  constexpr int b = D::NL1, NL = D::NL;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    double diag[b][b];
    HDG_UNROLL
    for (int r = 0; r < b; ++r)
      HDG_UNROLL
      for (int c = 0; c < b; ++c) diag[r][c] = 0.0;
    col[f] = f;
    HDG_UNROLL
    for (int side = 0; side < 2; ++side) {
      int cell = facet_cell[(size_t)side * nf + f];
      int e0 = facet_local[(size_t)side * nf + f];
      if (cell < 0) {
        HDG_UNROLL
        for (int j = 1; j <= 2; ++j) {
          int slot = 2 * side + j;
          col[(size_t)slot * nf + f] = f;
          HDG_UNROLL
          for (int r = 0; r < b; ++r)
            HDG_UNROLL
            for (int c = 0; c < b; ++c) val[(size_t)((slot * b + r) * b + c) * nf + f] = 0.0;
        }
        continue;
      }
      HDG_UNROLL
      for (int j = 0; j < 3; ++j) {
        int e2 = (e0 + j) % 3;
        HDG_UNROLL
        for (int r = 0; r < b; ++r) {
          HDG_UNROLL
          for (int c = 0; c < b; ++c) {
            double s = -SK[(size_t)((e0 * b + r) * NL + e2 * b + c) * nc + cell];
            if (j == 0)
              diag[r][c] += s;
            else
              val[(size_t)(((2 * side + j) * b + r) * b + c) * nf + f] = s;
          }
        }
        if (j > 0) col[(size_t)(2 * side + j) * nf + f] = cell_facet[(size_t)e2 * nc + cell];
      }
    }
    HDG_UNROLL
    for (int r = 0; r < b; ++r)
      HDG_UNROLL
      for (int c = 0; c < b; ++c) val[(size_t)(r * b + c) * nf + f] = diag[r][c];
    // inverse of the SPD diagonal block by Gauss-Jordan (no pivoting needed)
    double inv[b][b];
    HDG_UNROLL
    for (int r = 0; r < b; ++r)
      HDG_UNROLL
      for (int c = 0; c < b; ++c) inv[r][c] = (r == c) ? 1.0 : 0.0;
    HDG_UNROLL
    for (int p = 0; p < b; ++p) {
      double ip = 1.0 / diag[p][p];
      HDG_UNROLL
      for (int c = 0; c < b; ++c) {
        diag[p][c] *= ip;
        inv[p][c] *= ip;
      }
      HDG_UNROLL
      for (int r = 0; r < b; ++r) {
        if (r == p) continue;
        double fct = diag[r][p];
        HDG_UNROLL
        for (int c = 0; c < b; ++c) {
          diag[r][c] = fma(-fct, diag[p][c], diag[r][c]);
          inv[r][c] = fma(-fct, inv[p][c], inv[r][c]);
        }
      }
    }
    HDG_UNROLL
    for (int r = 0; r < b; ++r)
      HDG_UNROLL
      for (int c = 0; c < b; ++c) dinv[(size_t)(r * b + c) * nf + f] = inv[r][c];
  }
}

// ------------------------------------------------------------------------------------------------
// a3: forward elimination, per cell:  gK = C_K A_K^-1 (Ru, Rp)   (SCPC.apply, first half)
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(128) k_forward(const double* __restrict__ xy, const int* __restrict__ flip, int nc,
                                                 double tau, const double* __restrict__ Ru,
                                                 const double* __restrict__ Rp, double* __restrict__ gK) {
  using D = Dims<K>;
  constexpr int NQ1 = D::NQ1, NP = D::NP, NL1 = D::NL1;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    double L[D::NH];
    build_H<K>(g, tau, L);
    cholesky<NP>(L);
    double u[2][NQ1], phi[NP], lam[3][NL1];
    HDG_UNROLL
    for (int c = 0; c < 2; ++c)
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) u[c][i] = Ru ? Ru[(size_t)(c * NQ1 + i) * nc + cell] : 0.0;
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) phi[a] = Rp ? Rp[(size_t)a * nc + cell] : 0.0;
    local_solve<K, false>(g, tau, L, lam, u, phi);
    HDG_UNROLL
    for (int e = 0; e < 3; ++e)
      HDG_UNROLL
      for (int m = 0; m < NL1; ++m) lam[e][m] = 0.0;
    apply_E<K>(g, u, 1.0, lam);
    apply_F<K>(g, phi, tau, lam);
    HDG_UNROLL
    for (int e = 0; e < 3; ++e) {
      int fl = flip[(size_t)e * nc + cell];
      HDG_UNROLL
      for (int m = 0; m < NL1; ++m) gK[(size_t)(e * NL1 + m) * nc + cell] = flip_sign(fl, m) * lam[e][m];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K5 / a6: back-substitution per cell  (u,phi) = A_K^-1 ((Ru,Rp) - B_K lam_K)
// ------------------------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(128) k_back(const double* __restrict__ xy, const int* __restrict__ flip,
                                              const int* __restrict__ cell_facet, int nc, int nf, double tau,
                                              const double* __restrict__ Ru, const double* __restrict__ Rp,
                                              const double* __restrict__ lamg, double* __restrict__ uo,
                                              double* __restrict__ po) {
  using D = Dims<K>;
  constexpr int NQ1 = D::NQ1, NP = D::NP, NL1 = D::NL1;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    double L[D::NH];
    build_H<K>(g, tau, L);
    cholesky<NP>(L);
    double u[2][NQ1], phi[NP], lam[3][NL1];
    HDG_UNROLL
    for (int e = 0; e < 3; ++e) {
      int f = cell_facet[(size_t)e * nc + cell];
      int fl = flip[(size_t)e * nc + cell];
      HDG_UNROLL
      for (int m = 0; m < NL1; ++m) lam[e][m] = flip_sign(fl, m) * lamg[(size_t)m * nf + f];
    }
    HDG_UNROLL
    for (int c = 0; c < 2; ++c)
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) u[c][i] = Ru ? Ru[(size_t)(c * NQ1 + i) * nc + cell] : 0.0;
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) phi[a] = Rp ? Rp[(size_t)a * nc + cell] : 0.0;
    local_solve<K, true>(g, tau, L, lam, u, phi);
    HDG_UNROLL
    for (int c = 0; c < 2; ++c)
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) uo[(size_t)(c * NQ1 + i) * nc + cell] = u[c][i];
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) po[(size_t)a * nc + cell] = phi[a];
  }
}

// ------------------------------------------------------------------------------------------------
// K5 fused with the update of the caller (north_star item 5; a6 + a7): back-substitution whose result never goes to
// memory as (u, phi) -- the Richardson / projection update of the timesteppers is applied in registers:
//     Qacc <- cq Qacc + cb Qbase + cu u        (Chorin  Q = Q~ + dt u, hdg_implicit.py:150:      cq = 0, cb = 1, cu = dt;
//                                               IMEX    Q_i += Q~ + a dt u, hdg_imex.py:580-587: cq = 1, cb = 1, cu = a dt)
//     pacc <- cp pacc + phi                    (Chorin  p = phi, hdg_implicit.py:188;  IMEX  p_i += phi: cp = 1)
// together with the partial sums of  int_K phi dx = detJ phi_0 / sqrt(2)  over the owned cells, which the pressure shift
// (_shift_pressure, hdg_imex.py:471-478; hdg_implicit.py:189-190) needs: k_shift then only touches mode 0 of pacc.
// Saves writing u, phi and re-reading u, phi, Q~ (and the pressure pass of k_pmean_partial) per solve.
// Launched with a grid-stride grid of G blocks of 128 threads; partial has G entries.
// ------------------------------------------------------------------------------------------------
struct BackUpdate {
  double cq, cb, cu, cp;
  const double* Qbase;  // may be null when cb == 0
  double* Qacc;
  double* pacc;
  double* partial;
  int nc_own;
};

__device__ __forceinline__ double block_reduce128(double v) {  // fixed tree, 128 threads; result valid in thread 0
  __shared__ double sm128[4];
  HDG_UNROLL
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm128[threadIdx.x >> 5] = v;
  __syncthreads();
  return (sm128[0] + sm128[1]) + (sm128[2] + sm128[3]);
}

template <int K>
__global__ void __launch_bounds__(128) k_back_update(const double* __restrict__ xy, const int* __restrict__ flip,
                                                     const int* __restrict__ cell_facet, int nc, int nf, double tau,
                                                     const double* __restrict__ Ru, const double* __restrict__ Rp,
                                                     const double* __restrict__ lamg, BackUpdate U) {
  using D = Dims<K>;
  constexpr int NQ1 = D::NQ1, NP = D::NP, NL1 = D::NL1;
  double acc = 0.0;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    double L[D::NH];
    build_H<K>(g, tau, L);
    cholesky<NP>(L);
    double u[2][NQ1], phi[NP], lam[3][NL1];
    HDG_UNROLL
    for (int e = 0; e < 3; ++e) {
      int f = cell_facet[(size_t)e * nc + cell];
      int fl = flip[(size_t)e * nc + cell];
      HDG_UNROLL
      for (int m = 0; m < NL1; ++m) lam[e][m] = flip_sign(fl, m) * lamg[(size_t)m * nf + f];
    }
    HDG_UNROLL
    for (int c = 0; c < 2; ++c)
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) u[c][i] = Ru ? Ru[(size_t)(c * NQ1 + i) * nc + cell] : 0.0;
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) phi[a] = Rp ? Rp[(size_t)a * nc + cell] : 0.0;
    local_solve<K, true>(g, tau, L, lam, u, phi);
    if (cell < U.nc_own) acc = fma(g.detJ, phi[0], acc);
    // all loads of the update first, then all stores: the pointers come out of a struct, so the compiler must assume
    // that a store to Qacc / pacc may change what a later load of Qbase returns and would serialise 2 NQ1 round trips
    const double* __restrict__ qb = U.Qbase;
    double* __restrict__ qa = U.Qacc;
    double* __restrict__ pa = U.pacc;
    HDG_UNROLL
    for (int c = 0; c < 2; ++c)
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) {
        const size_t idx = (size_t)(c * NQ1 + i) * nc + cell;
        double v = U.cu * u[c][i];
        if (U.cb != 0.0) v = fma(U.cb, qb[idx], v);
        if (U.cq != 0.0) v = fma(U.cq, qa[idx], v);
        u[c][i] = v;
      }
    if (U.cp != 0.0) {
      HDG_UNROLL
      for (int a = 0; a < NP; ++a) phi[a] = fma(U.cp, pa[(size_t)a * nc + cell], phi[a]);
    }
    HDG_UNROLL
    for (int c = 0; c < 2; ++c)
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) qa[(size_t)(c * NQ1 + i) * nc + cell] = u[c][i];
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) pa[(size_t)a * nc + cell] = phi[a];
  }
#ifdef __CUDA_ARCH__
  acc = block_reduce128(acc);
#endif
  if (threadIdx.x == 0) U.partial[blockIdx.x] = acc;
}

