// Per-cell kernels of the condensed mixed-Poisson path for the higher degrees (k >= 3), with the packed Cholesky factor
// of H = T + B B^T / detJ in SHARED memory instead of registers.
//
// Why: the thread-per-cell kernels of hdg_poisson.cuh keep the factor (55 doubles at k = 3, 120 at k = 4) next to the
// local vectors in registers.  At k = 4 that is 120 + 42 + 15 + 15 doubles for k_back and 120 + 15 + hoisted W entries
// for k_condense: ptxas caps at 255 registers and spills 1.1 - 10.4 KB of stack per thread (cuobjdump -res-usage), and the
// measured rates are 0.05 (condense) and 0.24 (forward / back) of the HBM copy rate (profiles/r2/condense_bench_r2i.jsonl).
// Here the factor is a column of a [NH][BD] shared array (bank-conflict free: consecutive threads, consecutive words; no
// barrier anywhere, a thread only ever touches its own column), all indices stay compile-time, and
//   * k_condense_b solves the NL1 rows of one facet TOGETHER (one load of L(i, k) feeds NL1 FMAs) and forms the
//     Schur entries block-wise: facet block (e, e) in full and (e, e+1 mod 3) with its mirror image -- every unordered
//     pair of facets exactly once, the W entries of the partner facet formed once per block from the immediates of the
//     reference tables (structural zeros elided) and used for NL1 dot products;
//   * k_forward_s / k_back_s / k_back_update_s are the kernels of hdg_poisson.cuh with the factor behind an accessor.
// The arithmetic per entry is the one of hdg_poisson.cuh (same operation order per right-hand side), so the results
// agree to round-off; tests/test_poisson_host.py executes both on the CPU and compares, the GPU parity tests run both.
//
// Defaults for K >= 3 where they measured faster on a B200 (lsmem_mask in hdg_engine.cu; hdg_set_tuning "poisson_lsmem" is
// the bit mask 1 condensation | 2 forward | 4 back; profiles/r2/condense_bench_r2x/r2A-C.jsonl): k = 4 condensation
// 5.49 -> 2.53 ms per 10^6 cells, back 0.69 -> 0.57 ms; k = 3 condensation 0.515 -> 0.424 ms.  What limits k_condense_b<4>
// now is instruction fetch (ncu profiles/r2/ncu_r2y_condense_b4_*: 64 % "no instruction" stalls on a 192 KB straight-line
// body), see DESIGN.md 9.  K <= 2 has no instantiation: there the register versions are spill-free.
//
// Reference: the firedrake.SCPC static condensation of a_mixed_poisson, hdg_imex.py:123-135 (as hdg_poisson.cuh).
#pragma once
#include "hdg_poisson.cuh"

// threads per block: the shared column costs NH * 8 bytes per thread (k = 3: 440 B, k = 4: 960 B); static shared
// memory stays below 48 KB
template <int K>
struct LsBlock {
  static constexpr int BD = (K >= 4) ? 32 : 64;
};

// accessor of one thread's column of the [NH][BD] shared array
// The accesses are volatile on purpose: with compile-time offsets into one array the compiler otherwise forwards every
// stored entry to its later loads, i.e. keeps the whole factor in registers again (ptxas: 255 registers and 1.6 - 15 KB of
// stack at k = 4 without it).  Volatile loads stay in program order but are still pipelined by the hardware.
template <int BD>
struct LsCol {
  double* p;
  __device__ __forceinline__ volatile double& operator()(int a, int b) const {
    return const_cast<volatile double*>(p)[tri(a, b) * BD];
  }
};

// identity the compiler cannot see through: stops common-subexpression elimination from keeping the W entries of one
// facet block alive for the next one (which is what made the unrolled k_condense<4> spill)
__device__ __forceinline__ double opaque(double x) {
#ifdef __CUDA_ARCH__
  asm volatile("" : "+d"(x));
#endif
  return x;
}
// the same with an ordering edge: the new value of x only exists once `after` has been computed, so that work depending
// on x (the W entries of the next column) is not hoisted above the accumulation that produces `after`
__device__ __forceinline__ void opaque_after(double& x0, double& x1, double& x2, double& after) {
#ifdef __CUDA_ARCH__
  asm volatile("" : "+d"(x0), "+d"(x1), "+d"(x2), "+d"(after));
#endif
}

template <int K, int BD>
__device__ __forceinline__ void build_H_s(const Geo& g, double tau, const LsCol<BD>& L) {
  using T = RefTables<K>;
  constexpr int NP = Dims<K>::NP;
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
      L(a, b) = h;
    }
  }
}

// in-place Cholesky of the shared column (operation order of cholesky<N>); the diagonal holds 1 / L_aa.
// Row j is read into registers once per column (N doubles at most) and reused by the N - j - 1 rows below it.
template <int N, int BD>
__device__ __forceinline__ void cholesky_s(const LsCol<BD>& L) {
  HDG_UNROLL
  for (int j = 0; j < N; ++j) {
    double rowj[N];
    HDG_UNROLL
    for (int k = 0; k < j; ++k) rowj[k] = L(j, k);
    double d = L(j, j);
    HDG_UNROLL
    for (int k = 0; k < j; ++k) d = fma(-rowj[k], rowj[k], d);
    double inv = rsqrt(d);
    L(j, j) = inv;
    HDG_UNROLL
    for (int i = j + 1; i < N; ++i) {
      double s = L(i, j);
      HDG_UNROLL
      for (int k = 0; k < j; ++k) s = fma(-L(i, k), rowj[k], s);
      L(i, j) = s * inv;
    }
  }
}

// x_m <- (L L^T)^-1 x_m for M right-hand sides at once (operation order of chol_solve<N> per right-hand side)
template <int N, int M, int BD>
__device__ __forceinline__ void chol_solve_s(const LsCol<BD>& L, double (&x)[M][N]) {
  HDG_UNROLL
  for (int i = 0; i < N; ++i) {
    double s[M];
    HDG_UNROLL
    for (int m = 0; m < M; ++m) s[m] = x[m][i];
    HDG_UNROLL
    for (int k = 0; k < i; ++k) {
      const double l = L(i, k);
      HDG_UNROLL
      for (int m = 0; m < M; ++m) s[m] = fma(-l, x[m][k], s[m]);
    }
    const double d = L(i, i);
    HDG_UNROLL
    for (int m = 0; m < M; ++m) x[m][i] = s[m] * d;
  }
  HDG_UNROLL
  for (int i = N - 1; i >= 0; --i) {
    double s[M];
    HDG_UNROLL
    for (int m = 0; m < M; ++m) s[m] = x[m][i];
    HDG_UNROLL
    for (int k = i + 1; k < N; ++k) {
      const double l = L(k, i);
      HDG_UNROLL
      for (int m = 0; m < M; ++m) s[m] = fma(-l, x[m][k], s[m]);
    }
    const double d = L(i, i);
    HDG_UNROLL
    for (int m = 0; m < M; ++m) x[m][i] = s[m] * d;
  }
}

// ------------------------------------------------------------------------------------------------
// K1+K2 with the factor in shared memory, facet blocks of NL1 rows (see the file header)
//   S_K = -tau G - E E^T / detJ + W H^-1 W^T,   W = E B^T / detJ + tau F
// ------------------------------------------------------------------------------------------------
template <int K, int E, int E2, bool DIAG>
__device__ __forceinline__ void condense_block(const Geo& g, const double (&nu)[3][2], double tau, const int (&fl)[3],
                                               const double (&v)[Dims<K>::NL1][Dims<K>::NP], int nc, int cell,
                                               double* __restrict__ SK) {
  using T = RefTables<K>;
  constexpr int NP = Dims<K>::NP, NL1 = Dims<K>::NL1, NL = Dims<K>::NL;
  const double nn = (g.n[E][0] * g.n[E2][0] + g.n[E][1] * g.n[E2][1]) * g.le[E] * g.le[E2] * g.idetJ;
  double nub[3][2];
  HDG_UNROLL
  for (int e = 0; e < 3; ++e) nub[e][0] = nub[e][1] = 0.0;
  nub[E2][0] = opaque(nu[E2][0]);
  nub[E2][1] = opaque(nu[E2][1]);
  double taub = opaque(tau);
  HDG_UNROLL
  for (int m2 = 0; m2 < NL1; ++m2) {
    // diagonal block: rows m <= m2 only, mirrored on store (bitwise symmetric S_K, like the K <= 3 register kernel)
    constexpr int MEND_ALL = NL1;
    const int mend = DIAG ? m2 + 1 : MEND_ALL;
    double s[NL1];
    HDG_UNROLL
    for (int m = 0; m < NL1; ++m) s[m] = 0.0;
    HDG_UNROLL
    for (int a = 0; a < NP; ++a)
      if (W_nonzero<K>(E2, m2, a)) {
        const double w = W_entry<K>(g, nub, taub, E2, m2, a);
        HDG_UNROLL
        for (int m = 0; m < NL1; ++m)
          if (m < mend) s[m] = fma(v[m][a], w, s[m]);
      }
    const double sg2 = flip_sign(fl[E2], m2);
    HDG_UNROLL
    for (int m = 0; m < NL1; ++m) {
      if (m < mend) {
        double t = s[m];
        if (T::NN(E, E2, m, m2) != 0.0) t = fma(-nn, T::NN(E, E2, m, m2), t);
        if (DIAG && m == m2) t -= tau * g.le[E];
        t *= flip_sign(fl[E], m) * sg2;
        const int r = E * NL1 + m, c = E2 * NL1 + m2;
        SK[(size_t)(r * NL + c) * nc + cell] = t;
        if (c != r) SK[(size_t)(c * NL + r) * nc + cell] = t;
      }
    }
    opaque_after(nub[E2][0], nub[E2][1], taub, s[0]);
  }
}

template <int K, int E>
__device__ __forceinline__ void condense_facet_rows(const Geo& g, const double (&nu)[3][2], double tau,
                                                    const int (&fl)[3], const LsCol<LsBlock<K>::BD>& L, int nc,
                                                    int cell, double* __restrict__ SK) {
  constexpr int NP = Dims<K>::NP, NL1 = Dims<K>::NL1;
  double v[NL1][NP];
  HDG_UNROLL
  for (int m = 0; m < NL1; ++m)
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) v[m][a] = W_entry<K>(g, nu, tau, E, m, a);
  chol_solve_s<NP, NL1, LsBlock<K>::BD>(L, v);
  condense_block<K, E, E, true>(g, nu, tau, fl, v, nc, cell, SK);
  condense_block<K, E, (E + 1) % 3, false>(g, nu, tau, fl, v, nc, cell, SK);
}

template <int K>
__global__ void __launch_bounds__(LsBlock<K>::BD) k_condense_b(const double* __restrict__ xy,
                                                               const int* __restrict__ flip, int nc, double tau,
                                                               double* __restrict__ SK) {
  using D = Dims<K>;
  constexpr int BD = LsBlock<K>::BD;
  __shared__ double Lsh[D::NH * BD];
  const LsCol<BD> L{Lsh + threadIdx.x};
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    build_H_s<K, BD>(g, tau, L);
    cholesky_s<D::NP, BD>(L);
    double nu[3][2];
    int fl[3];
    HDG_UNROLL
    for (int e = 0; e < 3; ++e) {
      fl[e] = flip[(size_t)e * nc + cell];
      nu[e][0] = g.Ji[0][0] * g.n[e][0] + g.Ji[0][1] * g.n[e][1];
      nu[e][1] = g.Ji[1][0] * g.n[e][0] + g.Ji[1][1] * g.n[e][1];
    }
    condense_facet_rows<K, 0>(g, nu, tau, fl, L, nc, cell, SK);
    condense_facet_rows<K, 1>(g, nu, tau, fl, L, nc, cell, SK);
    condense_facet_rows<K, 2>(g, nu, tau, fl, L, nc, cell, SK);
  }
}

// ------------------------------------------------------------------------------------------------
// local_solve of hdg_local.cuh with the factor behind the accessor
// ------------------------------------------------------------------------------------------------
template <int K, bool HAS_LAM>
__device__ __forceinline__ void local_solve_s(const Geo& g, double tau, const LsCol<LsBlock<K>::BD>& L,
                                              const double (&lam)[3][Dims<K>::NL1], double (&u)[2][Dims<K>::NQ1],
                                              double (&phi)[1][Dims<K>::NP]) {
  constexpr int NQ1 = Dims<K>::NQ1;
  if (HAS_LAM) {
    apply_Et<K>(g, lam, -1.0, u);
    apply_Ft<K>(g, lam, tau, phi[0]);
  }
  apply_B_over_detJ<K>(g, u, -1.0, phi[0]);
  chol_solve_s<Dims<K>::NP, 1, LsBlock<K>::BD>(L, phi);
  HDG_UNROLL
  for (int c = 0; c < 2; ++c) {
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i) u[c][i] *= g.idetJ;
  }
  apply_Bt_over_detJ<K>(g, phi[0], 1.0, u);
}

// a3: forward elimination (k_forward of hdg_poisson.cuh)
template <int K>
__global__ void __launch_bounds__(LsBlock<K>::BD) k_forward_s(const double* __restrict__ xy,
                                                              const int* __restrict__ flip, int nc, double tau,
                                                              const double* __restrict__ Ru,
                                                              const double* __restrict__ Rp, double* __restrict__ gK) {
  using D = Dims<K>;
  constexpr int NQ1 = D::NQ1, NP = D::NP, NL1 = D::NL1, BD = LsBlock<K>::BD;
  __shared__ double Lsh[D::NH * BD];
  const LsCol<BD> L{Lsh + threadIdx.x};
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    build_H_s<K, BD>(g, tau, L);
    cholesky_s<NP, BD>(L);
    double u[2][NQ1], phi[1][NP], lam[3][NL1];
    HDG_UNROLL
    for (int c = 0; c < 2; ++c)
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) u[c][i] = Ru ? Ru[(size_t)(c * NQ1 + i) * nc + cell] : 0.0;
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) phi[0][a] = Rp ? Rp[(size_t)a * nc + cell] : 0.0;
    local_solve_s<K, false>(g, tau, L, lam, u, phi);
    HDG_UNROLL
    for (int e = 0; e < 3; ++e)
      HDG_UNROLL
      for (int m = 0; m < NL1; ++m) lam[e][m] = 0.0;
    apply_E<K>(g, u, 1.0, lam);
    apply_F<K>(g, phi[0], tau, lam);
    HDG_UNROLL
    for (int e = 0; e < 3; ++e) {
      int fl = flip[(size_t)e * nc + cell];
      HDG_UNROLL
      for (int m = 0; m < NL1; ++m) gK[(size_t)(e * NL1 + m) * nc + cell] = flip_sign(fl, m) * lam[e][m];
    }
  }
}

// shared by k_back_s and k_back_update_s: loads, factorisation and local solve of one cell
template <int K>
__device__ __forceinline__ void back_cell_s(const double* __restrict__ xy, const int* __restrict__ flip,
                                            const int* __restrict__ cell_facet, int nc, int nf, double tau,
                                            const double* __restrict__ Ru, const double* __restrict__ Rp,
                                            const double* __restrict__ lamg, int cell,
                                            const LsCol<LsBlock<K>::BD>& L, Geo& g, double (&u)[2][Dims<K>::NQ1],
                                            double (&phi)[1][Dims<K>::NP]) {
  using D = Dims<K>;
  constexpr int NQ1 = D::NQ1, NP = D::NP, NL1 = D::NL1, BD = LsBlock<K>::BD;
  g = make_geo(xy, nc, cell);
  build_H_s<K, BD>(g, tau, L);
  cholesky_s<NP, BD>(L);
  double lam[3][NL1];
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
  for (int a = 0; a < NP; ++a) phi[0][a] = Rp ? Rp[(size_t)a * nc + cell] : 0.0;
  local_solve_s<K, true>(g, tau, L, lam, u, phi);
}

// K5 / a6: back-substitution (k_back of hdg_poisson.cuh)
template <int K>
__global__ void __launch_bounds__(LsBlock<K>::BD) k_back_s(const double* __restrict__ xy, const int* __restrict__ flip,
                                                           const int* __restrict__ cell_facet, int nc, int nf,
                                                           double tau, const double* __restrict__ Ru,
                                                           const double* __restrict__ Rp,
                                                           const double* __restrict__ lamg, double* __restrict__ uo,
                                                           double* __restrict__ po) {
  using D = Dims<K>;
  constexpr int NQ1 = D::NQ1, NP = D::NP, BD = LsBlock<K>::BD;
  __shared__ double Lsh[D::NH * BD];
  const LsCol<BD> L{Lsh + threadIdx.x};
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g;
    double u[2][NQ1], phi[1][NP];
    back_cell_s<K>(xy, flip, cell_facet, nc, nf, tau, Ru, Rp, lamg, cell, L, g, u, phi);
    HDG_UNROLL
    for (int c = 0; c < 2; ++c)
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) uo[(size_t)(c * NQ1 + i) * nc + cell] = u[c][i];
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) po[(size_t)a * nc + cell] = phi[0][a];
  }
}

// fixed-tree block sum for BD = 32 or 64 threads; result valid in thread 0
template <int BD>
__device__ __forceinline__ double block_reduce_s(double v) {
  static_assert(BD == 32 || BD == 64, "block_reduce_s: one or two warps");
  HDG_UNROLL
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (BD == 64) {
    __shared__ double sm2[2];
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm2[threadIdx.x >> 5] = v;
    __syncthreads();
    v = sm2[0] + sm2[1];
  }
  return v;
}

// K5 fused with the caller's update (k_back_update of hdg_poisson.cuh; partial has gridDim.x entries)
template <int K>
__global__ void __launch_bounds__(LsBlock<K>::BD) k_back_update_s(const double* __restrict__ xy,
                                                                  const int* __restrict__ flip,
                                                                  const int* __restrict__ cell_facet, int nc, int nf,
                                                                  double tau, const double* __restrict__ Ru,
                                                                  const double* __restrict__ Rp,
                                                                  const double* __restrict__ lamg, BackUpdate U) {
  using D = Dims<K>;
  constexpr int NQ1 = D::NQ1, NP = D::NP, BD = LsBlock<K>::BD;
  __shared__ double Lsh[D::NH * BD];
  const LsCol<BD> L{Lsh + threadIdx.x};
  double acc = 0.0;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g;
    double u[2][NQ1], phi[1][NP];
    back_cell_s<K>(xy, flip, cell_facet, nc, nf, tau, Ru, Rp, lamg, cell, L, g, u, phi);
    if (cell < U.nc_own) acc = fma(g.detJ, phi[0][0], acc);
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
      for (int a = 0; a < NP; ++a) phi[0][a] = fma(U.cp, pa[(size_t)a * nc + cell], phi[0][a]);
    }
    HDG_UNROLL
    for (int c = 0; c < 2; ++c)
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) qa[(size_t)(c * NQ1 + i) * nc + cell] = u[c][i];
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) pa[(size_t)a * nc + cell] = phi[0][a];
  }
#ifdef __CUDA_ARCH__
  acc = block_reduce_s<BD>(acc);
#endif
  if (threadIdx.x == 0) U.partial[blockIdx.x] = acc;
}
