// Vector / reduction kernels of the two Krylov solvers and the small field kernels around them:
//   * trace solve: CG on P = -S with the blocked-ELL SpMV (k_cg_spmv, the roofline kernel of
//     bench.py), facet-block-Jacobi, null-space handling            -- hdg_imex.py:134-137 (condensed_field KSP)
//   * tentative velocity (a9): BiCGStab vector kernels                -- hdg_imex.py:223-228, hdg_implicit.py:129
//   * trace right-hand side, _shift_pressure (a7/a8, hdg_imex.py:471-478), AoS <-> SoA conversion
// All reductions are two-stage with a fixed tree (block_reduce / reduce_partials), Krylov scalars stay in device
// memory (CgScalars / BiScalars).  The kernels depend on hdg_local.cuh only, so tests/host_kernels can also compile
// them with g++ and run them with a single block of one thread (tests/test_krylov_host.py); the engine launches them
// from hdg_engine.cu.
#pragma once
#include "hdg_local.cuh"

struct CgScalars {
  double rz0;      // initial <r,z>
  double rz;       // current <r,z>
  double tol2;     // rtol^2
  int iters;
  int done;        // 0 running, 1 converged, 2 maxit
  int maxit;
  int pad;
};

struct BiScalars {
  double rho, rr0, rr, tol2;
  int iters, done, maxit, ticket;
  double rho_prev;  // rho of the iteration before the last one (k_bi_resume rebuilds the direction p from it)
};


// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
constexpr int BLOCK = 256;

// which entries of a flat SoA vector belong to owned entities.  The vector is one or two segments
// [ndof][stride] (cells then facets for the augmented tentative system); entity = index % stride.
struct OwnMask {
  int all;                 // 1: single GPU, everything owned
  unsigned long long n1;   // length of segment 1
  int stride1, own1, stride2, own2;
};
__device__ __forceinline__ bool is_owned(const OwnMask& m, size_t i) {
  if (m.all) return true;
  if (i < m.n1) return (int)(i % (size_t)m.stride1) < m.own1;
  return (int)((i - m.n1) % (size_t)m.stride2) < m.own2;
}

// deterministic block reduction (fixed tree); result valid in thread 0
__device__ __forceinline__ double block_reduce(double v) {
  __shared__ double sm[BLOCK / 32];
  HDG_UNROLL
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sm[w] = v;
  __syncthreads();
  if (w == 0) {
    v = (lane < BLOCK / 32) ? sm[lane] : 0.0;
    HDG_UNROLL
    for (int o = 4; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  }
  return v;
}

// every block sums the same `n` partials in the same order => identical result in all blocks
__device__ __forceinline__ double reduce_partials(const double* __restrict__ part, int n) {
  __shared__ double res;
  double v = 0.0;
  for (int i = threadIdx.x; i < n; i += BLOCK) v += part[i];
  v = block_reduce(v);
  if (threadIdx.x == 0) res = v;
  __syncthreads();
  return res;
}

// trace right-hand side of  P lam = b,  b = -(R_l - sum_K gK)  (P = -S); partial sums of the
// constant-mode component for the range projection
template <int K>
__global__ void __launch_bounds__(BLOCK) k_trace_rhs(const double* __restrict__ gK, const double* __restrict__ Rl,
                                                     const int* __restrict__ facet_cell,
                                                     const int* __restrict__ facet_local, int nc, int nf,
                                                     int nf_own, double* __restrict__ b,
                                                     double* __restrict__ partial) {
  constexpr int NL1 = Dims<K>::NL1;
  double acc = 0.0;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    int c0 = facet_cell[f], c1 = facet_cell[(size_t)nf + f];
    int e0 = facet_local[f], e1 = facet_local[(size_t)nf + f];
    HDG_UNROLL
    for (int m = 0; m < NL1; ++m) {
      double v = gK[(size_t)(e0 * NL1 + m) * nc + c0];
      if (c1 >= 0) v += gK[(size_t)(e1 * NL1 + m) * nc + c1];
      if (Rl) v -= Rl[(size_t)m * nf + f];
      b[(size_t)m * nf + f] = v;
      if (m == 0 && f < nf_own) acc += v;
    }
  }
  acc = block_reduce(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// ------------------------------------------------------------------------------------------------
// K4: preconditioned CG on P = -S with facet-block-Jacobi.  Three kernels per iteration, all
// reductions two-stage and deterministic (per-block partials re-reduced by every consumer block).
// ------------------------------------------------------------------------------------------------
// init: r = b - mean0(b) on mode 0 (projection onto range(S)), x = 0, z = Dinv r, p = z, <r,z>
template <int b>
__global__ void __launch_bounds__(BLOCK) k_cg_init(int nf, int nf_own, double inv_nf_glob,
                                                   const double* __restrict__ dinv,
                                                   const double* __restrict__ part_mean, double* __restrict__ r,
                                                   double* __restrict__ x, double* __restrict__ z,
                                                   double* __restrict__ p, double* __restrict__ part_rz) {
  double mean = reduce_partials(part_mean, gridDim.x) * inv_nf_glob;
  double acc = 0.0;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    double rv[b], zv[b];
    HDG_UNROLL
    for (int m = 0; m < b; ++m) rv[m] = r[(size_t)m * nf + f];
    rv[0] -= mean;
    r[f] = rv[0];
    HDG_UNROLL
    for (int i = 0; i < b; ++i) {
      double s = 0.0;
      HDG_UNROLL
      for (int j = 0; j < b; ++j) s = fma(dinv[(size_t)(i * b + j) * nf + f], rv[j], s);
      zv[i] = s;
      if (f < nf_own) acc = fma(s, rv[i], acc);
    }
    HDG_UNROLL
    for (int m = 0; m < b; ++m) {
      x[(size_t)m * nf + f] = 0.0;
      z[(size_t)m * nf + f] = zv[m];
      p[(size_t)m * nf + f] = zv[m];
    }
  }
  acc = block_reduce(acc);
  if (threadIdx.x == 0) part_rz[blockIdx.x] = acc;
}

// part_ref (optional): partial sums of <b, M^-1 b>, the reference of the relative tolerance when the
// iteration starts from a non-zero guess (PETSc's default test: ||r|| <= rtol ||b|| in the
// preconditioned norm); without it the reference is the initial <r,z>
__global__ void k_cg_start(CgScalars* s, const double* __restrict__ part_rz, const double* __restrict__ part_ref,
                           int n, double rtol, int maxit) {
  double rz = reduce_partials(part_rz, n);
  double ref = part_ref ? reduce_partials(part_ref, n) : rz;
  if (threadIdx.x == 0) {
    s->rz0 = ref;
    s->rz = rz;
    s->tol2 = rtol * rtol;
    s->iters = 0;
    s->maxit = maxit;
    s->done = (rz <= rtol * rtol * ref || rz <= 0.0 || maxit <= 0) ? 1 : 0;
  }
}

// r -= q (q = P x0) and partial sums of its mode-0 coefficients: the residual of a guess carries
// round-off along the constant null vector of P, which has to be projected out again before the
// correction equation P d = r is handed to the CG
template <int b>
__global__ void __launch_bounds__(BLOCK) k_cg_guess_resid(int nf, int nf_own, const double* __restrict__ q,
                                                          double* __restrict__ r, double* __restrict__ part_mean) {
  double acc = 0.0;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    HDG_UNROLL
    for (int m = 0; m < b; ++m) {
      size_t idx = (size_t)m * nf + f;
      double v = r[idx] - q[idx];
      r[idx] = v;
      if (m == 0 && f < nf_own) acc += v;
    }
  }
  acc = block_reduce(acc);
  if (threadIdx.x == 0) part_mean[blockIdx.x] = acc;
}

// A: q = P p, partial <p,q>
template <int b>
__global__ void __launch_bounds__(BLOCK) k_cg_spmv(int nf, int nf_own, const double* __restrict__ val,
                                                   const int* __restrict__ col,
                                                   const double* __restrict__ p, double* __restrict__ q,
                                                   double* __restrict__ part_pq, const CgScalars* __restrict__ s) {
  if (s && s->done) return;
  double acc = 0.0;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    double y[b];
    HDG_UNROLL
    for (int i = 0; i < b; ++i) y[i] = 0.0;
    HDG_UNROLL
    for (int j = 0; j < 5; ++j) {
      int cj = col[(size_t)j * nf + f];
      double xv[b];
      HDG_UNROLL
      for (int c = 0; c < b; ++c) xv[c] = p[(size_t)c * nf + cj];
      HDG_UNROLL
      for (int r = 0; r < b; ++r)
        HDG_UNROLL
        for (int c = 0; c < b; ++c) y[r] = fma(val[(size_t)((j * b + r) * b + c) * nf + f], xv[c], y[r]);
    }
    HDG_UNROLL
    for (int i = 0; i < b; ++i) {
      q[(size_t)i * nf + f] = y[i];
      if (part_pq && f < nf_own) acc = fma(y[i], p[(size_t)i * nf + f], acc);
    }
  }
  if (part_pq) {
    acc = block_reduce(acc);
    if (threadIdx.x == 0) part_pq[blockIdx.x] = acc;
  }
}

// B: alpha = <r,z>/<p,q>; x += alpha p; r -= alpha q; z = Dinv r; partial <r,z>
template <int b>
__global__ void __launch_bounds__(BLOCK) k_cg_update(int nf, int nf_own, const double* __restrict__ dinv,
                                                     const double* __restrict__ p, const double* __restrict__ q,
                                                     double* __restrict__ x, double* __restrict__ r,
                                                     double* __restrict__ z, const double* __restrict__ part_pq,
                                                     double* __restrict__ part_rz, const CgScalars* __restrict__ s) {
  if (s->done) return;
  double pq = reduce_partials(part_pq, gridDim.x);
  double alpha = s->rz / pq;
  double acc = 0.0;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    double rv[b];
    HDG_UNROLL
    for (int m = 0; m < b; ++m) {
      size_t idx = (size_t)m * nf + f;
      x[idx] = fma(alpha, p[idx], x[idx]);
      rv[m] = fma(-alpha, q[idx], r[idx]);
      r[idx] = rv[m];
    }
    HDG_UNROLL
    for (int i = 0; i < b; ++i) {
      double v = 0.0;
      HDG_UNROLL
      for (int j = 0; j < b; ++j) v = fma(dinv[(size_t)(i * b + j) * nf + f], rv[j], v);
      z[(size_t)i * nf + f] = v;
      if (f < nf_own) acc = fma(v, rv[i], acc);
    }
  }
  acc = block_reduce(acc);
  if (threadIdx.x == 0) part_rz[blockIdx.x] = acc;
}

// partial sums of the mode-0 coefficients of a trace vector over the owned facets
__global__ void __launch_bounds__(BLOCK) k_mode0_partial(int nf_own, const double* __restrict__ z,
                                                         double* __restrict__ part) {
  double acc = 0.0;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf_own; f += gridDim.x * blockDim.x) acc += z[f];
  acc = block_reduce(acc);
  if (threadIdx.x == 0) part[blockIdx.x] = acc;
}
// x_mode0 -= mean0 (the constant null vector of P has coefficient 1 in mode 0 of every facet)
__global__ void __launch_bounds__(BLOCK) k_sub_mode0(int nf, const double* __restrict__ part, double inv_nf_glob,
                                                     double* __restrict__ x) {
  const double mean = reduce_partials(part, gridDim.x) * inv_nf_glob;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) x[f] -= mean;
}

// C: beta = <r,z>_new / <r,z>_old; p = z + beta p; convergence bookkeeping.
// Every block derives the same decision from the same partials; block 0 publishes it.
// part_z0 (optional): partial sums of the mode-0 coefficients of z.  The multigrid preconditioner does not
// keep z orthogonal to the constant null vector of P; left alone, the search directions accumulate a
// constant component that P annihilates but that swamps <p, P p> with round-off once the residual is
// small (the CG then diverges again from ~1e-10, profiles/debug_cg_trace_r1o.log).  So the constant is
// removed from z before it enters p:  p = (z - mean0(z) n) + beta p.
template <int b>
__global__ void __launch_bounds__(BLOCK) k_cg_pupdate(int nf, const double* __restrict__ z, double* __restrict__ p,
                                                      const double* __restrict__ part_rz, CgScalars* s,
                                                      const double* __restrict__ part_z0 = nullptr,
                                                      double inv_nf_glob = 0.0) {
  __shared__ int done_in;
  __shared__ double rz_old, rz0, tol2;
  __shared__ int it, maxit;
  if (threadIdx.x == 0) {
    done_in = s->done;
    rz_old = s->rz;
    rz0 = s->rz0;
    tol2 = s->tol2;
    it = s->iters;
    maxit = s->maxit;
  }
  __syncthreads();
  if (done_in) return;
  double rz_new = reduce_partials(part_rz, gridDim.x);
  bool conv = rz_new <= tol2 * rz0;
  bool stop = conv || (it + 1 >= maxit);
  if (!stop) {
    double beta = rz_new / rz_old;
    const double zmean = part_z0 ? reduce_partials(part_z0, gridDim.x) * inv_nf_glob : 0.0;
    size_t n = (size_t)b * nf;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
      p[i] = fma(beta, p[i], z[i] - (i < (size_t)nf ? zmean : 0.0));
  }
  // publish after all blocks have read the old scalars: a grid-wide ordering is not available, so
  // the *last* block to arrive writes (ticket counter in s->pad)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    int ticket = atomicAdd(&s->pad, 1);
    if (ticket == (int)gridDim.x - 1) {
      s->pad = 0;
      s->rz = rz_new;
      s->iters = it + 1;
      s->done = conv ? 1 : (stop ? 2 : 0);
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// a7/a8: _shift_pressure  (hdg_imex.py:471-478):  p -= mean(p), lam -= mean(p)
// int_K p dx = detJ * p_0 / sqrt(2)  (Dubiner mode 0 is the constant sqrt(2))
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BLOCK) k_pmean_partial(const double* __restrict__ xy, int nc, int nc_own,
                                                         const double* __restrict__ p, double* __restrict__ partial) {
  double acc = 0.0;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc_own; cell += gridDim.x * blockDim.x) {
    double x0 = xy[cell], y0 = xy[(size_t)nc + cell];
    double x1 = xy[2 * (size_t)nc + cell], y1 = xy[3 * (size_t)nc + cell];
    double x2 = xy[4 * (size_t)nc + cell], y2 = xy[5 * (size_t)nc + cell];
    double detJ = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
    acc = fma(detJ, p[cell], acc);
  }
  acc = block_reduce(acc);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
// pressure shift after the fused back-substitution (k_back_update, hdg_poisson.cuh): `partial` holds the partial sums of
// sum_K detJ phi_0 of THIS solve's phi, which was accumulated into p (p <- cp p + phi), so removing the mean of phi
// from p is the same subtraction on mode 0; lam (the trace of this solve) is shifted like in k_shift
__global__ void __launch_bounds__(BLOCK) k_shift_n(int nc, int nf, double inv_volume, const double* __restrict__ partial,
                                                   int npartial, double* __restrict__ p, double* __restrict__ lam) {
  double integral = reduce_partials(partial, npartial) * 0.70710678118654752440;
  double shift = integral * inv_volume;
  double ps = shift * 0.70710678118654752440;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += gridDim.x * blockDim.x) p[i] -= ps;
  if (lam)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += gridDim.x * blockDim.x) lam[i] -= shift;
}
__global__ void __launch_bounds__(BLOCK) k_shift(int nc, int nf, double inv_volume, const double* __restrict__ partial,
                                                 double* __restrict__ p, double* __restrict__ lam) {
  double integral = reduce_partials(partial, gridDim.x) * 0.70710678118654752440;
  double shift = integral * inv_volume;
  double ps = shift * 0.70710678118654752440;  // coefficient of the constant in mode 0
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += gridDim.x * blockDim.x) p[i] -= ps;
  if (lam)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += gridDim.x * blockDim.x) lam[i] -= shift;
}

// ------------------------------------------------------------------------------------------------
// layout conversion AoS (entity major) <-> SoA (dof major); ndof is small
// ------------------------------------------------------------------------------------------------
__global__ void k_aos_to_soa(const double* __restrict__ aos, double* __restrict__ soa, int n, int ndof) {
  size_t total = (size_t)n * ndof;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t ent = i / ndof;
    int d = (int)(i - ent * ndof);
    soa[(size_t)d * n + ent] = aos[i];
  }
}
__global__ void k_soa_to_aos(const double* __restrict__ soa, double* __restrict__ aos, int n, int ndof) {
  size_t total = (size_t)n * ndof;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t ent = i / ndof;
    int d = (int)(i - ent * ndof);
    aos[i] = soa[(size_t)d * n + ent];
  }
}
__global__ void k_int_transpose(const int* __restrict__ aos, int* __restrict__ soa, int n, int ndof) {
  size_t total = (size_t)n * ndof;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t ent = i / ndof;
    int d = (int)(i - ent * ndof);
    soa[(size_t)d * n + ent] = aos[i];
  }
}

// ------------------------------------------------------------------------------------------------
// BiCGStab for the tentative-velocity system  (I - a dt M^-1 f_impl(.;Q*)) x = b   (Riesz form of
// hdg_imex.py:233-247 / hdg_implicit.py:103-129; the reference uses GMRES+ILU resp. direct LU).
// Five kernels per iteration, reductions deterministic as in the CG.
// ------------------------------------------------------------------------------------------------
template <typename S = double>  // S: storage type of the vectors (float in the mixed-precision solver); sums in FP64
__global__ void __launch_bounds__(BLOCK) k_dot2(size_t n, OwnMask own, const S* __restrict__ a,
                                                const S* __restrict__ b, const S* __restrict__ c,
                                                double* __restrict__ p_ab, double* __restrict__ p_cc) {
  // partial <a,b> and (optionally) <c,c> over the owned entries
  double s0 = 0.0, s1 = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (!is_owned(own, i)) continue;
    s0 = fma((double)a[i], (double)b[i], s0);
    if (c) s1 = fma((double)c[i], (double)c[i], s1);
  }
  s0 = block_reduce(s0);
  if (threadIdx.x == 0) p_ab[blockIdx.x] = s0;
  if (c) {
    s1 = block_reduce(s1);
    if (threadIdx.x == 0) p_cc[blockIdx.x] = s1;
  }
}

// r = b - t (t = A x0, or r = b if t == nullptr; b may alias r); rhat = r; p = r; partial <r,r>
template <typename S = double>
__global__ void __launch_bounds__(BLOCK) k_bi_init(size_t n, OwnMask own, const S* b, const S* __restrict__ t,
                                                   S* r, S* __restrict__ rhat,
                                                   S* __restrict__ p, double* __restrict__ part) {
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    S v = t ? b[i] - t[i] : b[i];
    r[i] = v;
    rhat[i] = v;
    p[i] = v;
    if (is_owned(own, i)) s = fma((double)v, (double)v, s);
  }
  s = block_reduce(s);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}
// part: partial sums of the initial <r,r>; part_ref (optional): partial sums of the squared norm the
// tolerance refers to (||b||^2, the PETSc convention); convergence: <r,r> <= rtol^2 * reference
__global__ void k_bi_start(BiScalars* s, const double* __restrict__ part, const double* __restrict__ part_ref, int n,
                           double rtol, int maxit) {
  double rr = reduce_partials(part, n);
  double ref = part_ref ? reduce_partials(part_ref, n) : rr;
  if (threadIdx.x == 0) {
    s->rho = rr;
    s->rr0 = ref;
    s->rr = rr;
    s->tol2 = rtol * rtol;
    s->iters = 0;
    s->maxit = maxit;
    s->ticket = 0;
    s->done = (rr <= rtol * rtol * ref || maxit <= 0) ? 1 : 0;
  }
}
// s = r - alpha v, alpha = rho / <rhat, v>
__global__ void __launch_bounds__(BLOCK) k_bi_s(size_t n, const double* __restrict__ r, const double* __restrict__ v,
                                                double* __restrict__ sv, const double* __restrict__ p_rv,
                                                const BiScalars* __restrict__ s) {
  if (s->done) return;
  double alpha = s->rho / reduce_partials(p_rv, gridDim.x);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    sv[i] = fma(-alpha, v[i], r[i]);
}
// omega = <t,s>/<t,t>; x += alpha p + omega s; r = s - omega t; partials <rhat,r>, <r,r>
__global__ void __launch_bounds__(BLOCK) k_bi_xr(size_t n, OwnMask own, const double* __restrict__ p,
                                                 const double* __restrict__ sv,
                                                 const double* __restrict__ t, const double* __restrict__ rhat,
                                                 double* __restrict__ x, double* __restrict__ r,
                                                 const double* __restrict__ p_rv, const double* __restrict__ p_ts,
                                                 const double* __restrict__ p_tt, double* __restrict__ p_rho,
                                                 double* __restrict__ p_rr, const BiScalars* __restrict__ s) {
  if (s->done) return;
  double alpha = s->rho / reduce_partials(p_rv, gridDim.x);
  double tt = reduce_partials(p_tt, gridDim.x);
  double omega = tt > 0.0 ? reduce_partials(p_ts, gridDim.x) / tt : 0.0;
  double a0 = 0.0, a1 = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double si = sv[i];
    x[i] += alpha * p[i] + omega * si;
    double ri = fma(-omega, t[i], si);
    r[i] = ri;
    if (is_owned(own, i)) {
      a0 = fma(rhat[i], ri, a0);
      a1 = fma(ri, ri, a1);
    }
  }
  a0 = block_reduce(a0);
  if (threadIdx.x == 0) p_rho[blockIdx.x] = a0;
  a1 = block_reduce(a1);
  if (threadIdx.x == 0) p_rr[blockIdx.x] = a1;
}
// Flexible variants (experimental, hdg_set_tuning "tent_flex"): the solution is accumulated from the *preconditioned*
// directions instead of being recovered from the accumulated Krylov vector at the end.  xh = [Phat^-1 .]_x is what the
// operator application leaves behind for its argument (the first nx entries of the augmented vectors are the velocity
// part).  The recursive residual then stays the residual of the accumulated x even if the preconditioner is not an
// exactly linear operator (FP32 storage inside it; tests/experiments/tent_fp32_sweeps.py).
//   k_bi_s_flex :  s = r - alpha v ;  x += alpha xh(p)
//   k_bi_xr_flex:  x += omega xh(s);  r = s - omega t ;  partials <rhat,r>, <r,r>
template <typename S = double>
__global__ void __launch_bounds__(BLOCK) k_bi_s_flex(size_t n, const S* __restrict__ r,
                                                     const S* __restrict__ v, S* __restrict__ sv,
                                                     const double* __restrict__ p_rv, const BiScalars* __restrict__ s,
                                                     size_t nx, const S* __restrict__ xh, S* __restrict__ x) {
  if (s->done) return;
  double alpha = s->rho / reduce_partials(p_rv, gridDim.x);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    sv[i] = (S)fma(-alpha, (double)v[i], (double)r[i]);
    if (i < nx) x[i] = (S)fma(alpha, (double)xh[i], (double)x[i]);
  }
}
template <typename S = double>
__global__ void __launch_bounds__(BLOCK) k_bi_xr_flex(size_t n, OwnMask own, const S* __restrict__ sv,
                                                      const S* __restrict__ t, const S* __restrict__ rhat,
                                                      S* __restrict__ r, const double* __restrict__ p_ts,
                                                      const double* __restrict__ p_tt, double* __restrict__ p_rho,
                                                      double* __restrict__ p_rr, const BiScalars* __restrict__ s,
                                                      size_t nx, const S* __restrict__ xh, S* __restrict__ x) {
  if (s->done) return;
  double tt = reduce_partials(p_tt, gridDim.x);
  double omega = tt > 0.0 ? reduce_partials(p_ts, gridDim.x) / tt : 0.0;
  double a0 = 0.0, a1 = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (i < nx) x[i] = (S)fma(omega, (double)xh[i], (double)x[i]);
    const S rs = (S)fma(-omega, (double)t[i], (double)sv[i]);
    r[i] = rs;
    const double ri = rs;  // the stored (rounded) residual is the one the recurrence continues with
    if (is_owned(own, i)) {
      a0 = fma((double)rhat[i], ri, a0);
      a1 = fma(ri, ri, a1);
    }
  }
  a0 = block_reduce(a0);
  if (threadIdx.x == 0) p_rho[blockIdx.x] = a0;
  a1 = block_reduce(a1);
  if (threadIdx.x == 0) p_rr[blockIdx.x] = a1;
}
// beta = (rho_new/rho)(alpha/omega); p = r + beta (p - omega v); bookkeeping (last block publishes)
template <typename S = double>
__global__ void __launch_bounds__(BLOCK) k_bi_p(size_t n, const S* __restrict__ r, const S* __restrict__ v,
                                                S* __restrict__ p, const double* __restrict__ p_rv,
                                                const double* __restrict__ p_ts, const double* __restrict__ p_tt,
                                                const double* __restrict__ p_rho, const double* __restrict__ p_rr,
                                                BiScalars* s) {
  __shared__ int done_in, it, maxit;
  __shared__ double rho_old, rr0, tol2;
  if (threadIdx.x == 0) {
    done_in = s->done;
    it = s->iters;
    maxit = s->maxit;
    rho_old = s->rho;
    rr0 = s->rr0;
    tol2 = s->tol2;
  }
  __syncthreads();
  if (done_in) return;
  double alpha = rho_old / reduce_partials(p_rv, gridDim.x);
  double tt = reduce_partials(p_tt, gridDim.x);
  double omega = tt > 0.0 ? reduce_partials(p_ts, gridDim.x) / tt : 0.0;
  double rho_new = reduce_partials(p_rho, gridDim.x);
  double rr = reduce_partials(p_rr, gridDim.x);
  bool conv = rr <= tol2 * rr0;
  bool stop = conv || (it + 1 >= maxit) || !(omega != 0.0) || !(rho_new != 0.0);
  if (!stop) {
    double beta = (rho_new / rho_old) * (alpha / omega);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
      p[i] = (S)fma(beta, fma(-omega, (double)v[i], (double)p[i]), (double)r[i]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    int ticket = atomicAdd(&s->ticket, 1);
    if (ticket == (int)gridDim.x - 1) {
      s->ticket = 0;
      s->rho_prev = rho_old;
      s->rho = rho_new;
      s->rr = rr;
      s->iters = it + 1;
      s->done = conv ? 1 : (stop ? 2 : 0);
      __threadfence();
    }
  }
}

// Continue a BiCGStab run that stopped as converged (done == 1) with the tolerance tightened by `factor` (< 1): the
// direction update k_bi_p skipped at the converged iteration is done now -- the partial sums of that iteration are
// still in place -- and the done flag is cleared.  Used when the acceptance test of the caller (the preconditioned
// primal residual of the tentative-velocity system) asks for more than the recurrence residual delivered.
__global__ void __launch_bounds__(BLOCK) k_bi_resume(size_t n, const double* __restrict__ r, const double* __restrict__ v,
                                                     double* __restrict__ p, const double* __restrict__ p_rv,
                                                     const double* __restrict__ p_ts, const double* __restrict__ p_tt,
                                                     BiScalars* s, double factor) {
  __shared__ int done_in, it, maxit;
  __shared__ double rho_prev, rho;
  if (threadIdx.x == 0) {
    done_in = s->done;
    it = s->iters;
    maxit = s->maxit;
    rho_prev = s->rho_prev;
    rho = s->rho;
  }
  __syncthreads();
  if (done_in != 1 || it >= maxit) return;
  double alpha = rho_prev / reduce_partials(p_rv, gridDim.x);
  double tt = reduce_partials(p_tt, gridDim.x);
  double omega = tt > 0.0 ? reduce_partials(p_ts, gridDim.x) / tt : 0.0;
  const bool ok = omega != 0.0 && rho != 0.0 && rho_prev != 0.0;
  if (ok) {
    double beta = (rho / rho_prev) * (alpha / omega);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
      p[i] = fma(beta, fma(-omega, v[i], p[i]), r[i]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    int ticket = atomicAdd(&s->ticket, 1);
    if (ticket == (int)gridDim.x - 1) {
      s->ticket = 0;
      s->tol2 *= factor * factor;
      s->done = ok ? 0 : 2;
      __threadfence();
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Restarted flexible GMRES for the tentative-velocity system: the robust path, taken when BiCGStab stagnates or breaks
// down (large CFL numbers: the reference's default dt = 0.04, src/driver.py:80-86, is CFL 10 - 40 on the benchmark
// meshes; it hands the system to GMRES+ILU, hdg_imex.py:224-228, or a direct LU, hdg_implicit.py:126-129).
// Classical Gram-Schmidt in chunks of GM_CHUNK basis vectors per pass over w; all reductions two-stage with a fixed
// tree, finished per slot by k_part_finish (hdg_comm.cuh).  The small Hessenberg least-squares problem lives on the host.
// ------------------------------------------------------------------------------------------------
constexpr int GM_CHUNK = 8;

// part[(i0 + c) G + block] = sum_owned w V_{i0+c},  c < cnt <= GM_CHUNK;  ww_slot >= 0: part[ww_slot G + block] = sum w w
__global__ void __launch_bounds__(BLOCK) k_gm_dots(size_t n, OwnMask own, const double* __restrict__ w,
                                                   const double* __restrict__ V, size_t ldv, int i0, int cnt,
                                                   int ww_slot, double* __restrict__ part) {
  double acc[GM_CHUNK], ww = 0.0;
  HDG_UNROLL
  for (int c = 0; c < GM_CHUNK; ++c) acc[c] = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (!is_owned(own, i)) continue;
    const double wi = w[i];
    ww = fma(wi, wi, ww);
    HDG_UNROLL
    for (int c = 0; c < GM_CHUNK; ++c)
      if (c < cnt) acc[c] = fma(wi, V[(size_t)(i0 + c) * ldv + i], acc[c]);
  }
  HDG_UNROLL
  for (int c = 0; c < GM_CHUNK; ++c) {
    if (c < cnt) {  // cnt is uniform over the grid: every thread reaches the barriers of block_reduce
      double s = block_reduce(acc[c]);
      if (threadIdx.x == 0) part[(size_t)(i0 + c) * gridDim.x + blockIdx.x] = s;
    }
  }
  if (ww_slot >= 0) {
    ww = block_reduce(ww);
    if (threadIdx.x == 0) part[(size_t)ww_slot * gridDim.x + blockIdx.x] = ww;
  }
}

// w -= sum_{c < cnt} coef[i0 + c] V_{i0+c};  norm_slot >= 0: part[norm_slot G + block] = sum_owned w_new^2
__global__ void __launch_bounds__(BLOCK) k_gm_axpy(size_t n, OwnMask own, double* __restrict__ w,
                                                   const double* __restrict__ V, size_t ldv, int i0, int cnt,
                                                   const double* __restrict__ coef, int norm_slot,
                                                   double* __restrict__ part) {
  double h[GM_CHUNK];
  HDG_UNROLL
  for (int c = 0; c < GM_CHUNK; ++c) h[c] = c < cnt ? coef[i0 + c] : 0.0;
  double nn = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double wi = w[i];
    HDG_UNROLL
    for (int c = 0; c < GM_CHUNK; ++c)
      if (c < cnt) wi = fma(-h[c], V[(size_t)(i0 + c) * ldv + i], wi);
    w[i] = wi;
    if (is_owned(own, i)) nn = fma(wi, wi, nn);
  }
  if (norm_slot >= 0) {
    nn = block_reduce(nn);
    if (threadIdx.x == 0) part[(size_t)norm_slot * gridDim.x + blockIdx.x] = nn;
  }
}

// w *= 1 / sqrt(red[slot])   (red: finished sums, k_part_finish)
__global__ void __launch_bounds__(BLOCK) k_gm_scale(size_t n, double* __restrict__ w, const double* __restrict__ red,
                                                    int slot) {
  const double s = 1.0 / sqrt(red[slot]);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) w[i] *= s;
}

// x += sum_{c < cnt} y[j0 + c] Z_{j0+c}   (solution update from the stored preconditioned directions)
__global__ void __launch_bounds__(BLOCK) k_gm_update(size_t n, double* __restrict__ x, const double* __restrict__ Z,
                                                     size_t ldz, int j0, int cnt, const double* __restrict__ y) {
  double c_[GM_CHUNK];
  HDG_UNROLL
  for (int c = 0; c < GM_CHUNK; ++c) c_[c] = c < cnt ? y[j0 + c] : 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double xi = x[i];
    HDG_UNROLL
    for (int c = 0; c < GM_CHUNK; ++c)
      if (c < cnt) xi = fma(c_[c], Z[(size_t)(j0 + c) * ldz + i], xi);
    x[i] = xi;
  }
}

// r = b - t (r may be null: norm only); part[block] = sum_owned (b - t)^2
__global__ void __launch_bounds__(BLOCK) k_resid_norm(size_t n, OwnMask own, const double* __restrict__ b,
                                                      const double* __restrict__ t, double* __restrict__ r,
                                                      double* __restrict__ part) {
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double v = b[i] - t[i];
    if (r) r[i] = v;
    if (is_owned(own, i)) s = fma(v, v, s);
  }
  s = block_reduce(s);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

// ------------------------------------------------------------------------------------------------
// conversions of the mixed-precision tentative-velocity solver (run_tentative_mixed, hdg_engine.cu)
// ------------------------------------------------------------------------------------------------
// out = (float)(scale * in)
__global__ void __launch_bounds__(BLOCK) k_mx_to_float(size_t n, const double* __restrict__ in, double scale,
                                                       float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = (float)(scale * in[i]);
}
// x += scale * dx   (the FP32 correction of one outer step, accumulated in FP64)
__global__ void __launch_bounds__(BLOCK) k_mx_axpy(size_t n, double* __restrict__ x, double scale,
                                                   const float* __restrict__ dx) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    x[i] = fma(scale, (double)dx[i], x[i]);
}

