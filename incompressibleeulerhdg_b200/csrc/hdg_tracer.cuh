// Passive tracer advection:
//
//   _tracer_advection(chi, q, u, project_onto_cg=True)                       src/timesteppers/common.py:110-129
//       u_ = project(u, [CG_{k+1}]^2)                                         (global L2 projection)
//       un = (u_.n + |u_.n|) / 2
//       q div(chi u_) dx - (chi^+ - chi^-) (un^+ q^+ - un^- q^-) dS           (interior facets only)
//   mass solves  chi sigma dx == chi q dx + dt a_ij * advection               hdg_implicit.py:93-96,192-193;
//                                                                             hdg_imex.py:415-448,622-623,638-639
//
// Device design.  The tracer space is DG_k in the orthonormal Dubiner basis (mass matrix detJ I), so
// the mass solve is a division by detJ that is folded into the advection kernel.  The only global
// solve is the CG_{k+1} mass matrix of the velocity projection; it is applied matrix-free: with the
// nodal (Lagrange) basis phi_j = sum_i W[i][j] psi_i of the orthonormal modal basis psi,
//     M_K = detJ W^T W,       load_K = detJ W^T U_K,       U^cg_K = W x_K
// (W = inverse Vandermonde matrix of the equispaced Lagrange nodes, the compile-time table VINV), i.e.
// two small dense products per cell and a
// deterministic gather over the cells that share a dof (incidence CSR, fixed order, no atomics).
// Jacobi-PCG with both velocity components advanced together (one alpha/beta per component); all
// Krylov scalars stay in device memory, the host polls the residual every few iterations.  On a partitioned
// mesh the dofs are numbered owned-first (partition.cg_plan): dots run over the owned prefix and are summed
// over the ranks, ghost entries of p and x are refreshed before the kernels that gather through the cell map.
//
// The advection kernel evaluates the volume and interior-facet integrals by quadrature (the non-polynomial
// |u.n| rules out closed-form reference tensors), one thread per cell, the neighbour's tracer trace with
// the facet parameter reversed (both cells are counter-clockwise).  Two variants: k_tracer_adv_t reads the
// compile-time tables of hdg_tables.inc (default facet rule; table entries are immediate operands),
// k_tracer_adv reads runtime tables with warp-uniform (broadcast) loads for any other facet rule.
#pragma once
#include "hdg_local.cuh"

struct TracerScalars {
  double rz[2][2];   // <r,z> per parity and component
  double pAp[2];
  double rz0[2];
};

struct TracerState {
  int ncg = 0, ncg_own = 0, nloc = 0, nq_cell = 0, nq_facet = 0;  // ncg_own < ncg on a partitioned mesh
  int *cellmap = nullptr;   // [nloc][nc]
  int *inc_ptr = nullptr;   // [ncg+1]
  int *inc_idx = nullptr;   // [nloc*nc]  entries j*nc + cell
  double *W = nullptr;      // [nloc][nloc]  modal <- nodal
  double *dinv = nullptr;   // [ncg] inverse diagonal of the CG mass matrix
  double *tab_cell = nullptr, *tab_facet = nullptr;
  double *yK = nullptr;     // [2][nloc][nc]
  double *x = nullptr, *r = nullptr, *z = nullptr, *p = nullptr, *Ap = nullptr;  // [2][ncg]
  double *part = nullptr;   // [2][grid]
  TracerScalars* scal = nullptr;
  TracerScalars* scal_host = nullptr;  // pinned
};

// ---- CG projection ------------------------------------------------------------------------------
// yK[c][j][cell] = detJ sum_i W[i][j] U[c][i][cell]            (load vector, element contributions)
// W = RefTables<K>::VINV (compile-time: the entries are immediate operands of the DFMAs; the runtime
// copy handed to hdg_tracer_setup is checked against it on the host)
template <int K>
__global__ void __launch_bounds__(128) k_cgp_load(const double* __restrict__ xy, int nc,
                                                  const double* __restrict__ U, double* __restrict__ yK) {
  using T = RefTables<K>;
  constexpr int NLOC = Dims<K>::NQ1;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    HDG_UNROLL
    for (int c = 0; c < 2; ++c) {
      double u[NLOC];
      HDG_UNROLL
      for (int i = 0; i < NLOC; ++i) u[i] = U[(size_t)(c * NLOC + i) * nc + cell];
      HDG_UNROLL
      for (int j = 0; j < NLOC; ++j) {
        double s = 0.0;
        HDG_UNROLL
        for (int i = 0; i < NLOC; ++i)
          if (T::VINV(i, j) != 0.0) s = fma(T::VINV(i, j), u[i], s);
        yK[(size_t)(c * NLOC + j) * nc + cell] = g.detJ * s;
      }
    }
  }
}

// out[0] = max |W - VINV|, out[1] = max |VINV|  (set-up check, one thread)
template <int K>
__global__ void k_cgp_check_w(const double* __restrict__ W, double* __restrict__ out) {
  using T = RefTables<K>;
  constexpr int NLOC = Dims<K>::NQ1;
  double err = 0.0, mx = 0.0;
  for (int i = 0; i < NLOC; ++i)
    for (int j = 0; j < NLOC; ++j) {
      err = fmax(err, fabs(W[i * NLOC + j] - T::VINV(i, j)));
      mx = fmax(mx, fabs(T::VINV(i, j)));
    }
  out[0] = err;
  out[1] = mx;
}

// yK = M_K x_K = detJ W^T (W x_K)  with x_K gathered through the cell -> dof map
template <int K>
__global__ void __launch_bounds__(128) k_cgp_cellop(const double* __restrict__ xy, int nc, int ncg,
                                                    const int* __restrict__ cellmap,
                                                    const double* __restrict__ x, double* __restrict__ yK) {
  using T = RefTables<K>;
  constexpr int NLOC = Dims<K>::NQ1;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    int dof[NLOC];
    HDG_UNROLL
    for (int j = 0; j < NLOC; ++j) dof[j] = cellmap[(size_t)j * nc + cell];
    HDG_UNROLL
    for (int c = 0; c < 2; ++c) {
      double xl[NLOC], m[NLOC];
      HDG_UNROLL
      for (int j = 0; j < NLOC; ++j) xl[j] = x[(size_t)c * ncg + dof[j]];
      HDG_UNROLL
      for (int i = 0; i < NLOC; ++i) {
        double s = 0.0;
        HDG_UNROLL
        for (int j = 0; j < NLOC; ++j)
          if (T::VINV(i, j) != 0.0) s = fma(T::VINV(i, j), xl[j], s);
        m[i] = s;
      }
      HDG_UNROLL
      for (int j = 0; j < NLOC; ++j) {
        double s = 0.0;
        HDG_UNROLL
        for (int i = 0; i < NLOC; ++i)
          if (T::VINV(i, j) != 0.0) s = fma(T::VINV(i, j), m[i], s);
        yK[(size_t)(c * NLOC + j) * nc + cell] = g.detJ * s;
      }
    }
  }
}

// cell-wise modal representation of a CG field:  Ucg[c][i][cell] = sum_j W[i][j] x[c][dof_j]
template <int K>
__global__ void __launch_bounds__(128) k_cgp_tocell(int nc, int ncg, const int* __restrict__ cellmap,
                                                    const double* __restrict__ x, double* __restrict__ Ucg) {
  using T = RefTables<K>;
  constexpr int NLOC = Dims<K>::NQ1;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    int dof[NLOC];
    HDG_UNROLL
    for (int j = 0; j < NLOC; ++j) dof[j] = cellmap[(size_t)j * nc + cell];
    HDG_UNROLL
    for (int c = 0; c < 2; ++c) {
      double xl[NLOC];
      HDG_UNROLL
      for (int j = 0; j < NLOC; ++j) xl[j] = x[(size_t)c * ncg + dof[j]];
      HDG_UNROLL
      for (int i = 0; i < NLOC; ++i) {
        double s = 0.0;
        HDG_UNROLL
        for (int j = 0; j < NLOC; ++j)
          if (T::VINV(i, j) != 0.0) s = fma(T::VINV(i, j), xl[j], s);
        Ucg[(size_t)(c * NLOC + i) * nc + cell] = s;
      }
    }
  }
}

// deterministic gather of element contributions: out[c][g] = sum_{t in inc(g)} yK[c][inc_idx[t]].
// MODE 0: first PCG step  b = out; x = 0; r = b; z = dinv r; p = z; partial <r,z>
// MODE 1: Ap = out; partial <p,Ap>
template <int MODE>
__global__ void __launch_bounds__(BLOCK) k_cgp_gather(int ncg, int ncg_own, size_t comp_stride,
                                                      const int* __restrict__ inc_ptr,
                                                      const int* __restrict__ inc_idx,
                                                      const double* __restrict__ yK, const double* __restrict__ dinv,
                                                      double* __restrict__ x, double* __restrict__ r,
                                                      double* __restrict__ z, double* __restrict__ p,
                                                      double* __restrict__ Ap, double* __restrict__ part) {
  double acc0 = 0.0, acc1 = 0.0;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < ncg; g += gridDim.x * blockDim.x) {
    int t0 = inc_ptr[g], t1 = inc_ptr[g + 1];
    double s0 = 0.0, s1 = 0.0;
    for (int t = t0; t < t1; ++t) {
      int idx = inc_idx[t];
      s0 += yK[idx];
      s1 += yK[comp_stride + idx];
    }
    size_t g1 = (size_t)ncg + g;
    if (MODE == 0) {
      double d = dinv[g];
      x[g] = 0.0;
      x[g1] = 0.0;
      r[g] = s0;
      r[g1] = s1;
      double z0 = d * s0, z1 = d * s1;
      z[g] = z0;
      z[g1] = z1;
      p[g] = z0;
      p[g1] = z1;
      if (g < ncg_own) {  // reductions run over the owned dofs only
        acc0 = fma(s0, z0, acc0);
        acc1 = fma(s1, z1, acc1);
      }
    } else {
      Ap[g] = s0;
      Ap[g1] = s1;
      if (g < ncg_own) {
        acc0 = fma(p[g], s0, acc0);
        acc1 = fma(p[g1], s1, acc1);
      }
    }
  }
  acc0 = block_reduce(acc0);
  __syncthreads();
  acc1 = block_reduce(acc1);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = acc0;
    part[gridDim.x + blockIdx.x] = acc1;
  }
}

// single block: finish the two partial sums; what = 0: rz[par] (and rz0 if init), 1: pAp
__global__ void __launch_bounds__(BLOCK) k_cgp_finish(const double* __restrict__ part, int n, TracerScalars* s,
                                                      int what, int par, int init) {
  double a = reduce_partials(part, n);
  __syncthreads();
  double b = reduce_partials(part + n, n);
  if (threadIdx.x == 0) {
    if (what == 0) {
      s->rz[par][0] = a;
      s->rz[par][1] = b;
      if (init) {
        s->rz0[0] = a;
        s->rz0[1] = b;
      }
    } else {
      s->pAp[0] = a;
      s->pAp[1] = b;
    }
  }
}

// x += alpha p; r -= alpha Ap; z = dinv r; partial <r,z>      (alpha = rz[par] / pAp per component)
__global__ void __launch_bounds__(BLOCK) k_cgp_update(int ncg, int ncg_own, const TracerScalars* __restrict__ s, int par,
                                                      const double* __restrict__ dinv, const double* __restrict__ p,
                                                      const double* __restrict__ Ap, double* __restrict__ x,
                                                      double* __restrict__ r, double* __restrict__ z,
                                                      double* __restrict__ part) {
  double al[2];
  HDG_UNROLL
  for (int c = 0; c < 2; ++c) al[c] = (s->pAp[c] > 0.0) ? s->rz[par][c] / s->pAp[c] : 0.0;
  double acc[2] = {0.0, 0.0};
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < ncg; g += gridDim.x * blockDim.x) {
    double d = dinv[g];
    HDG_UNROLL
    for (int c = 0; c < 2; ++c) {
      size_t i = (size_t)c * ncg + g;
      x[i] = fma(al[c], p[i], x[i]);
      double rr = fma(-al[c], Ap[i], r[i]);
      r[i] = rr;
      double zz = d * rr;
      z[i] = zz;
      if (g < ncg_own) acc[c] = fma(rr, zz, acc[c]);
    }
  }
  double a0 = block_reduce(acc[0]);
  __syncthreads();
  double a1 = block_reduce(acc[1]);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = a0;
    part[gridDim.x + blockIdx.x] = a1;
  }
}

// p = z + beta p,  beta = rz[new] / rz[old]
__global__ void __launch_bounds__(BLOCK) k_cgp_dir(int ncg, const TracerScalars* __restrict__ s, int par_old,
                                                   const double* __restrict__ z, double* __restrict__ p) {
  double be[2];
  HDG_UNROLL
  for (int c = 0; c < 2; ++c) be[c] = (s->rz[par_old][c] > 0.0) ? s->rz[par_old ^ 1][c] / s->rz[par_old][c] : 0.0;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < ncg; g += gridDim.x * blockDim.x) {
    HDG_UNROLL
    for (int c = 0; c < 2; ++c) {
      size_t i = (size_t)c * ncg + g;
      p[i] = fma(be[c], p[i], z[i]);
    }
  }
}

// ---- tracer advection -----------------------------------------------------------------------------
// out = c0 * acc + c1 * M^-1 adv(chi; q, u)   with u the cell-wise representation of a CG velocity.
// tab_cell  [nq_cell][1 + 3 NP + 3 NQ1]:  w, chi_a, d0 chi_a, d1 chi_a, psi_i, d0 psi_i, d1 psi_i
// tab_facet [3][nq_facet][1 + NP + NQ1]:  w, chi_a, psi_i at the point s_q of local facet e; the rule
//           must be symmetric (s_{n-1-q} = 1 - s_q) so that the neighbour's values at the same physical
//           point are the entry (e', n-1-q).
template <int K>
__global__ void __launch_bounds__(128) k_tracer_adv(const double* __restrict__ xy, const int* __restrict__ nbr,
                                                    const int* __restrict__ nbr_e, int nc, int nq_cell,
                                                    const double* __restrict__ tab_cell, int nq_facet,
                                                    const double* __restrict__ tab_facet,
                                                    const double* __restrict__ U, const double* __restrict__ q,
                                                    double c0, const double* acc, double c1, double* out) {
  constexpr int NQ1 = Dims<K>::NQ1, NP = Dims<K>::NP;
  constexpr int SC = 1 + 3 * NP + 3 * NQ1, SF = 1 + NP + NQ1;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    double u[2][NQ1], qk[NP], res[NP];
    HDG_UNROLL
    for (int c = 0; c < 2; ++c)
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) u[c][i] = U[(size_t)(c * NQ1 + i) * nc + cell];
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) {
      qk[a] = q[(size_t)a * nc + cell];
      res[a] = 0.0;
    }
    // volume term: int q div(chi u) dx / detJ = sum_qp w q (beta . grad^ chi + chi div u)
    for (int qp = 0; qp < nq_cell; ++qp) {
      const double* t = tab_cell + (size_t)qp * SC;
      double w = __ldg(t);
      double qv = 0.0;
      HDG_UNROLL
      for (int a = 0; a < NP; ++a) qv = fma(__ldg(t + 1 + a), qk[a], qv);
      double uv[2] = {0.0, 0.0}, du[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) {
        double ps = __ldg(t + 1 + 3 * NP + i), d0 = __ldg(t + 1 + 3 * NP + NQ1 + i),
               d1 = __ldg(t + 1 + 3 * NP + 2 * NQ1 + i);
        HDG_UNROLL
        for (int c = 0; c < 2; ++c) {
          uv[c] = fma(ps, u[c][i], uv[c]);
          du[c][0] = fma(d0, u[c][i], du[c][0]);
          du[c][1] = fma(d1, u[c][i], du[c][1]);
        }
      }
      double b0 = g.Ji[0][0] * uv[0] + g.Ji[0][1] * uv[1];
      double b1 = g.Ji[1][0] * uv[0] + g.Ji[1][1] * uv[1];
      double divu = g.Ji[0][0] * du[0][0] + g.Ji[1][0] * du[0][1] + g.Ji[0][1] * du[1][0] + g.Ji[1][1] * du[1][1];
      double wq = w * qv;
      HDG_UNROLL
      for (int a = 0; a < NP; ++a) {
        double v = b0 * __ldg(t + 1 + NP + a) + b1 * __ldg(t + 1 + 2 * NP + a) + divu * __ldg(t + 1 + a);
        res[a] = fma(wq, v, res[a]);
      }
    }
    // interior facets: - int_e chi (max(un,0) q_K + min(un,0) q_nbr) ds / detJ
    for (int e = 0; e < 3; ++e) {
      int nb = nbr[(size_t)e * nc + cell];
      if (nb < 0) continue;
      int ne = nbr_e[(size_t)e * nc + cell];
      double qn[NP];
      HDG_UNROLL
      for (int a = 0; a < NP; ++a) qn[a] = q[(size_t)a * nc + nb];
      double nx = e == 0 ? g.n[0][0] : (e == 1 ? g.n[1][0] : g.n[2][0]);
      double ny = e == 0 ? g.n[0][1] : (e == 1 ? g.n[1][1] : g.n[2][1]);
      double le = e == 0 ? g.le[0] : (e == 1 ? g.le[1] : g.le[2]);
      double fscale = le * g.idetJ;
      for (int qf = 0; qf < nq_facet; ++qf) {
        const double* t = tab_facet + (size_t)(e * nq_facet + qf) * SF;
        const double* tn = tab_facet + (size_t)(ne * nq_facet + (nq_facet - 1 - qf)) * SF;
        double w = __ldg(t);
        double un = 0.0;
        HDG_UNROLL
        for (int i = 0; i < NQ1; ++i) un = fma(__ldg(t + 1 + NP + i), nx * u[0][i] + ny * u[1][i], un);
        double qin = 0.0, qout = 0.0;
        HDG_UNROLL
        for (int a = 0; a < NP; ++a) {
          qin = fma(__ldg(t + 1 + a), qk[a], qin);
          qout = fma(tn[1 + a], qn[a], qout);
        }
        double flux = fmax(un, 0.0) * qin + fmin(un, 0.0) * qout;
        double wf = -w * fscale * flux;
        HDG_UNROLL
        for (int a = 0; a < NP; ++a) res[a] = fma(wf, __ldg(t + 1 + a), res[a]);
      }
    }
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) {
      size_t idx = (size_t)a * nc + cell;
      out[idx] = (c0 != 0.0 ? c0 * acc[idx] : 0.0) + c1 * res[a];
    }
  }
}

// Same operator from the compile-time tables of hdg_tables.inc (cell rule WQ/PHI/DPHI/DPSI exact to degree
// 3k+2 >= 3k, facet rule WF/PHIF/PSIF with NQF = ceil((3k+4)/2) Gauss points = the default facet rule):
// every table entry is an immediate operand, so the kernel issues no table loads at all (the runtime-table
// kernel above is LSU bound: one broadcast load per DFMA).  The tracer basis is the first NP functions of
// the hierarchical velocity basis, so chi = PHI[.][a < NP].  Used when nq_facet == NQF; the neighbour's local
// facet index is a runtime value, so its trace is evaluated for the three candidates and selected.
template <int K>
__global__ void __launch_bounds__(128) k_tracer_adv_t(const double* __restrict__ xy, const int* __restrict__ nbr,
                                                      const int* __restrict__ nbr_e, int nc,
                                                      const double* __restrict__ U, const double* __restrict__ q,
                                                      double c0, const double* acc, double c1, double* out) {
  using T = RefTables<K>;
  constexpr int NQ1 = Dims<K>::NQ1, NP = Dims<K>::NP, NQ = T::NQ, NQF = T::NQF;
  for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nc; cell += gridDim.x * blockDim.x) {
    Geo g = make_geo(xy, nc, cell);
    double u[2][NQ1], qk[NP], res[NP];
    HDG_UNROLL
    for (int c = 0; c < 2; ++c)
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) u[c][i] = U[(size_t)(c * NQ1 + i) * nc + cell];
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) {
      qk[a] = q[(size_t)a * nc + cell];
      res[a] = 0.0;
    }
    // contravariant velocity coefficients beta_d = sum_c Ji[d][c] u_c: the volume term only needs beta and
    // div u = sum_d d_d beta_d
    double be[2][NQ1];
    HDG_UNROLL
    for (int i = 0; i < NQ1; ++i) {
      be[0][i] = g.Ji[0][0] * u[0][i] + g.Ji[0][1] * u[1][i];
      be[1][i] = g.Ji[1][0] * u[0][i] + g.Ji[1][1] * u[1][i];
    }
    HDG_UNROLL
    for (int qp = 0; qp < NQ; ++qp) {
      double qv = 0.0;
      HDG_UNROLL
      for (int a = 0; a < NP; ++a)
        if (T::PHI(qp, a) != 0.0) qv = fma(T::PHI(qp, a), qk[a], qv);
      double b0 = 0.0, b1 = 0.0, divu = 0.0;
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) {
        if (T::PHI(qp, i) != 0.0) {
          b0 = fma(T::PHI(qp, i), be[0][i], b0);
          b1 = fma(T::PHI(qp, i), be[1][i], b1);
        }
        if (T::DPHI(0, qp, i) != 0.0) divu = fma(T::DPHI(0, qp, i), be[0][i], divu);
        if (T::DPHI(1, qp, i) != 0.0) divu = fma(T::DPHI(1, qp, i), be[1][i], divu);
      }
      const double wq = T::WQ(qp) * qv;
      const double w0 = wq * b0, w1 = wq * b1, wd = wq * divu;
      HDG_UNROLL
      for (int a = 0; a < NP; ++a) {
        if (T::DPSI(0, qp, a) != 0.0) res[a] = fma(w0, T::DPSI(0, qp, a), res[a]);
        if (T::DPSI(1, qp, a) != 0.0) res[a] = fma(w1, T::DPSI(1, qp, a), res[a]);
        if (T::PHI(qp, a) != 0.0) res[a] = fma(wd, T::PHI(qp, a), res[a]);
      }
    }
    HDG_UNROLL
    for (int e = 0; e < 3; ++e) {
      const int nb = nbr[(size_t)e * nc + cell];
      if (nb < 0) continue;
      const int ne = nbr_e[(size_t)e * nc + cell];
      double qn[NP], un_c[NQ1];
      HDG_UNROLL
      for (int a = 0; a < NP; ++a) qn[a] = q[(size_t)a * nc + nb];
      HDG_UNROLL
      for (int i = 0; i < NQ1; ++i) un_c[i] = g.n[e][0] * u[0][i] + g.n[e][1] * u[1][i];
      const double fscale = g.le[e] * g.idetJ;
      HDG_UNROLL
      for (int qf = 0; qf < NQF; ++qf) {
        double un = 0.0;
        HDG_UNROLL
        for (int i = 0; i < NQ1; ++i)
          if (T::PHIF(e, qf, i) != 0.0) un = fma(T::PHIF(e, qf, i), un_c[i], un);
        double qin = 0.0, o0 = 0.0, o1 = 0.0, o2 = 0.0;
        HDG_UNROLL
        for (int a = 0; a < NP; ++a) {
          if (T::PSIF(e, qf, a) != 0.0) qin = fma(T::PSIF(e, qf, a), qk[a], qin);
          if (T::PSIF(0, NQF - 1 - qf, a) != 0.0) o0 = fma(T::PSIF(0, NQF - 1 - qf, a), qn[a], o0);
          if (T::PSIF(1, NQF - 1 - qf, a) != 0.0) o1 = fma(T::PSIF(1, NQF - 1 - qf, a), qn[a], o1);
          if (T::PSIF(2, NQF - 1 - qf, a) != 0.0) o2 = fma(T::PSIF(2, NQF - 1 - qf, a), qn[a], o2);
        }
        const double qout = ne == 0 ? o0 : (ne == 1 ? o1 : o2);
        const double flux = fmax(un, 0.0) * qin + fmin(un, 0.0) * qout;
        const double wf = -T::WF(qf) * fscale * flux;
        HDG_UNROLL
        for (int a = 0; a < NP; ++a)
          if (T::PSIF(e, qf, a) != 0.0) res[a] = fma(wf, T::PSIF(e, qf, a), res[a]);
      }
    }
    HDG_UNROLL
    for (int a = 0; a < NP; ++a) {
      size_t idx = (size_t)a * nc + cell;
      out[idx] = (c0 != 0.0 ? c0 * acc[idx] : 0.0) + c1 * res[a];
    }
  }
}
