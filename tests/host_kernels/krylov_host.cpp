// TEST INFRASTRUCTURE -- not part of the product (see cuda_shim.h).  The vector / reduction kernels of the trace CG and
// of BiCGStab (csrc/hdg_krylov.cuh, plus k_cg_update_plain of hdg_mg.cuh) as plain C functions, run with one block of
// one thread: every two-stage reduction then has a single partial (G = 1).  The Krylov scalars live in this library
// the way they live in device memory in the engine; tests/test_krylov_host.py replays the launch sequences of
// run_pcg_mg and bicgstab_loop (csrc/hdg_engine.cu).
#include "cuda_shim.h"
#include "hdg_krylov.cuh"
#include "hdg_mg.cuh"

static CgScalars g_cg;
static BiScalars g_bi;
static const OwnMask ALL = {1, 0ull, 1, 1, 1, 1};

#define BY_B(b_, ...)               \
  switch (b_) {                     \
    case 2: { constexpr int b = 2; __VA_ARGS__; } return 0; \
    case 3: { constexpr int b = 3; __VA_ARGS__; } return 0; \
    case 4: { constexpr int b = 4; __VA_ARGS__; } return 0; \
    case 5: { constexpr int b = 5; __VA_ARGS__; } return 0; \
    default: return 1;              \
  }

extern "C" {

// ---- trace CG -------------------------------------------------------------------------------------------------
int kh_trace_rhs(int k, int nc, int nf, const double* gK, const double* Rl, const int* facet_cell, const int* facet_local,
                 double* b, double* partial) {
  switch (k) {
    case 1: k_trace_rhs<1>(gK, Rl, facet_cell, facet_local, nc, nf, nf, b, partial); return 0;
    case 2: k_trace_rhs<2>(gK, Rl, facet_cell, facet_local, nc, nf, nf, b, partial); return 0;
    case 3: k_trace_rhs<3>(gK, Rl, facet_cell, facet_local, nc, nf, nf, b, partial); return 0;
    case 4: k_trace_rhs<4>(gK, Rl, facet_cell, facet_local, nc, nf, nf, b, partial); return 0;
  }
  return 1;
}
int kh_cg_init(int bs, int nf, const double* dinv, const double* part_mean, double* r, double* x, double* z, double* p,
               double* part_rz) {
  BY_B(bs, k_cg_init<b>(nf, nf, 1.0 / nf, dinv, part_mean, r, x, z, p, part_rz))
}
int kh_cg_start(const double* part_rz, double rtol, int maxit) {
  k_cg_start(&g_cg, part_rz, nullptr, 1, rtol, maxit);
  return 0;
}
int kh_cg_spmv(int bs, int nf, const double* val, const int* col, const double* p, double* q, double* part_pq,
               int with_scalars) {
  BY_B(bs, k_cg_spmv<b>(nf, nf, val, col, p, q, part_pq, with_scalars ? &g_cg : nullptr))
}
int kh_cg_update(int bs, int nf, const double* dinv, const double* p, const double* q, double* x, double* r, double* z,
                 const double* part_pq, double* part_rz) {
  BY_B(bs, k_cg_update<b>(nf, nf, dinv, p, q, x, r, z, part_pq, part_rz, &g_cg))
}
int kh_mode0_partial(int nf, const double* z, double* part) {
  k_mode0_partial(nf, z, part);
  return 0;
}
int kh_sub_mode0(int nf, const double* part, double* x) {
  k_sub_mode0(nf, part, 1.0 / nf, x);
  return 0;
}
int kh_cg_update_plain(int bs, int nf, const double* p, const double* q, double* x, double* r, const double* part_pq,
                       const double* part_q0) {
  k_cg_update_plain((size_t)bs * nf, p, q, x, r, part_pq, &g_cg, part_q0, 1.0 / nf, (size_t)nf);
  return 0;
}
int kh_cg_pupdate(int bs, int nf, const double* z, double* p, const double* part_rz, const double* part_z0) {
  BY_B(bs, k_cg_pupdate<b>(nf, z, p, part_rz, &g_cg, part_z0, 1.0 / nf))
}
int kh_cg_state(double* rz0, double* rz, int* iters, int* done) {
  *rz0 = g_cg.rz0; *rz = g_cg.rz; *iters = g_cg.iters; *done = g_cg.done;
  return 0;
}

// ---- _shift_pressure, layout conversion ---------------------------------------------------------------------------
int kh_shift_pressure(int nc, int nf, const double* xy, double volume, double* p, double* lam, double* partial) {
  k_pmean_partial(xy, nc, nc, p, partial);
  k_shift(nc, nf, 1.0 / volume, partial, p, lam);
  return 0;
}
int kh_aos_to_soa(const double* aos, double* soa, int n, int ndof) { k_aos_to_soa(aos, soa, n, ndof); return 0; }
int kh_soa_to_aos(const double* soa, double* aos, int n, int ndof) { k_soa_to_aos(soa, aos, n, ndof); return 0; }

// ---- BiCGStab -------------------------------------------------------------------------------------------------------
int kh_dot2(size_t n, const double* a, const double* b, const double* c, double* p_ab, double* p_cc) {
  k_dot2(n, ALL, a, b, c, p_ab, p_cc);
  return 0;
}
int kh_bi_init(size_t n, const double* b, double* r, double* rhat, double* p, double* part) {
  k_bi_init(n, ALL, b, (const double*)nullptr, r, rhat, p, part);
  return 0;
}
int kh_bi_start(const double* part, const double* part_ref, double rtol, int maxit) {
  k_bi_start(&g_bi, part, part_ref, 1, rtol, maxit);
  return 0;
}
int kh_bi_s(size_t n, const double* r, const double* v, double* sv, const double* p_rv) {
  k_bi_s(n, r, v, sv, p_rv, &g_bi);
  return 0;
}
int kh_bi_xr(size_t n, const double* p, const double* sv, const double* t, const double* rhat, double* x, double* r,
             const double* p_rv, const double* p_ts, const double* p_tt, double* p_rho, double* p_rr) {
  k_bi_xr(n, ALL, p, sv, t, rhat, x, r, p_rv, p_ts, p_tt, p_rho, p_rr, &g_bi);
  return 0;
}
int kh_bi_s_flex(size_t n, const double* r, const double* v, double* sv, const double* p_rv, size_t nx, const double* xh,
                 double* x) {
  k_bi_s_flex(n, r, v, sv, p_rv, &g_bi, nx, xh, x);
  return 0;
}
int kh_bi_xr_flex(size_t n, const double* sv, const double* t, const double* rhat, double* r, const double* p_ts,
                  const double* p_tt, double* p_rho, double* p_rr, size_t nx, const double* xh, double* x) {
  k_bi_xr_flex(n, ALL, sv, t, rhat, r, p_ts, p_tt, p_rho, p_rr, &g_bi, nx, xh, x);
  return 0;
}
int kh_bi_p(size_t n, const double* r, const double* v, double* p, const double* p_rv, const double* p_ts,
            const double* p_tt, const double* p_rho, const double* p_rr) {
  k_bi_p(n, r, v, p, p_rv, p_ts, p_tt, p_rho, p_rr, &g_bi);
  return 0;
}
// continue a converged run with the tolerance tightened by `factor` (k_bi_resume, the acceptance loop of run_tentative_aug)
int kh_bi_resume(size_t n, const double* r, const double* v, double* p, const double* p_rv, const double* p_ts,
                 const double* p_tt, double factor) {
  k_bi_resume(n, r, v, p, p_rv, p_ts, p_tt, &g_bi, factor);
  return 0;
}
int kh_bi_state(int* iters, int* done) {
  *iters = g_bi.iters; *done = g_bi.done;
  return 0;
}

// ---- restarted flexible GMRES (run_fgmres, csrc/hdg_engine.cu) ------------------------------------------------------
int kh_gm_dots(long n, const double* w, const double* V, long ldv, int i0, int cnt, int ww_slot, double* part) {
  k_gm_dots((size_t)n, ALL, w, V, (size_t)ldv, i0, cnt, ww_slot, part);
  return 0;
}
int kh_gm_axpy(long n, double* w, const double* V, long ldv, int i0, int cnt, const double* coef, int norm_slot,
               double* part) {
  k_gm_axpy((size_t)n, ALL, w, V, (size_t)ldv, i0, cnt, coef, norm_slot, part);
  return 0;
}
int kh_gm_scale(long n, double* w, const double* red, int slot) {
  k_gm_scale((size_t)n, w, red, slot);
  return 0;
}
int kh_gm_update(long n, double* x, const double* Z, long ldz, int j0, int cnt, const double* y) {
  k_gm_update((size_t)n, x, Z, (size_t)ldz, j0, cnt, y);
  return 0;
}
int kh_resid_norm(long n, const double* b, const double* t, double* r, double* part) {
  k_resid_norm((size_t)n, ALL, b, t, r, part);
  return 0;
}
}
