/*
 * hdg_b200.h -- C-ABI of the B200-native HDG solve engine.
 *
 * Drop-in boundary for the hybridisation / static-condensation path of
 * eikehmueller/IncompressibleEulerHDG (citations are into the reference tree):
 *
 *   - the statically condensed mixed-Poisson solver objects built in
 *     src/timesteppers/hdg_imex.py:123-221 ("pc_python_type": "firedrake.SCPC",
 *     "pc_sc_eliminate_fields": "0, 1", condensed_field gmres rtol 1e-12) and called through
 *     pressure_solve() at hdg_imex.py:257-272, and the same operator solved directly at
 *     src/timesteppers/hdg_implicit.py:133-146;
 *   - project_bdm() at src/timesteppers/common.py:91-108;
 *   - the tentative-velocity solves at hdg_imex.py:223-255,274-281 and hdg_implicit.py:103-129.
 *
 * Conventions
 *   * FP64 everywhere, int32 indices, all arrays C-contiguous.
 *   * Every function returns 0 on success and a non-zero HDG_E* code otherwise; the message is
 *     available from hdg_last_error().  No exceptions and no callbacks cross this boundary.
 *   * Host ("AoS") layout, used by the *_host entry points -- cell/facet major like a
 *     Firedrake Function.dat.data array in the engine's modal basis:
 *         Q[nc][2][NQ1]   p[nc][NP]   lam[nf][K+1]
 *     with NQ1 = (K+2)(K+3)/2, NP = (K+1)(K+2)/2.
 *   * Device ("SoA") layout, used by the *_dev entry points -- dof major, entity minor, so that a
 *     thread-per-entity kernel is perfectly coalesced:
 *         Q[(c*NQ1+i)*nc + cell]   p[a*nc + cell]   lam[m*nf + facet]
 *   * One handle drives one GPU (one process per GPU); multi-GPU runs create one handle per rank
 *     on that rank's local mesh and join them with hdg_comm_init() (see the multi-GPU section).
 *   * Calls are asynchronous on the engine stream unless stated; *_host entry points and
 *     functions returning scalars synchronise before returning.
 */
#ifndef HDG_B200_H
#define HDG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hdg_engine* hdg_handle;

enum {
  HDG_OK = 0,
  HDG_EINVAL = 1,   /* bad argument */
  HDG_ECUDA = 2,    /* CUDA runtime error */
  HDG_ENOGPU = 3,   /* no CUDA device: the engine has no CPU fallback */
  HDG_ESTATE = 4,   /* call order violated (e.g. apply before setup) */
  HDG_ENCCL = 5,    /* NCCL error */
  HDG_ENOCONV = 6,  /* Krylov solver hit maxit (results are still written) */
  HDG_ECOMM = 7     /* peer-memory transport: a halo exchange or all-reduce timed out (fatal for the handle: ghost data
                       and reductions after that point are invalid; sticky until hdg_p2p_enable is called again) */
};

/* ---- life cycle ------------------------------------------------------------------------------ */

/* Library version string and the pressure degrees compiled in (bit k set => degree k available). */
const char* hdg_version(void);
int hdg_supported_degrees(void);
/* Number of CUDA devices visible (0 => every other call fails with HDG_ENOGPU). */
int hdg_device_count(void);

/* Create an engine on CUDA device `device` for pressure degree k (spaces [DG_{k+1}]^2 x DG_k x
 * DGT_k, hdg_imex.py:65-69) and stabilisation tau (hdg_imex.py:58).
 *   cell_xy[nc][3][2]   vertex coordinates per cell, counter-clockwise
 *   cell_facet[nc][3]   global facet of local facet e (opposite local vertex e, running e+1 -> e+2)
 *   cell_flip[nc][3]    1 if the cell traverses the facet against its global direction
 *   facet_cell[nf][2]   adjacent cells (-1 in slot 1 on the domain boundary)
 *   facet_local[nf][2]  local facet index within each adjacent cell (-1 if absent)
 * The host arrays are copied; the engine keeps no host pointer. */
int hdg_create(int k, double tau, int nc, int nf, const double* cell_xy, const int32_t* cell_facet,
               const int32_t* cell_flip, const int32_t* facet_cell, const int32_t* facet_local,
               int device, hdg_handle* out);
int hdg_destroy(hdg_handle h);
/* Message of the last error on this handle (or of the last failed hdg_create if h == NULL). */
const char* hdg_last_error(hdg_handle h);
/* Run all engine work on an externally owned cudaStream_t (e.g. torch's current stream). */
int hdg_set_stream(hdg_handle h, void* cuda_stream);
int hdg_synchronize(hdg_handle h);

/* ---- condensed mixed-Poisson path (SCPC replacement) ----------------------------------------- */

/* K1-K3: per-cell local operators + Schur complements S_K = D - C A^-1 B (hdg_imex.py:123-133),
 * deterministic gather into the global trace matrix (mat_type aij, hdg_imex.py:135) stored as
 * blocked ELL, and the facet-block-Jacobi inverse (the ASMStarPC patches of hdg_imex.py:143-152).
 * Repeatable (re-assembles).  keep_local != 0 keeps the S_K array for hdg_get_local_schur(). */
int hdg_setup_poisson(hdg_handle h, int keep_local);

/* Copy S_K (SoA: SK[(r*NL+c)*nc + cell], NL = 3(K+1)) to a host buffer of nc*NL*NL doubles. */
int hdg_get_local_schur(hdg_handle h, double* SK_host);
/* Copy the assembled trace matrix as dense facet blocks: val[nf][5][b][b], col[nf][5] (b = K+1;
 * slot 0 is the diagonal block, slots 1-2 the other facets of cell 0, slots 3-4 of cell 1). */
int hdg_get_trace_matrix(hdg_handle h, double* val_host, int32_t* col_host);

/* a3+a4+a6: forward elimination, trace Krylov solve (CG on -S, facet-block-Jacobi, relative
 * tolerance rtol on the preconditioned residual norm, at most maxit iterations) and local
 * back-substitution for the residual (rhs_Q, rhs_p, rhs_l) in the dual space; any rhs pointer may
 * be NULL (= zero).  Writes the solution (Q, p, l); *iters receives the Krylov iteration count
 * that the reference reads at hdg_imex.py:265-271.  shift != 0 additionally applies
 * _shift_pressure (hdg_imex.py:471-478).  Host buffers, AoS layout; synchronous. */
/* Pipelined host transfers (the host-buffer route a PETSc Vec / numpy caller takes, `hdg_imex.py:257-272`, without
 * stalling the solver): the copy runs on an engine-owned copy stream and overlaps the kernels of the compute stream.
 * `slot` (0 or 1) selects one of two staging buffers per direction.  hdg_upload_begin starts host -> device and
 * returns; hdg_upload_end orders the compute stream behind it and converts AoS -> SoA into dev_soa.
 * hdg_download_begin converts SoA -> AoS on the compute stream, starts device -> host and returns; hdg_copy_wait blocks
 * until all started transfers have completed.  Host buffers should be pinned. */
int hdg_upload_begin(hdg_handle h, int kind, const double* host_aos, int slot);
int hdg_upload_end(hdg_handle h, int kind, int slot, double* dev_soa);
int hdg_download_begin(hdg_handle h, int kind, const double* dev_soa, double* host_aos, int slot);
int hdg_copy_wait(hdg_handle h);

int hdg_poisson_apply_host(hdg_handle h, const double* rhs_Q, const double* rhs_p, const double* rhs_l,
                           double* Q, double* p, double* l, double rtol, int maxit, int shift,
                           int* iters);
/* on != 0: the trace Krylov iteration of hdg_poisson_apply_dev starts from the incoming content of
 * `l` (e.g. the trace of the previous timestep; the analogue of ksp_initial_guess_nonzero) instead of
 * zero.  The tolerance then refers to the preconditioned norm of the right-hand side, so the
 * converged result does not depend on the guess.  Default off (the reference starts from zero). */
int hdg_set_initial_guess(hdg_handle h, int on);
/* warm-started trace solves that did not converge within their iteration cap and were repeated from a
 * zero guess (robustness fallback; counted for diagnostics) */
int hdg_guess_restarts(hdg_handle h, int64_t* restarts);
/* Same with device pointers in SoA layout; asynchronous except for the iteration-count read. */
int hdg_poisson_apply_dev(hdg_handle h, const double* rhs_Q, const double* rhs_p, const double* rhs_l,
                          double* Q, double* p, double* l, double rtol, int maxit, int shift,
                          int* iters);

/* The same solve with the back-substitution FUSED with the caller's update (north_star item 5: "back-substitution
 * fused with the Richardson/projection update"; replaces SCPC.apply followed by the assignments of
 * src/timesteppers/hdg_imex.py:580-587 / hdg_implicit.py:150,188-190).  With (u, phi, lam) the solution after
 * _shift_pressure (hdg_imex.py:471-478), the kernel k_back_update computes in registers, without writing u or phi:
 *     Q_acc <- cq Q_acc + cb Q_base + cu u      (Chorin: cq = 0, cb = 1, Q_base = tentative velocity, cu = dt;
 *                                                IMEX Richardson stage: cq = 1, cb = 1, cu = a_ii dt)
 *     p_acc <- cp p_acc + phi                   (Chorin: cp = 0; IMEX: cp = 1)
 *     l     <- lam
 * Q_base may be NULL when cb == 0.  Device SoA buffers as for hdg_poisson_apply_dev. */
int hdg_poisson_apply_update_dev(hdg_handle h, const double* rhs_Q, const double* rhs_p, const double* rhs_l, double cq,
                                 double* Q_acc, double cb, const double* Q_base, double cu, double cp, double* p_acc,
                                 double* l, double rtol, int maxit, int* iters);

/* ---- multigrid preconditioner for the trace solve (GTMGPC replacement, hdg_imex.py:138-169) ---- */
typedef struct {
  int32_t nrows, ncols;
  const int32_t* rowptr; /* [nrows+1] */
  const int32_t* col;    /* [nnz] */
  const double* val;     /* [nnz] */
} hdg_csr;
/* Upload a hierarchy built on the host (multigrid.py) and switch the trace solve to
 * multigrid-preconditioned CG.  A[0..nlevels): P1 operators (level 0 = P1 on the mesh, the
 * reference's get_coarse_operator hdg_imex.py:101-106 with the sign of P = -S); P[l]/R[l]:
 * prolongation level l+1 -> l and its transpose; T/Tt: P1 level 0 -> trace space (SoA numbering
 * mode*nf+facet) and its transpose; lmax[l]: estimate of lambda_max(D^-1 A[l]); coarsest_pinv:
 * dense pseudo-inverse of A[nlevels-1]; smooth_fine / smooth_coarse: Chebyshev sweeps per
 * pre-/post-smoothing on the trace level (facet-block-Jacobi, hdg_imex.py:143-152) and on the P1
 * levels (point Jacobi); cheb_ratio: the smoother targets [lmax/cheb_ratio, 1.1 lmax]. */
int hdg_mg_setup(hdg_handle h, int nlevels, const hdg_csr* A, const hdg_csr* P, const hdg_csr* R,
                 const hdg_csr* T, const hdg_csr* Tt, const double* lmax, const double* coarsest_pinv,
                 int smooth_fine, int smooth_coarse, double cheb_ratio);
/* on = 0 falls back to facet-block-Jacobi CG, on != 0 re-enables the hierarchy */
int hdg_mg_enable(hdg_handle h, int on);
int hdg_mg_info(hdg_handle h, int* nlevels, double* fine_lmax);

/* y = P x with P = -S (the symmetric positive semi-definite matrix the CG iterates on), device
 * SoA pointers -- the SpMV of the Krylov loop, exposed for tests and the roofline microbenchmark. */
int hdg_trace_spmv_dev(hdg_handle h, const double* x, double* y);
/* Per-cell pieces exposed for tests / microbenchmarks (device SoA pointers):
 *   forward elimination  r_l = rhs_l - sum_K C_K A_K^-1 (rhs_Q, rhs_p)
 *   back-substitution    (Q,p) = A_K^-1 ((rhs_Q, rhs_p) - B_K l)                               */
int hdg_forward_eliminate_dev(hdg_handle h, const double* rhs_Q, const double* rhs_p,
                              const double* rhs_l, double* r_l);
int hdg_back_substitute_dev(hdg_handle h, const double* rhs_Q, const double* rhs_p, const double* l,
                            double* Q, double* p);

/* ---- velocity side of the timesteppers (device SoA pointers) -------------------------------- */

/* penalty parameter alpha of f_impl (hdg_imex.py:56, default 1) */
int hdg_set_penalty(hdg_handle h, double alpha);
/* project_bdm (common.py:91-108): Qstar has averaged facet normal moments (zero on the boundary)
 * and the interior moments of Q; kept in the cell-wise [P_{k+1}]^2 representation. */
int hdg_project_bdm_dev(hdg_handle h, const double* Q, double* Qstar);
/* Y = c0 X + c1 M^-1 f_impl(w, X; Qstar)  (hdg_imex.py:313-331), upwind != 0 for flux="upwind".
 * Qstar must be H(div)-conforming (output of hdg_project_bdm_dev).  X and Y must not alias. */
int hdg_fimpl_apply_dev(hdg_handle h, const double* Qstar, const double* X, double c0, double c1, int upwind,
                        double* Y);
/* tentative velocity (hdg_imex.py:233-255,274-281; hdg_implicit.py:103-129), in Riesz form:
 *   (I - adt M^-1 f_impl(.;Qstar)) x = rhs     by BiCGStab, relative residual tolerance rtol.
 * zero_guess != 0 starts from x = 0, otherwise from the incoming x.  Converged when
 * ||rhs - A x||_2 <= rtol ||rhs||_2 (the PETSc convention of hdg_imex.py:224-228). */
int hdg_tentative_solve_dev(hdg_handle h, const double* Qstar, double adt, int upwind, const double* rhs, double* x,
                            double rtol, int maxit, int zero_guess, int* iters);
/* Krylov strategy of hdg_tentative_solve_dev: mode 0 = plain BiCGStab, mode 1 (default) = BiCGStab on
 * the facet-multiplier formulation of the normal-jump penalty, right-preconditioned by the
 * advection-free operator whose facet Schur complement is inverted by `sweeps` Chebyshev /
 * facet-block-Jacobi sweeps (mesh-independent iteration counts; see csrc/hdg_tent.cuh). */
int hdg_set_tentative_solver(hdg_handle h, int mode, int sweeps);
/* Counters of the tentative-velocity solver since hdg_create: out6 = {solves, BiCGStab iterations, FGMRES iterations,
 * fallbacks from BiCGStab to the restarted flexible GMRES (taken when BiCGStab stagnates, e.g. at the reference's
 * default dt = 0.04, src/driver.py:80-86, where it solves with GMRES+ILU / LU, hdg_imex.py:224-228,
 * hdg_implicit.py:126-129), failed true-residual verifications, FGMRES restart cycles}. */
int hdg_tentative_stats(hdg_handle h, int64_t* out6);
/* Counters of the mixed-precision tentative-velocity solver (FP64 iterative refinement around an FP32 BiCGStab, the
 * default; hdg_set_tuning("tent_mixed", 0) selects the all-FP64 solver): out4 = {solves, outer refinement steps,
 * inner FP32 iterations, hand-overs to the FP64 solver}. */
int hdg_mixed_stats(hdg_handle h, int64_t* out4);
/* multi-GPU only.  local_sweeps == 0 (default): the ghost facets are refreshed before every Chebyshev
 * sweep of the facet Schur preconditioner, which reproduces the single-GPU iteration exactly.
 * local_sweeps != 0: the sweeps run without halo exchanges in between (restricted overlapping Schwarz
 * on owned + ghost facets; still a fixed linear preconditioner, the converged velocity is unchanged to
 * the solver tolerance).  Measured on 2 B200: 3 instead of 10 exchanges per operator application but
 * ~+40 % BiCGStab iterations, so it only pays when exchanges dominate. */
int hdg_set_tentative_comm(hdg_handle h, int local_sweeps);
/* dual vector on the pressure space: mode 0  scale * int psi div Q dx      (hdg_implicit.py:145)
 *                                    mode 1  scale * _weak_divergence      (hdg_imex.py:353-365) */
int hdg_weak_divergence_dev(hdg_handle h, const double* Q, double scale, int mode, double* Rp);
/* Y = c0 Y + c1 M^-1 g(w, p, l)   (_pressure_gradient, hdg_imex.py:333-340) */
int hdg_pressure_gradient_dev(hdg_handle h, const double* p, const double* l, double c0, double c1, double* Y);
/* _reconstruct_trace (hdg_imex.py:450-469) and _shift_pressure (hdg_imex.py:471-478; l may be NULL) */
int hdg_reconstruct_trace_dev(hdg_handle h, const double* Q, const double* p, double* l);
int hdg_shift_pressure_dev(hdg_handle h, double* p, double* l);
/* pressure-reconstruction right-hand side (hdg_imex.py:204-207):
 *   Rp = _weak_divergence(psi, -b + (grad Q) Q),   Rl = - mu n.b ds  (zero on interior facets) */
int hdg_reconstruction_rhs_dev(hdg_handle h, const double* Q, const double* b, double* Rp, double* Rl);
/* constraint rows of the mixed operator applied to a state, Gamma(psi, mu; Q, p, l) of
 * hdg_imex.py:342-351, as dual vectors Rp [pressure space] and Rl [trace space]: the (psi, mu) part of
 * the monolithic residual of the fully implicit stage (hdg_imex.py:602-610, hdg_implicit.py:172-183) */
int hdg_gamma_apply_dev(hdg_handle h, const double* Q, const double* p, const double* l, double* Rp, double* Rl);
/* *result = sum_i x_i y_i over the owned entries of two fields of `kind` (summed over ranks; synchronous) */
int hdg_dot_dev(hdg_handle h, int kind, const double* x, const double* y, double* result);
/* *result = int_Omega x . y dx for two cell fields of the same kind (synchronous) */
int hdg_l2_inner_dev(hdg_handle h, int kind, const double* x, const double* y, double* result);
/* out = sum_t coefs[t] * ptrs[t]  over n doubles (nterms <= 8; out may alias any input) */
int hdg_lincomb_dev(hdg_handle h, int64_t n, double* out, int nterms, const double* coefs,
                    const double* const* ptrs);
/* y = M x (inverse == 0) or M^-1 x for a cell field (kind 0 velocity, 1 pressure); M = detJ I */
int hdg_mass_dev(hdg_handle h, int kind, int inverse, const double* x, double* y);

/* ---- passive tracer advection ------------------------------------------------
 * Replaces `_tracer_advection(chi, q, u, project_onto_cg=True)` (common.py:110-129) and the tracer
 * mass solves of hdg_implicit.py:93-96,192-193 and hdg_imex.py:415-448,622-623,638-639.
 *
 * hdg_tracer_setup hands over the [CG_{k+1}]^2 space of the velocity projection and the quadrature
 * tables (all host arrays, copied; NLOC = (k+2)(k+3)/2 Lagrange nodes per cell, NP = (k+1)(k+2)/2):
 *   ncg, ncg_owned          CG dofs held by this handle and how many of them it owns (equal on one GPU).  On a
 *                           partitioned mesh the dofs are numbered owned-first and their halo plan is set
 *                           beforehand with hdg_set_halo_plan(h, 18, ...) (partition.cg_plan); the multi-GPU
 *                           path has host-side (gloo) tests only -- it has not run on 2 GPUs yet
 *   cellmap [NLOC][nc]      global CG dof of local node j of every cell
 *   inc_ptr [ncg+1], inc_idx [NLOC*nc]   incidence CSR of the dofs: entries j*nc + cell in a fixed order
 *   W [NLOC][NLOC]          modal <- nodal (inverse Vandermonde matrix of the Lagrange nodes)
 *   dinv [ncg]              inverse diagonal of the CG mass matrix (Jacobi preconditioner)
 *   tab_cell [nq_cell][1+3NP+3NLOC]    per point: weight, chi, d0 chi, d1 chi, psi, d0 psi, d1 psi
 *   tab_facet [3][nq_facet][1+NP+NLOC] per local facet and point: weight, chi, psi (symmetric rule)
 * hdg_project_cg_dev: Qcg = cell-wise [P_{k+1}]^2 representation of the L2 projection of Q onto
 *   [CG_{k+1}]^2 (`Function(V_CG).project(u)`), Jacobi-PCG on the matrix-free mass matrix to
 *   <r,z> <= rtol^2 <r0,z0> per component.
 * hdg_tracer_advection_dev: out = c0 acc + c1 M^-1 [ q div(chi u) dx - (chi+ - chi-)(un+ q+ - un- q-) dS ]
 *   with u = Qcg, un = (u.n + |u.n|)/2; acc may be NULL when c0 == 0 and may alias out; q must not. */
int hdg_tracer_setup(hdg_handle h, int ncg, int ncg_owned, const int32_t* cellmap, const int32_t* inc_ptr,
                     const int32_t* inc_idx, const double* W, const double* dinv, int nq_cell, const double* tab_cell,
                     int nq_facet, const double* tab_facet);
int hdg_project_cg_dev(hdg_handle h, const double* Q, double* Qcg, double rtol, int maxit, int* iters);
int hdg_tracer_advection_dev(hdg_handle h, const double* Qcg, const double* q, double c0, const double* acc,
                             double c1, double* out);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink/NVSwitch ---------------------
 * Replaces what Firedrake/PETSc do implicitly under mpiexec (DMPlex partition, PyOP2/PetscSF halo
 * exchanges, MPI_Allreduce in every KSP; the reference itself only names COMM_WORLD at
 * hdg_imex.py:110).  Each rank creates its handle on its *local mesh* (owned cells + ghost layer,
 * owned entities numbered first; partition.py), then:
 *   hdg_comm_unique_id   rank 0 obtains a 128-byte ncclUniqueId and broadcasts it out of band
 *   hdg_comm_init        collective: joins the handle to the communicator
 *   hdg_set_partition    owned counts, global facet count and global volume (reductions run over
 *                        owned entities only and are summed over ranks)
 *   hdg_set_halo_plan    exchange plan of one entity kind: 0 cells, 1 facets, 2+l P1 level l (l < 16),
 *                        18 CG dofs of the tracer path.
 *                        Peer j sends the owned entities send_idx[send_ptr[j]..send_ptr[j+1]) and
 *                        fills the ghost block [recv_off[j], recv_off[j]+recv_cnt[j]); ghost blocks
 *                        are contiguous, ordered by peer and cover [n_owned, n_local).
 * Afterwards every entry point refreshes the ghost entries of the inputs whose neighbours it reads
 * (inputs are const except for their ghost entries) and all Krylov reductions are global. */
int hdg_comm_unique_id(void* id128);
int hdg_comm_init(hdg_handle h, int rank, int nranks, const void* id128);
int hdg_set_partition(hdg_handle h, int nc_owned, int nf_owned, int64_t nf_global, double global_volume);
int hdg_set_halo_plan(hdg_handle h, int kind, int n_owned, int n_local, int npeers, const int32_t* peer_rank,
                      const int32_t* send_ptr, const int32_t* send_idx, const int32_t* recv_off,
                      const int32_t* recv_cnt);
/* Peer-memory transport (NVLink P2P through CUDA IPC, csrc/hdg_comm.cuh): halo exchanges become a push
 * kernel that stores straight into the peers' mailboxes plus a wait/unpack kernel, and all-reduces one
 * single-CTA kernel; NCCL is then only used for the multigrid all-gather.
 *   hdg_p2p_alloc   allocates this rank's mailbox (slab_doubles per sender and parity; must hold the
 *                   largest halo block times its dofs) and returns its 64-byte cudaIpcMemHandle
 *   hdg_p2p_attach  handles[nranks][64]: maps every peer's mailbox and switches the transport on
 *   hdg_p2p_enable  0 = back to NCCL send/recv + all-reduce, 1 = peer memory (collective: same on all ranks)
 *   hdg_p2p_status  *error != 0 if a bounded flag wait timed out (a peer died or fell out of step) */
int hdg_p2p_alloc(hdg_handle h, int64_t slab_doubles, void* handle64);
int hdg_p2p_attach(hdg_handle h, const void* handles);
int hdg_p2p_enable(hdg_handle h, int on);
int hdg_p2p_status(hdg_handle h, int* error);
/* refresh the ghost entries of an SoA device field [ndof][n_local] (asynchronous on the engine stream) */
int hdg_halo_exchange_dev(hdg_handle h, int kind, int ndof, double* field);
/* in-place sum over ranks of n <= 16 device doubles */
int hdg_allreduce_sum_dev(hdg_handle h, double* values, int n);
/* Measurement aid: device time (microseconds) of one halo exchange of an [ndof][n_local] field of plan `kind` and of one
 * all-reduce of `nred` partial-sum slots, each averaged over `nrep` back-to-back calls on the engine stream.  Collective:
 * every rank calls it with the same arguments.  Zeros on a single rank. */
int hdg_comm_probe(hdg_handle h, int kind, int ndof, int nred, int nrep, double* us_exchange, double* us_allreduce);
/* multigrid levels l < repl_level are row-distributed (their CSR blocks passed to hdg_mg_setup hold
 * the owned rows with local column numbering and need halo plan 2+l); levels >= repl_level are
 * replicated.  gather_counts[q] / gather_gid: rows of level repl_level owned by rank q and their
 * global ids (rank-major), used to all-gather the restricted residual at the interface. */
int hdg_mg_set_distribution(hdg_handle h, int repl_level, const int32_t* gather_counts, const int32_t* gather_gid);
int hdg_comm_stats(hdg_handle h, int* rank, int* nranks, int64_t* exchanges, int64_t* allreduces);

/* ---- layout conversion / buffer helpers -------------------------------------------------------- */

/* kind: 0 = velocity (2*NQ1 per cell), 1 = pressure (NP per cell), 2 = trace (K+1 per facet). */
int hdg_field_size(hdg_handle h, int kind, int64_t* n);
int hdg_upload(hdg_handle h, int kind, const double* host_aos, double* dev_soa);
int hdg_download(hdg_handle h, int kind, const double* dev_soa, double* host_aos);

/* ---- timers (CUDA-event time accumulated per reference PerformanceLog label) ------------------ */
/* labels: 0 setup_poisson, 1 forward_elimination, 2 trace_solve, 3 back_substitution,
 *         4 bdm_projection, 5 tentative_velocity_solve, 6 h2d, 7 d2h,
 *         8 spmv_sampled (every 16th trace SpMV of the CG, per-launch events),
 *         9 fimpl_sampled (first f_impl application of every BiCGStab iteration),
 *         10 condense (k_condense inside setup_poisson), 11 assemble (k_assemble inside setup_poisson) */
int hdg_get_timers(hdg_handle h, double* ms, int64_t* ncalls, int n);
int hdg_reset_timers(hdg_handle h);
/* Measured FP64 FMA throughput of the device in TFLOP/s (denominator of "% of FP64 peak"). */
int hdg_measure_fp64_peak(hdg_handle h, double* tflops);
/* mean duration (ms) of nrep back-to-back launches of one Chebyshev / facet-block-Jacobi sweep on the facet Schur
 * complement of the tentative-velocity preconditioner (k_tent_sweep, the kernel with the largest share of a Chorin
 * step) for a = adt; measurement aid of bench.py, no reference counterpart */
int hdg_tent_sweep_probe(hdg_handle h, double adt, int nrep, double* ms_per_launch);
/* The iteration bodies of the three Krylov loops (BiCGStab chunk of the tentative solve, one
 * multigrid-PCG iteration, Jacobi-CG chunk) are captured once into CUDA graphs and replayed, which
 * removes the per-kernel launch cost that dominates at small per-GPU sizes (strong scaling).  On by
 * default on one GPU and with the peer-memory transport; hdg_set_graphs(h, 0) launches kernel by kernel. */
/* diagnostics: the Krylov scalars of the last trace CG (out[0..4]: reference <b,M^-1 b>, <r,z>, rtol^2,
 * iterations, done) and of the last BiCGStab (out[5..9]: ||b||^2, ||r||^2, rtol^2, iterations, done) */
int hdg_debug_scalars(hdg_handle h, double* out10);
int hdg_set_graphs(hdg_handle h, int on);
/* Engine knobs by name (no reference counterpart; also settable at start-up with HDG_TUNING="name=value,...").  They
 * select between implementations of the same operation and never change what is computed beyond the solver tolerance:
 *   "sweep_minblocks" 5 | 6 | 8   register-allocation variant of k_tent_sweep (96 / 80 / 64 registers at k = 2; default 6)
 *   "tent_cellblock", "tent_scaledx", "tent_flex", "tent_fp32", "tent_verify"   0 | 1, parts of the default
 *                                  tentative-velocity preconditioner / BiCGStab (all default 1)
 *   "tent_mixed" 0 | 1             FP64 refinement around an FP32 BiCGStab (default 0), with "tent_inner_tol", "tent_inner_cap"
 *   "tent_krylov" 0 | 1 | 2        BiCGStab with FGMRES fallback | BiCGStab only | FGMRES only; "tent_gmres_m", "tent_bicg_cap"
 *   "fimpl_split" 0 | 1 | 3        operator of the tentative iteration: k_fimpl | k_fimpl_c (default) | k_fimpl_t; "fimpl_pre" 0 | 1
 *   "condense_rows" 0 | 1          k >= 3: row-loop condensation kernel (default 0)
 *   "poisson_lsmem" -1 | 0..7      k >= 3: bit mask of the per-cell Poisson kernels that keep the Cholesky factor in shared
 *                                  memory (1 condensation, 2 forward elimination, 4 back-substitution); -1 = measured default
 *   "tracer_tables" 0 | 1          tracer advection from runtime | compile-time tables (default 1)
 *   "p2p_fused" 0 | 1              halo exchange as one kernel (default 1)
 *   "ktime" 0 | 1                  event pair around every launch (hdg_kernel_times), CUDA graphs off
 *   "tent_trace" 0 | 1             residual norms of the tentative solves on stderr
 * Unknown names return HDG_EINVAL. */
int hdg_set_tuning(hdg_handle h, const char* name, int value);
int hdg_graph_replays(hdg_handle h, int64_t* replays);
/* Number of engine kernels launched since creation (bench.py "gpu_launches"). */
int64_t hdg_launch_count(hdg_handle h);
/* The same per kernel: "name=launches\n" lines (names without template arguments) written into buf (at most len
 * bytes, NUL-terminated); returns the number of bytes the complete text needs. */
int64_t hdg_kernel_counts(hdg_handle h, char* buf, int64_t len);
/* In-situ device time per kernel (diagnostics): after hdg_set_tuning(h, "ktime", 1) every launch is bracketed by an
 * event pair and the CUDA graphs are off; "name=launches:milliseconds\n" lines since the last read (reading with a
 * buffer synchronises and clears the record); returns the number of bytes the complete text needs. */
int64_t hdg_kernel_times(hdg_handle h, char* buf, int64_t len);

#ifdef __cplusplus
}
#endif
#endif /* HDG_B200_H */
