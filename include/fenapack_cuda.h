/* libfenapack_cuda -- C ABI of the B200-native PCD preconditioner / FGMRES hot path.
 *
 * Drop-in boundary for the ONE data-parallel path of FENaPack (blechta/fenapack):
 * applying the block-triangular PCD preconditioner inside right-preconditioned
 * (F)GMRES for the P2/P1 Oseen / Navier-Stokes saddle-point system.  Each entry
 * point below names the reference interface (file:line under /root/reference)
 * whose work it takes over.  Plain pointers and sizes only; no PyTorch, no PETSc
 * types.  All floating point is IEEE double, all matrices are CSR with 32-bit
 * indices ("PetscInt" of a default PETSc build), rows sorted or unsorted.
 *
 * Numbering.  The library works in the reference's *split* numbering: velocity
 * ("u") dofs and pressure ("p") dofs as produced by the index sets is_u / is_p
 * (fenapack/field_split.py:71-82).  In a multi-rank job every rank owns a
 * contiguous range of each split numbering (exactly PETSc's MPIAIJ row
 * partition, fenapack/SubfieldBC.h:138-140), passes its *local rows* with
 * *global column ids*, and the library builds the halo plan.
 *
 * Call order:  fnp_create[_dist] -> fnp_set_option* -> fnp_set_layout ->
 *   fnp_set_pattern (once per operator) -> fnp_set_values -> fnp_set_bc ->
 *   fnp_setup -> { fnp_pc_apply | fnp_schur_apply | fnp_solve }*  ->
 *   [fnp_set_values (A00/P00/KP) -> fnp_setup]  per Newton step -> ... -> fnp_destroy
 *
 * Error convention (replaces PETSc error codes + Python exceptions,
 * fenapack/__init__.py:31): every function returns FNP_OK (0) or a negative
 * code; fnp_last_error() returns a thread-local message.  There is no CPU
 * fallback: without a usable CUDA device fnp_create fails with FNP_ERR_CUDA.
 *
 * Threading: calls on one context must be serialised by the caller;
 * in a multi-rank job every call that touches vectors is collective.
 */
#ifndef FENAPACK_CUDA_H
#define FENAPACK_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FNP_OK 0
#define FNP_ERR_ARG (-1)      /* bad argument / call order (RuntimeError/ValueError in the reference) */
#define FNP_ERR_CUDA (-2)     /* CUDA runtime failure or no device */
#define FNP_ERR_OPTION (-3)   /* unknown option name or unparsable value */
#define FNP_ERR_NCCL (-4)
#define FNP_ERR_STATE (-5)    /* e.g. pattern changed on a value refresh, setup missing */
#define FNP_ERR_NUMERIC (-6)  /* breakdown (zero diagonal, NaN) */

/* Operators owned on the device (SURVEY 8b "Data handed across"). */
enum fnp_operator {
  FNP_MAT_A00 = 0, /* velocity block of the system matrix          PCFIELDSPLIT A00, field_split.py:82-93 */
  FNP_MAT_A01 = 1, /* u-rows x p-cols block (discrete gradient)    PCFIELDSPLIT A01 */
  FNP_MAT_A10 = 2, /* p-rows x u-cols block (discrete divergence)  PCFIELDSPLIT A10 */
  FNP_MAT_AP = 3,  /* pressure Laplacian with PCD Dirichlet rows   field_split_backend.py:67-70 */
  FNP_MAT_MP = 4,  /* pressure mass matrix                         field_split_backend.py:73-76 */
  FNP_MAT_KP = 5,  /* pressure convection (+reaction) matrix       field_split_backend.py:79-83 */
  FNP_MAT_P00 = 6, /* optional preconditioning velocity block (stabilised a_pc,
                      nonlinear_solvers.py:75-76); defaults to A00 */
  FNP_MAT_P01 = 7, /* optional u-rows x p-cols block of the PRECONDITIONING matrix: PCFIELDSPLIT cuts the
                      blocks of its triangular apply from Pmat (useAmat = false), the Krylov MatMult uses
                      Amat; set it only when the two differ.  Defaults to A01 */
  FNP_MAT_A11 = 8, /* optional p-rows x p-cols block of the system matrix (zero for Taylor-Hood; non-zero
                      for pressure-stabilised discretisations).  Enters the Krylov MatMult only */
  FNP_MAT_COUNT = 9,
  FNP_MAT_RP = 100 /* derived (PCDR): Rp = Bt^T diag(Mu)^-1 Bt, built by fnp_setup; valid for
                      fnp_spmv and the AMG introspection calls only */
};

typedef struct fnp_context fnp_context;

/* ---- life cycle ------------------------------------------------------- */

/* One context on CUDA device `device` of this process (single-GPU job).
 * Replaces PCDKSP.__init__ (field_split.py:46-57): GMRES, right PC, fieldsplit
 * SCHUR/UPPER/USER are fixed properties of the context. */
int fnp_create(fnp_context **out, int device);

/* Multi-rank job: one process (or host thread) per GPU.  `nccl_id` is the
 * 128-byte ncclUniqueId made by fnp_nccl_unique_id on rank 0 and broadcast by
 * the host communicator (mpi4py / torch.distributed) -- the role of the MPI
 * communicator argument `comm` of PCDKSP (field_split.py:46,75-77). */
int fnp_nccl_unique_id(void *out128);
int fnp_create_dist(fnp_context **out, int device, const void *nccl_id, int rank, int nranks);

int fnp_destroy(fnp_context *ctx);
const char *fnp_last_error(void);
const char *fnp_version(void);

/* Launch every kernel of this context on the caller's stream (cudaStream_t),
 * so that a host framework can order its own work and CUDA events with it. */
int fnp_set_stream(fnp_context *ctx, void *cuda_stream);
int fnp_synchronize(fnp_context *ctx);

/* ---- options ---------------------------------------------------------- */

/* PETSc options database names, without the user prefix, exactly as the
 * reference sets them (demo_navier-stokes-pcd.py:146-165):
 *   ksp_type gmres|fgmres, ksp_gmres_restart, ksp_rtol, ksp_atol, ksp_max_it
 *   fieldsplit_p_pc_python_type fenapack.PCDPC_BRM1|fenapack.PCDPC_BRM2|fenapack.PCDRPC_BRM1|fenapack.PCDRPC_BRM2
 *   fieldsplit_u_ksp_type richardson, fieldsplit_u_ksp_max_it, fieldsplit_u_pc_type amg|jacobi
 *   fieldsplit_p_PCD_Ap_ksp_type richardson|cg, ..._ksp_max_it, ..._ksp_rtol, ..._pc_type amg|jacobi
 *   fieldsplit_p_PCD_Rp_* like ..._Ap_* (PCDR only)
 *   fieldsplit_p_PCD_Mp_ksp_type chebyshev, ..._ksp_max_it, ..._ksp_chebyshev_eigenvalues "lo, hi",
 *   ..._pc_type jacobi
 *   <prefix>pc_amg_threshold, pc_amg_levels, pc_amg_coarse_size, pc_amg_smooth_steps,
 *   pc_amg_eig_ratio, pc_amg_prolongator_truncation, pc_amg_coarse_drop, pc_amg_replicate_size,
 *   pc_amg_lag (velocity block: rebuild the coarse levels at every lag-th refresh only),
 *   pc_amg_refresh galerkin|rebuild (velocity block; galerkin, the default on single-rank contexts: on a
 *   value refresh keep the prolongators and recompute the coarse operators on the device; rebuild: redo
 *   the host set-up) for the prefixes fieldsplit_u_ and fieldsplit_p_PCD_Ap_
 * "hypre"/"boomeramg"/"gamg" are accepted as aliases of amg (the smoothed-
 * aggregation hierarchy of this library).  Unknown names -> FNP_ERR_OPTION.
 * Library tuning knobs (prefix fnp_, not PETSc names): fnp_timers, fnp_cuda_graph, fnp_spmv_kernel
 * auto|csr|sell, fnp_sell_max_mean_row, fnp_sell_sigma (sorting window, before fnp_set_pattern),
 * fnp_sell_gather (bit mask: 1 16-byte gathers, 2 six CTAs/SM, 4 L2 bulk prefetch, 8 16-byte epilogue
 * loads, 16 L2 bulk prefetch in the CSR kernel, 64 default instead of evict-first cache policy for
 * operators of at most 64 MB; default 79, results are bit-identical for every value),
 * fnp_sell_warps (warps per SELL slice, 0 = from rows x mean row), fnp_sell_warps_rows, fnp_gmres_sync,
 * fnp_refresh_chunk_terms,
 * fnp_kronecker, fnp_prune_zeros, fnp_halo_overlap, fnp_halo_p2p. */
int fnp_set_option(fnp_context *ctx, const char *name, const char *value);

/* ---- operators -------------------------------------------------------- */

/* Ownership ranges of this rank in the two split numberings (single rank:
 * begin = 0, local = global).  Replaces the index-set set-up of
 * PCDKSP.init_pcd (field_split.py:71-82). */
int fnp_set_layout(fnp_context *ctx, int64_t n_u_local, int64_t u_begin, int64_t n_u_global,
                   int64_t n_p_local, int64_t p_begin, int64_t n_p_global);

/* Sparsity pattern of one operator: local rows, global column ids.  Copied.
 * Takes over Mat.createSubMatrix(is_row, is_col) (field_split_backend.py:331-334)
 * / the fieldsplit block extraction of PCSetUp_FieldSplit (field_split.py:90). */
int fnp_set_pattern(fnp_context *ctx, int which, const int32_t *rowptr, const int32_t *colidx);

/* Values for the pattern set before, length nnz, in the caller's entry order.  `values` may be a
 * host pointer (pageable or pinned; copied to the device as it is) or a device pointer (used in
 * place) -- the library asks the CUDA runtime which.  The value-only refresh of the reference's
 * MAT_REUSE_MATRIX path (field_split_backend.py:82-83, 285-291): same pattern, new numbers, once
 * per Newton step for A00/P00/KP.  All per-entry work (scatter into the stored format, Jacobi
 * diagonal, Kronecker and pruning checks) runs on the device. */
int fnp_set_values(fnp_context *ctx, int which, const double *values);

/* PCD Dirichlet dofs in local pressure numbering and their values:
 * SubfieldBC (SubfieldBC.h:48-53,92-160). */
int fnp_set_bc(fnp_context *ctx, const int32_t *idx_local, const double *values, int32_t n);

/* PCDR variants (PCDRPC_BRM1/2, preconditioners.py:173-298): diagonal of the velocity mass
 * matrix Mu (Mat.getDiagonal of setup_mat_Mu, field_split_backend.py:100-105,147-148), local
 * u numbering.  With it and A01 (= Bt, setup_mat_Bt :108-118) fnp_setup builds
 * Rp = Bt^T diag(Mu)^-1 Bt (PCDInterface._build_approx_Ap :142-166) and its solver
 * (options fieldsplit_p_PCD_Rp_*).  Single rank so far. */
int fnp_set_mu_diag(fnp_context *ctx, const double *diag_local);

/* Optional: positions of the local split dofs in the local monolithic vector
 * (dofmap_dofs_is, _field_split_utils.py:39-50).  Enables fnp_solve_monolithic. */
int fnp_set_index_sets(fnp_context *ctx, const int64_t *is_u_local, const int64_t *is_p_local);

/* (Re)build everything derived from values: Jacobi diagonals, SpMV format
 * choice, AMG hierarchies (first call: full set-up; later calls: numeric
 * refresh of the operators whose values changed).  The work of
 * BasePCDPC.setUp (preconditioners.py:71-85) and of ksp.setUp() for the inner
 * solvers (field_split_backend.py:250-255, field_split.py:103-106). */
int fnp_setup(fnp_context *ctx);

/* ---- the hot path ------------------------------------------------------ */
/* `on_device` = 0: host pointers, copies inside the call; 1: device pointers. */

/* y = A x for one stored operator (Mat.mult, preconditioners.py:131,164). */
int fnp_spmv(fnp_context *ctx, int which, const double *x, double *y, int on_device);

/* Inner solves with the configured KSP/PC:
 *   fnp_mp_solve  ksp_Mp.solve  (preconditioners.py:133,162)
 *   fnp_ap_solve  ksp_Ap.solve  (preconditioners.py:130,166)
 *   fnp_u_solve   fieldsplit "u" sub-KSP (field_split.py:93-106) */
int fnp_mp_solve(fnp_context *ctx, const double *b, double *x, int on_device);
int fnp_ap_solve(fnp_context *ctx, const double *b, double *x, int on_device);
int fnp_u_solve(fnp_context *ctx, const double *b, double *x, int on_device);
int fnp_rp_solve(fnp_context *ctx, const double *b, double *x, int on_device);   /* ksp_Rp.solve, preconditioners.py:259,293 */

/* y_p = -S^-1 x_p : PCDPC_BRM1.apply / PCDPC_BRM2.apply
 * (preconditioners.py:98-135, 148-169) including apply_pcd_bcs
 * (field_split_backend.py:62-64 -> SubfieldBC.h:162-182).  x_p is not modified. */
int fnp_schur_apply(fnp_context *ctx, const double *x_p, double *y_p, int on_device);

/* Block-triangular apply of PCFIELDSPLIT SCHUR/UPPER (field_split.py:54-57):
 * y_p = schur(x_p); y_u = A00^-1 (x_u - A01 y_p). */
int fnp_pc_apply(fnp_context *ctx, const double *x_u, const double *x_p, double *y_u, double *y_p,
                 int on_device);

/* Right-preconditioned restarted (F)GMRES on [A00 A01; A10 0] with the PCD
 * preconditioner, zero initial guess: KSPSolve of PCDKSP (field_split.py:36-57),
 * the call DOLFIN's NewtonSolver makes per Newton step (nonlinear_solvers.py:53-60).
 * Outputs: iterations, final residual-norm estimate, number of PC applies. */
int fnp_solve(fnp_context *ctx, const double *b_u, const double *b_p, double *x_u, double *x_p,
              int on_device, int32_t *iterations, double *residual_norm, int32_t *pc_applies);

/* Same with monolithic local vectors (needs fnp_set_index_sets). */
int fnp_solve_monolithic(fnp_context *ctx, const double *b, double *x, int on_device,
                         int32_t *iterations, double *residual_norm, int32_t *pc_applies);

/* KSPConvergedReason of the last fnp_solve[_monolithic], with PETSc's codes (KSPConvergedDefault on
 * the recurrence estimate): 2 KSP_CONVERGED_RTOL, 3 KSP_CONVERGED_ATOL, -3 KSP_DIVERGED_ITS; 0 before
 * the first solve.  What PCDKrylovSolver.solve checks through dolfin.PETScKrylovSolver
 * (field_split.py:153-187). */
int fnp_get_converged_reason(fnp_context *ctx, int32_t *reason);

/* Residual history of the last fnp_solve (entry 0 = ||b||). Returns count copied. */
int fnp_get_residual_history(fnp_context *ctx, double *out, int32_t capacity);

/* ---- introspection (used by the parity tests; not on the hot path) ----- */

/* AMG hierarchy of `which` (FNP_MAT_AP or FNP_MAT_A00): number of levels, then
 * per level the CSR of A_l / P_l / R_l copied to caller buffers sized from
 * fnp_amg_level_info.  kind: 0 = A, 1 = P (level l+1 -> l), 2 = R. */
/* Kronecker block size of a stored operator: bs > 1 means the library recognised
 * A = S (x) I_bs (interleaved components; the Picard/Oseen velocity block) and stores /
 * coarsens the scalar operator S only; the AMG introspection below then returns the
 * levels of S.  Option fnp_kronecker 0 (before fnp_set_pattern) disables the detection. */
int fnp_operator_block_size(fnp_context *ctx, int which, int32_t *bs);
/* The derived PCDR operator Rp = Bt^T diag(Mu)^-1 Bt as the library assembled it (PCDInterface.
 * _build_approx_Ap, field_split_backend.py:142-166): this rank's rows, GLOBAL column ids.  Sizes from
 * fnp_rp_info.  Lets a checker build its own hierarchy from bit-identical values on several ranks. */
int fnp_rp_info(fnp_context *ctx, int64_t *nrows_local, int64_t *nnz);
int fnp_rp_get(fnp_context *ctx, int32_t *rowptr, int32_t *colidx_global, double *values);
int fnp_amg_num_levels(fnp_context *ctx, int which, int32_t *levels);
int fnp_amg_level_info(fnp_context *ctx, int which, int level, int kind, int64_t *nrows,
                       int64_t *ncols, int64_t *nnz, double *rho);
int fnp_amg_level_get(fnp_context *ctx, int which, int level, int kind, int32_t *rowptr,
                      int32_t *colidx, double *values);
int fnp_amg_coarse_inverse(fnp_context *ctx, int which, double *dense_row_major);
/* One V-cycle, zero initial guess. */
int fnp_amg_vcycle(fnp_context *ctx, int which, const double *b, double *x, int on_device);

/* Stage timers (names follow the reference's dolfin Timer names, e.g.
 * "FENaPack: PCDPC_BRM1 apply", preconditioners.py:98).  Accumulated CUDA-event
 * milliseconds and call counts since the last reset; enabled by
 * fnp_set_option(ctx, "fnp_timers", "1"). */
int fnp_get_timer(fnp_context *ctx, const char *name, double *ms, int64_t *calls);
/* Algorithmic bytes accumulated by the launches of a per-kernel timer
 * (fnp_timers 2: "spmv <operator>", "multidot", "maxpy+norm", "gemv coarse"); 0 for stage timers
 * that span several kernels.  bytes / ms is the achieved bandwidth the bench reports per stage. */
int fnp_get_timer_bytes(fnp_context *ctx, const char *name, double *bytes);
/* Names of all timers recorded since the last reset, newline separated, NUL terminated; returns the
 * length needed (call with out = NULL to size the buffer). */
int fnp_timer_names(fnp_context *ctx, char *out, int64_t capacity);
int fnp_reset_timers(fnp_context *ctx);

/* Number of CUDA kernel launches issued by this context since creation. */
int64_t fnp_kernel_launches(fnp_context *ctx);

/* Timed region helpers: CUDA events on the context's stream. */
int fnp_event_tic(fnp_context *ctx);
int fnp_event_toc(fnp_context *ctx, double *elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* FENAPACK_CUDA_H */
