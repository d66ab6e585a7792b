// The PCD hot path: inner solves, the two Schur-complement approximations of
// fenapack/preconditioners.py (PCDPC_BRM1.apply :98-135, PCDPC_BRM2.apply
// :148-169) and the block-triangular apply of PCFIELDSPLIT SCHUR/UPPER that
// fenapack/field_split.py:54-57 selects.  The vector updates of the reference
// (copy, axpy, scale -- one petsc4py call each) are folded into the epilogues
// of the neighbouring SpMV-class kernels.
#include <algorithm>

#include "fnp_internal.cuh"

namespace fnp {

// ---------------------------------------------------------------------------
// CG with device-resident scalars (no host round trip per iteration)
// ---------------------------------------------------------------------------
__global__ void cg_xr_kernel(int64_t n, const double *__restrict__ rz, const double *__restrict__ pq,
                             const double *__restrict__ p, const double *__restrict__ q, double *__restrict__ x,
                             double *__restrict__ r) {
  const double alpha = pq[0] != 0.0 ? rz[0] / pq[0] : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += alpha * p[i];
    r[i] -= alpha * q[i];
  }
}
__global__ void cg_p_kernel(int64_t n, const double *__restrict__ rz_new, const double *__restrict__ rz_old,
                            const double *__restrict__ z, double *__restrict__ p) {
  const double beta = rz_old[0] != 0.0 ? rz_new[0] / rz_old[0] : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = z[i] + beta * p[i];
}

static void allreduce1(Ctx &c, double *dev) {
  if (c.nranks > 1) FNP_NCCL(nccl().AllReduce(dev, dev, 1, ncclDouble, ncclSum, c.comm, c.stream));
}

static void apply_inner_pc(Ctx &c, const InnerOpts &o, const DevCsr &A, DevHierarchy *H, const double *r, double *z) {
  if (o.pc == PC_AMG) {
    amg_vcycle(c, *H, r, z);
  } else if (o.pc == PC_JACOBI) {
    vec_pointwise_scale(c, A.vec_rows(), 1.0, A.dinv.p, r, nullptr, z);
  } else {
    vec_copy(c, A.vec_rows(), r, z);
  }
}

// Generic inner solve  x = KSP(A, PC)(b), zero initial guess.  work: 4 vectors.
static void inner_solve(Ctx &c, const InnerOpts &o, const DevCsr &A, DevHierarchy *H, const double *b, double *x,
                        double *w0, double *w1, double *w2, double *w3) {
  const int64_t n = A.vec_rows();
  switch (o.ksp) {
    case KSP_PREONLY:
      apply_inner_pc(c, o, A, H, b, x);
      break;
    case KSP_RICHARDSON: {
      // x1 = B b ; x_{k+1} = x_k + B (b - A x_k)      (KSPRICHARDSON, scale 1)
      apply_inner_pc(c, o, A, H, b, x);
      for (int k = 1; k < o.max_it; ++k) {
        spmv_axpby(c, A, x, -1.0, 1.0, b, w0);
        apply_inner_pc(c, o, A, H, w0, w1);
        vec_axpy(c, n, 1.0, w1, x);
      }
      break;
    }
    case KSP_CHEBYSHEV:
      FNP_REQUIRE(o.pc == PC_JACOBI, FNP_ERR_OPTION, "chebyshev is implemented with pc_type jacobi only");
      cheb_jacobi(c, A, b, o.emin, o.emax, o.max_it, 1.0, nullptr, x, w0, w1);
      break;
    case KSP_CG: {
      // preconditioned CG, fixed iteration count, scalars stay on the device
      double *r = w0, *z = w1, *p = w2, *q = w3;
      double *rz = c.red_out.p + 1000, *rz2 = c.red_out.p + 1001, *pq = c.red_out.p + 1002;   // above the Hessenberg column
      vec_zero(c, n, x);
      vec_copy(c, n, b, r);
      apply_inner_pc(c, o, A, H, r, z);
      vec_copy(c, n, z, p);
      dot(c, n, r, z, rz);
      allreduce1(c, rz);
      const int nb = 148 * 8;
      for (int k = 0; k < o.max_it; ++k) {
        spmv_store(c, A, p, q);
        dot(c, n, p, q, pq);
        allreduce1(c, pq);
        cg_xr_kernel<<<nb, 256, 0, c.stream>>>(n, rz, pq, p, q, x, r);
        c.launches++;
        if (k + 1 == o.max_it) break;
        apply_inner_pc(c, o, A, H, r, z);
        dot(c, n, r, z, rz2);
        allreduce1(c, rz2);
        cg_p_kernel<<<nb, 256, 0, c.stream>>>(n, rz2, rz, z, p);
        c.launches++;
        std::swap(rz, rz2);
      }
      FNP_CUDA(cudaPeekAtLastError());
      break;
    }
  }
}

void mp_solve(Ctx &c, const double *b, double out_scale, const double *add, double *x) {
  StageTimer t(c, "FENaPack: PCD_Mp solve");
  const DevCsr &Mp = c.dmat[FNP_MAT_MP];
  const InnerOpts &o = c.opt_mp;
  if (o.ksp == KSP_CHEBYSHEV && o.pc == PC_JACOBI) {
    cheb_jacobi(c, Mp, b, o.emin, o.emax, o.max_it, out_scale, add, x, c.p_w[3].p, c.p_w[4].p);
    return;
  }
  FNP_REQUIRE(add == nullptr || add != x, FNP_ERR_ARG, "mp_solve: aliasing not supported for this KSP");
  inner_solve(c, o, Mp, nullptr, b, x, c.p_w[1].p, c.p_w[2].p, c.p_w[3].p, c.p_w[4].p);
  if (add) vec_axpby(c, Mp.nrows, out_scale, x, 1.0, add, x);
  else if (out_scale != 1.0) vec_scale(c, Mp.nrows, out_scale, x);
}

void ap_solve(Ctx &c, const double *b, double *x) {
  StageTimer t(c, "FENaPack: PCD_Ap solve");
  inner_solve(c, c.opt_ap, c.dmat[FNP_MAT_AP], &c.amg_ap, b, x, c.p_w[1].p, c.p_w[2].p, c.p_w[3].p, c.p_w[4].p);
}

void rp_solve(Ctx &c, const double *b, double *x) {
  StageTimer t(c, "FENaPack: PCD_Rp solve");
  inner_solve(c, c.opt_rp, c.rp, &c.amg_rp, b, x, c.p_w[1].p, c.p_w[2].p, c.p_w[3].p, c.p_w[4].p);
}

void u_solve(Ctx &c, const double *b, double *x) {
  StageTimer t(c, "FENaPack: fieldsplit_u solve");
  inner_solve(c, c.opt_u, c.velocity_pc_matrix(), &c.amg_u, b, x, c.u_w[1].p, c.u_w[2].p, c.u_w[3].p, c.u_w[4].p);
}

void schur_apply(Ctx &c, const double *x, double *y) {
  const int64_t n = c.n_p;
  const DevCsr &Kp = c.dmat[FNP_MAT_KP];
  double *z = c.p_w[0].p;
  const bool pcdr = c.variant >= 3;            // PCDR: an extra -Rp^-1 x (preconditioners.py:251-262, 284-297)
  if (c.variant == 1 || c.variant == 3) {
    StageTimer t(c, pcdr ? "FENaPack: PCDRPC_BRM1 apply" : "FENaPack: PCDPC_BRM1 apply");
    // z = x ; z[bc] = g                        preconditioners.py:128-129
    vec_copy_bc(c, n, x, z, c.bc_idx.p, c.bc_val.p, c.nbc);
    // y = Ap^-1 z                              :130
    ap_solve(c, z, y);
    // z = Kp y + x                             :131-132
    spmv_axpby(c, Kp, y, 1.0, 1.0, x, z);
    // y = -Mp^-1 z                             :133-135
    mp_solve(c, z, -1.0, nullptr, y);
  } else {
    StageTimer t(c, pcdr ? "FENaPack: PCDRPC_BRM2 apply" : "FENaPack: PCDPC_BRM2 apply");
    double *z0 = c.p_w[5].p;
    // y = Mp^-1 x                              preconditioners.py:162
    mp_solve(c, x, 1.0, nullptr, y);
    // z = Kp y ; z[bc] = g                     :163-165
    spmv_store(c, Kp, y, z);
    vec_scatter_bc(c, z, c.bc_idx.p, c.bc_val.p, c.nbc);
    // z0 = Ap^-1 z                             :166
    ap_solve(c, z, z0);
    // y = -(y + z0)                            :167-169
    vec_axpby(c, n, -1.0, y, -1.0, z0, y);
  }
  if (pcdr) {
    // z = Rp^-1 x ; y = -(y_pcd + z)           :259-262 / :293-297 (the sign is already in y)
    double *zr = c.p_w[6].p;
    rp_solve(c, x, zr);
    vec_axpy(c, n, -1.0, zr, y);
  }
}

void pc_apply(Ctx &c, const double *x_u, const double *x_p, double *y_u, double *y_p) {
  FNP_REQUIRE(c.have_values[FNP_MAT_A00] && c.have_values[FNP_MAT_A01], FNP_ERR_STATE,
              "this context holds the Schur-complement operators only (no velocity block)");
  StageTimer t(c, "FENaPack: PCD fieldsplit apply");
  // y_p = S^-1 x_p
  schur_apply(c, x_p, y_p);
  // t = x_u - A01 y_p
  double *tu = c.u_w[0].p;
  {
    // the triangular apply uses the block of the preconditioning matrix (PCFIELDSPLIT, useAmat = false)
    StageTimer t2(c, "FENaPack: A01 mult");
    const DevCsr &B01 = c.have_values[FNP_MAT_P01] ? c.dmat[FNP_MAT_P01] : c.dmat[FNP_MAT_A01];
    spmv_axpby(c, B01, y_p, -1.0, 1.0, x_u, tu);
  }
  // y_u = A00^-1 t
  u_solve(c, tu, y_u);
}

void Ctx::drop_graph() {
  if (pc_graph) cudaGraphExecDestroy(pc_graph);
  pc_graph = nullptr;
  pc_graph_nodes = 0;
}

void pc_apply_vec(Ctx &c, const double *x, double *y) {
  const int64_t n = c.n_u + c.n_p;
  // multi-rank: the apply contains peer-memory exchanges (device-resident sequence numbers) or NCCL
  // send/recv and all-gathers (capturable since NCCL 2.9); fnp_cuda_graph 1 restricts graphs to one rank
  const bool graphable = c.use_graph && (c.nranks == 1 || c.use_graph >= 2) && c.timers_on == 0 &&
                         c.stream != nullptr;   // the legacy default stream cannot be captured
  if (!graphable) {
    pc_apply(c, x, x + c.n_u, y, y + c.n_u);
    return;
  }
  if (!c.pc_graph) {
    c.g_in.ensure((size_t)n);
    c.g_out.ensure((size_t)n);
    const int64_t before = c.launches;
    cudaGraph_t graph = nullptr;
    FNP_CUDA(cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeThreadLocal));
    try {
      pc_apply(c, c.g_in.p, c.g_in.p + c.n_u, c.g_out.p, c.g_out.p + c.n_u);
    } catch (...) {
      cudaStreamEndCapture(c.stream, &graph);
      if (graph) cudaGraphDestroy(graph);
      throw;
    }
    FNP_CUDA(cudaStreamEndCapture(c.stream, &graph));
    c.pc_graph_nodes = c.launches - before;
    c.launches = before;                       // nothing ran yet; replays are counted below
    FNP_CUDA(cudaGraphInstantiate(&c.pc_graph, graph, 0));
    FNP_CUDA(cudaGraphDestroy(graph));
  }
  FNP_CUDA(cudaMemcpyAsync(c.g_in.p, x, n * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
  FNP_CUDA(cudaGraphLaunch(c.pc_graph, c.stream));
  FNP_CUDA(cudaMemcpyAsync(y, c.g_out.p, n * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
  c.launches += c.pc_graph_nodes;
}

void system_matvec(Ctx &c, const double *x, double *y) {
  StageTimer t(c, "FENaPack: system MatMult");
  const double *xu = x, *xp = x + c.n_u;
  double *yu = y, *yp = y + c.n_u;
  spmv_store(c, c.dmat[FNP_MAT_A00], xu, yu);
  spmv_axpby(c, c.dmat[FNP_MAT_A01], xp, 1.0, 1.0, yu, yu);
  spmv_store(c, c.dmat[FNP_MAT_A10], xu, yp);
  if (c.have_values[FNP_MAT_A11]) spmv_axpby(c, c.dmat[FNP_MAT_A11], xp, 1.0, 1.0, yp, yp);
}

// ---------------------------------------------------------------------------
// set-up
// ---------------------------------------------------------------------------
static void build_amg(Ctx &c, int which, DevHierarchy &H, const AmgParams &p) {
  H.params = p;
  const bool is_u = which == FNP_MAT_A00 || which == FNP_MAT_P00;
  const int bs = c.kron_bs[which];
  std::vector<int64_t> begins = is_u ? c.u_begins : c.p_begins;
  for (auto &b : begins) b /= bs;              // Kronecker mode: the hierarchy is built on the scalar operator
  ensure_host_values(c, which);                // the set-up runs on the host; values live on the device
  c.drop_graph();                              // the captured apply addresses the old levels
  amg_build_host(c, c.hmat[which], begins, p, H.host);
  amg_upload(c, H, which == FNP_MAT_AP ? "Ap" : "A00", &c.dmat[which], bs);
  H.host_vals_stale = false;
}

std::vector<double> comm_allgather_padded(Ctx &c, const double *v, int64_t count, int64_t maxcount);
double comm_allreduce(Ctx &c, double v, bool max_op);
std::shared_ptr<HaloPlan> build_halo(Ctx &c, HostCsr &h, const std::vector<int64_t> &begins,
                                     std::vector<int64_t> *ghost_global_out);

// Rp = Bt^T diag(Mu)^-1 Bt, row partitioned like every pressure operator.  A rank holds the u-rows of
// Bt, so its product T_r = Bt_r^T D_r^-1 Bt_r contributes to rows of Rp owned by its neighbours (the
// transposeMatMult of the reference, field_split_backend.py:161-166, which PETSc resolves with its own
// communication): the foreign rows travel as (row, column, value) triplets in one padded all-gather
// and every rank adds what falls into its range, own contribution first, then by rank -- a fixed order.
static void build_rp(Ctx &c) {
  const HostCsr &Bt = c.hmat[FNP_MAT_A01];     // local u rows, GLOBAL p columns
  HostCsr S = Bt, B, T;                        // S = D^-1 Bt (rows scaled), B = Bt^T
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < S.nrows; ++i) {
    const double m = c.mu_diag[i];
    const double f = m != 0.0 ? 1.0 / m : 0.0;
    for (int32_t k = S.rowptr[i]; k < S.rowptr[i + 1]; ++k) S.val[k] *= f;
  }
  host_transpose(Bt, B);                       // n_p_global x n_u_local
  host_spgemm(B, S, T);                        // n_p_global x n_p_global, rows sorted by column
  const int64_t p0 = c.p_begin, p1 = c.p_begin + c.n_p;
  HostCsr &R = c.h_rp;
  if (c.nranks == 1) {
    R = std::move(T);
  } else {
    std::vector<double> mine;                  // foreign rows as triplets
    for (int64_t i = 0; i < T.nrows; ++i) {
      if (i >= p0 && i < p1) continue;
      for (int32_t k = T.rowptr[i]; k < T.rowptr[i + 1]; ++k) {
        mine.push_back((double)i);
        mine.push_back((double)T.col[k]);
        mine.push_back(T.val[k]);
      }
    }
    const int64_t cnt = (int64_t)mine.size();
    const int64_t maxcnt = std::max<int64_t>(3, (int64_t)comm_allreduce(c, (double)cnt, true));
    const double marker = -1.0;                // padding: row id -1
    mine.resize((size_t)maxcnt, marker);
    for (int64_t t = cnt; t < maxcnt; ++t) mine[(size_t)t] = marker;
    std::vector<double> all = comm_allgather_padded(c, mine.data(), maxcnt, maxcnt);
    // own rows first, then the neighbours' contributions in rank order
    std::vector<std::vector<std::pair<int32_t, double>>> rows((size_t)c.n_p);
    for (int64_t i = p0; i < p1; ++i)
      for (int32_t k = T.rowptr[i]; k < T.rowptr[i + 1]; ++k) rows[(size_t)(i - p0)].push_back({T.col[k], T.val[k]});
    for (int q = 0; q < c.nranks; ++q) {
      if (q == c.rank) continue;
      const double *tq = all.data() + (size_t)q * maxcnt;
      for (int64_t t = 0; t + 2 < maxcnt; t += 3) {
        const int64_t i = (int64_t)tq[t];
        if (i < p0 || i >= p1) continue;
        rows[(size_t)(i - p0)].push_back({(int32_t)tq[t + 1], tq[t + 2]});
      }
    }
    R = HostCsr();
    R.nrows = c.n_p;
    R.ncols = c.n_p_global;
    R.rowptr.assign((size_t)c.n_p + 1, 0);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < c.n_p; ++i) {
      auto &r = rows[(size_t)i];
      std::stable_sort(r.begin(), r.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
      size_t o = 0;
      for (size_t t = 0; t < r.size(); ++t) {
        if (o > 0 && r[o - 1].first == r[t].first) r[o - 1].second += r[t].second;
        else r[o++] = r[t];
      }
      r.resize(o);
    }
    for (int64_t i = 0; i < c.n_p; ++i) R.rowptr[(size_t)i + 1] = R.rowptr[(size_t)i] + (int32_t)rows[(size_t)i].size();
    R.col.resize((size_t)R.rowptr[(size_t)c.n_p]);
    R.val.resize((size_t)R.rowptr[(size_t)c.n_p]);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < c.n_p; ++i) {
      int32_t o = R.rowptr[(size_t)i];
      for (const auto &e : rows[(size_t)i]) {
        R.col[(size_t)o] = e.first;
        R.val[(size_t)o] = e.second;
        ++o;
      }
    }
  }
  // device copy: local column numbering [owned | ghost] and a halo plan, like the uploaded operators
  HostCsr loc = R;
  std::shared_ptr<HaloPlan> plan = build_halo(c, loc, c.p_begins, nullptr);
  const int64_t n_own = c.n_p;
  csr_upload_pattern(c, c.rp, loc, "Rp", (plan && c.split_rows(loc.nrows)) ? n_own : -1, 1, plan ? n_own : -1);
  c.rp.halo = plan;
  if (plan) {
    c.rp.ncols_own = (int32_t)n_own;
    c.rp.nghost = plan->nghost;
  }
  csr_set_values(c, c.rp, loc, R.val.data(), true);
}

void setup_all(Ctx &c) {
  StageTimer t(c, "FENaPack: setup");
  FNP_REQUIRE(c.have_layout, FNP_ERR_STATE, "fnp_set_layout must be called before fnp_setup");
  // pressure side is always needed; the velocity side only when the context owns the
  // whole block-triangular apply (n_u == 0: Schur-complement-only mode, the python-PC path)
  for (int w : {(int)FNP_MAT_AP, (int)FNP_MAT_MP, (int)FNP_MAT_KP})
    FNP_REQUIRE(c.have_values[w], FNP_ERR_STATE, "fnp_setup: operator " + std::to_string(w) + " has no values");
  const bool have_u = c.have_values[FNP_MAT_A00];   // PCDR Schur-only contexts hold A01 but no velocity block
  if (have_u)
    for (int w : {(int)FNP_MAT_A00, (int)FNP_MAT_A01, (int)FNP_MAT_A10})
      FNP_REQUIRE(c.have_values[w], FNP_ERR_STATE, "fnp_setup: operator " + std::to_string(w) + " has no values");
  FNP_REQUIRE(c.variant >= 1 && c.variant <= 4, FNP_ERR_STATE, "PCD variant not set");
  {
    // work space: sized once; the captured apply addresses it, so a reallocation drops the graph
    bool grow = c.red_out.n < 1024;
    for (auto &b : c.p_w) grow = grow || b.n < (size_t)c.n_p;
    for (auto &b : c.u_w) grow = grow || b.n < (size_t)c.n_u;
    // block partials of the reductions: room for a full restart cycle (restart + 1 basis vectors and
    // one batch of padding), so that nothing is reallocated inside the hot path or under a graph
    const size_t red_need = (size_t)c.num_sms * 4 * (size_t)(std::max(c.restart, 192) + 1 + 8);
    grow = grow || c.red_partial.n < red_need;
    if (grow) {
      c.drop_graph();
      for (auto &b : c.p_w) b.ensure((size_t)c.n_p);
      for (auto &b : c.u_w) b.ensure((size_t)c.n_u);
      c.red_out.ensure(1024);
      c.red_partial.ensure(red_need);
    }
  }
  const int uidx = c.velocity_pc_index();
  // AMG hierarchies (Ap: once; velocity block: whenever its values changed)
  if (c.opt_ap.pc == PC_AMG && (c.dirty[FNP_MAT_AP] || !c.amg_ap.built)) build_amg(c, FNP_MAT_AP, c.amg_ap, c.opt_ap.amg);
  if (have_u && c.opt_u.pc == PC_AMG && (c.dirty[uidx] || !c.amg_u.built)) {
    // value refresh of an existing hierarchy.  Level 0 aliases the context's operator and its
    // Jacobi diagonal, so it always follows the new values; the coarse levels are
    //   * recomputed on the device with frozen prolongators (pc_amg_refresh galerkin, the default), or
    //   * kept for `lag` refreshes (pc_amg_lag), or
    //   * rebuilt on the host.
    // A pattern rebuild of the block (pruning found a new non-zero) invalidates the hierarchy.
    const bool same_shape = c.amg_u.built && !c.amg_u.levels.empty() && c.amg_u.levels[0].Ap == &c.dmat[uidx] &&
                            c.amg_u_gen == c.pattern_gen[uidx] && c.amg_u_which == uidx;
    if (same_shape && c.opt_u.amg.refresh == 1 && !c.amg_u.tail && c.opt_u.amg.coarse_drop == 0.0) {
      amg_refresh_device(c, c.amg_u, c.kron_bs[uidx]);
    } else if (same_shape && c.amg_u_age + 1 < c.opt_u.amg.lag) {
      ++c.amg_u_age;
    } else {
      build_amg(c, uidx, c.amg_u, c.opt_u.amg);
      c.amg_u_age = 0;
      c.amg_u_gen = c.pattern_gen[uidx];
      c.amg_u_which = uidx;
    }
  }
  if (c.variant >= 3) {
    // PCDR: Rp = Bt^T diag(Mu)^-1 Bt with Bt = A01 (PCDInterface._build_approx_Ap,
    // field_split_backend.py:142-166), rebuilt when A01 or Mu changed
    FNP_REQUIRE(c.have_values[FNP_MAT_A01] && (int64_t)c.mu_diag.size() == c.n_u, FNP_ERR_STATE,
                "PCDR needs A01 (the discrete pressure gradient) and fnp_set_mu_diag");
    if (c.dirty[FNP_MAT_A01] || c.mu_dirty || !c.amg_rp.built) {
      ensure_host_values(c, FNP_MAT_A01);
      c.drop_graph();
      build_rp(c);
      if (c.opt_rp.pc == PC_AMG) {
        c.amg_rp.params = c.opt_rp.amg;
        amg_build_host(c, c.h_rp, c.p_begins, c.opt_rp.amg, c.amg_rp.host);
        amg_upload(c, c.amg_rp, "Rp", &c.rp, 1);
      }
      c.mu_dirty = false;
    }
  }
  for (int w = 0; w < FNP_MAT_COUNT; ++w) c.dirty[w] = false;
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  c.is_setup = true;
}

}  // namespace fnp
