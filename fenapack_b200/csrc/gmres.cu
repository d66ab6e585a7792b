// Right-preconditioned restarted (F)GMRES, device resident: KSPGMRES as PCDKSP
// configures it (fenapack/field_split.py:52-53; restart / rtol from
// demo_navier-stokes-pcd.py:146-148).  Classical Gram-Schmidt without
// refinement (PETSc's default).
//
// One iteration = PC apply (CUDA graph) + system MatMult + four kernels:
//   multidot   h_i = v_i . w for i <= j AND w . w, one pass over w per 8 vectors
//   [ONE all-reduce of j + 2 doubles on multi-rank contexts]
//   maxpy      w -= sum_i h_i v_i, fused with the explicit ||w||^2 of this rank's part
//   givens     one thread: norm (explicit on one rank; ||w||^2 - sum h_i^2 from the single
//              reduction on several, with a cancellation guard), Givens rotations of the new
//              Hessenberg column, residual estimate -- H, cs, sn, g never leave the device
//   scale      v_{j+1} = w / ||w||
// The host does not synchronise with the device inside an iteration: the residual estimate of
// iteration j travels to pinned memory asynchronously and is looked at after iteration j + 1
// has been enqueued (the convergence test of KSPGMRES on the recurrence estimate against
// max(rtol ||b||, atol), one iteration late), except when the estimates predict convergence
// within the next iteration -- then the host waits, so that no iteration is enqueued in vain.
// The solution is formed from exactly the columns up to the first iteration that met the
// tolerance (device-side back substitution), so iteration counts and iterates are those of
// the synchronous algorithm.
#include <cmath>

#include "fnp_internal.cuh"

namespace fnp {

namespace {

struct Basis {
  std::vector<DevBuf<double>> &store;
  DevBuf<double *> ptrs;   // device array of vector pointers
  std::vector<double *> host_ptrs;
  int64_t n;
  Basis(std::vector<DevBuf<double>> &s, int64_t n_, int capacity) : store(s), n(n_) {
    ptrs.alloc(capacity);
    host_ptrs.assign(capacity, nullptr);
  }
  // make sure vectors [0, upto] exist
  void ensure(Ctx &c, int upto) {
    bool changed = false;
    if ((int)store.size() <= upto) store.resize(upto + 1);
    for (int i = 0; i <= upto; ++i) {
      if (store[i].n < (size_t)n) { store[i].alloc((size_t)n); changed = true; }
      if (host_ptrs[i] != store[i].p) { host_ptrs[i] = store[i].p; changed = true; }
    }
    if (changed)
      FNP_CUDA(cudaMemcpyAsync(ptrs.p, host_ptrs.data(), host_ptrs.size() * sizeof(double *), cudaMemcpyHostToDevice,
                               c.stream));
  }
  double *operator[](int i) { return store[i].p; }
};

constexpr double ST_OK = 0.0, ST_NONFINITE = 1.0, ST_CANCELLATION = 2.0;

}  // namespace

// Device-side state of one restart cycle: H (m+1) x m column major, cs, sn, g, and per iteration
// the pair (residual estimate, status).
//   hcol[0..j]  = h_ij (all-reduced), hcol[j+1] = w . w before the orthogonalisation (all-reduced)
//   nrm2_local  = explicit ||w||^2 after the orthogonalisation (this rank's part unless all-reduced)
__global__ void givens_kernel(int j, int m, const double *__restrict__ hcol, const double *__restrict__ nrm2_local,
                              int pythagoras, double *__restrict__ H, double *__restrict__ cs, double *__restrict__ sn,
                              double *__restrict__ g, double *__restrict__ inv_norm, double *__restrict__ out /* res, status */) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double status = ST_OK;
  double nrm2;
  if (pythagoras) {
    const double ww = hcol[j + 1];
    double s = 0.0;
    for (int i = 0; i <= j; ++i) s += hcol[i] * hcol[i];
    nrm2 = ww - s;
    if (!(nrm2 > 1e-6 * ww)) status = ST_CANCELLATION;      // also catches NaN
  } else {
    nrm2 = nrm2_local[0];
  }
  const double hn = sqrt(nrm2 > 0.0 ? nrm2 : 0.0);
  if (!isfinite(hn) && status == ST_OK) status = ST_NONFINITE;
  inv_norm[0] = hn > 0.0 ? 1.0 / hn : 0.0;
  double *col = H + (size_t)j * (m + 1);
  for (int i = 0; i <= j; ++i) col[i] = hcol[i];
  col[j + 1] = hn;
  for (int i = 0; i < j; ++i) {
    const double a = col[i], b = col[i + 1];
    col[i] = cs[i] * a + sn[i] * b;
    col[i + 1] = -sn[i] * a + cs[i] * b;
  }
  const double a = col[j], b = col[j + 1];
  const double rho = hypot(a, b);
  double c_, s_;
  if (rho == 0.0) { c_ = 1.0; s_ = 0.0; } else { c_ = a / rho; s_ = b / rho; }
  cs[j] = c_;
  sn[j] = s_;
  col[j] = rho;
  col[j + 1] = 0.0;
  const double gj = g[j];
  g[j + 1] = -s_ * gj;
  g[j] = c_ * gj;
  const double r = fabs(g[j + 1]);
  if (!isfinite(r) && status == ST_OK) status = ST_NONFINITE;
  out[0] = r;
  out[1] = status;
}

// y = H(0:jdone, 0:jdone)^-1 g (upper triangular after the rotations)
__global__ void backsolve_kernel(int jdone, int m, const double *__restrict__ H, const double *__restrict__ g, double *__restrict__ y) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int i = jdone - 1; i >= 0; --i) {
    double s = g[i];
    for (int k = i + 1; k < jdone; ++k) s -= H[(size_t)k * (m + 1) + i] * y[k];
    y[i] = s / H[(size_t)i * (m + 1) + i];
  }
}

__global__ void set_g0_kernel(int m, double beta, double *__restrict__ g) {
  for (int i = threadIdx.x; i <= m; i += blockDim.x) g[i] = i == 0 ? beta : 0.0;
}

__global__ void scale_by_kernel(int64_t n, const double *__restrict__ s, const double *__restrict__ w, double *__restrict__ v) {
  const double a = s[0];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = w[i] * a;
}

static void allreduce_sum(Ctx &c, double *dev, int count) {
  if (c.nranks > 1) FNP_NCCL(nccl().AllReduce(dev, dev, count, ncclDouble, ncclSum, c.comm, c.stream));
}

static void solve_impl(Ctx &c, const double *b, double *x, int32_t *its_out, double *rnorm_out, int32_t *napply_out) {
  const int64_t n = c.n_u + c.n_p;
  const int m = c.restart;
  FNP_REQUIRE(m >= 1 && m <= 990, FNP_ERR_OPTION, "ksp_gmres_restart must be in [1, 990]");
  c.kr_w.ensure((size_t)n);
  c.red_out.ensure(1024);
  {
    // block partials of a full restart cycle, sized before the apply is captured: nothing that a
    // captured kernel addresses may be reallocated later
    const size_t need = (size_t)c.num_sms * 4 * (size_t)(m + 2 + 8);
    if (c.red_partial.n < need) {
      c.drop_graph();
      c.red_partial.ensure(need);
    }
  }
  // pinned mirror: [0] ||b||^2 / restart norm, then (res, status) per iteration of a cycle
  const size_t pin_need = 2 + 2 * (size_t)(m + 1);
  if (c.pinned_n < pin_need) {
    if (c.pinned) cudaFreeHost(c.pinned);
    c.pinned = nullptr;
    c.pinned_n = 0;
    FNP_CUDA(cudaMallocHost(reinterpret_cast<void **>(&c.pinned), pin_need * sizeof(double)));
    c.pinned_n = pin_need;
  }
  // device Krylov state
  c.kr_H.ensure((size_t)(m + 1) * m);
  c.kr_small.ensure((size_t)(5 * (m + 1) + 8));
  double *Hd = c.kr_H.p;
  double *cs = c.kr_small.p, *sn = cs + (m + 1), *g = sn + (m + 1), *resd = g + (m + 1), *invn = resd + 2 * (m + 1);
  if (c.ev_iter.empty()) {
    c.ev_iter.resize(2);
    for (auto &e : c.ev_iter) FNP_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  Basis V(c.V, n, m + 1), Z(c.Z, n, m + 1);
  double *w = c.kr_w.p;
  double *hdev = c.red_out.p;          // [0, j]: h column, [j+1]: w.w   ([1000..]: CG scalars, pcd.cu)
  double *nrm_loc = c.red_out.p + 996; // explicit ||w||^2 after the orthogonalisation
  double *hpin = c.pinned;

  // ||b||
  dot(c, n, b, b, hdev);
  allreduce_sum(c, hdev, 1);
  FNP_CUDA(cudaMemcpyAsync(hpin, hdev, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  const double bnorm = std::sqrt(hpin[0]);
  FNP_REQUIRE(std::isfinite(bnorm), FNP_ERR_NUMERIC, "right-hand side is not finite");
  const double tol = std::max(c.rtol * bnorm, c.atol);
  c.res_hist.clear();
  c.res_hist.push_back(bnorm);
  vec_zero(c, n, x);
  int its = 0, napply = 0;
  double res = bnorm;
  c.converged_reason = 0;
  *its_out = 0; *rnorm_out = bnorm; *napply_out = 0;
  if (bnorm <= tol) {
    c.converged_reason = bnorm <= c.atol ? 3 : 2;          // KSP_CONVERGED_ATOL / KSP_CONVERGED_RTOL
    return;
  }
  // first cycle: r = b (zero initial guess)
  V.ensure(c, 0);
  vec_scale_inv_sqrt(c, n, hdev, b, V[0]);
  double beta = bnorm;
  while (true) {
    set_g0_kernel<<<1, 256, 0, c.stream>>>(m, beta, g);
    c.launches++;
    int jdone = 0;                 // columns that enter the solution update
    bool converged = false;
    int enq = 0;                   // iterations enqueued in this cycle
    int seen = 0;                  // iterations whose result the host has looked at
    double prev_res = beta, last_res = beta;
    bool cancelled = false;
    // look at the result of iteration `seen` (waits for its event); true when the cycle ends there
    auto consume = [&]() -> bool {
      FNP_CUDA(cudaEventSynchronize(c.ev_iter[seen & 1]));
      const double r = hpin[2 + 2 * seen], st = hpin[2 + 2 * seen + 1];
      if (st == ST_CANCELLATION) {
        // the norm from the single reduction lost its digits: close the cycle before this column, update
        // x, restart from the true residual with the explicit norm (every rank sees the same numbers)
        cancelled = true;
        jdone = seen;
        return true;
      }
      FNP_REQUIRE(st == ST_OK && std::isfinite(r), FNP_ERR_NUMERIC, "GMRES breakdown: non-finite Hessenberg entry");
      prev_res = last_res;
      last_res = r;
      ++seen;
      ++its;
      res = r;
      c.res_hist.push_back(r);
      jdone = seen;
      if (r <= tol || its >= c.max_it) {
        converged = r <= tol;
        return true;
      }
      return false;
    };
    bool cycle_over = false;
    for (int j = 0; j < m && !cycle_over; ++j) {
      V.ensure(c, j + 1);
      double *zj;
      if (c.flexible) {
        Z.ensure(c, j);
        zj = Z[j];
      } else {
        c.kr_x.ensure((size_t)n);
        zj = c.kr_x.p;
      }
      // z = M^-1 v_j ; w = A z
      pc_apply_vec(c, V[j], zj);
      ++napply;
      system_matvec(c, zj, w);
      // Several ranks: ||w||^2 - sum h^2 from the single reduction presumes an orthonormal basis; classical
      // Gram-Schmidt loses orthogonality like (residual reduction)^2 * eps, so the fused norm is used while
      // the reduction inside the cycle is below 1e4 (norm error < 1e-6 relative) and the explicit norm with
      // its own all-reduce afterwards.  Decided from the last residual estimate seen: identical on all ranks.
      const bool pythagoras = c.nranks > 1 && !c.gmres_two_reductions && last_res >= 1e-4 * beta;
      {
        // classical Gram-Schmidt, one reduction
        StageTimer tgs(c, "FENaPack: GMRES orthogonalization");
        multi_dot_ww_ptrs(c, n, V.ptrs.p, j + 1, w, hdev);
        allreduce_sum(c, hdev, j + 2);
        multi_axpy_norm_ptrs(c, n, V.ptrs.p, j + 1, hdev, w, nrm_loc);
        if (c.nranks > 1 && !pythagoras) allreduce_sum(c, nrm_loc, 1);
        givens_kernel<<<1, 32, 0, c.stream>>>(j, m, hdev, nrm_loc, pythagoras ? 1 : 0, Hd, cs, sn, g, invn, resd + 2 * j);
        c.launches++;
        scale_by_kernel<<<c.num_sms * 8, 256, 0, c.stream>>>(n, invn, w, V[j + 1]);
        c.launches++;
        FNP_CUDA(cudaPeekAtLastError());
      }
      // the entry of the ring used by iteration j - 2 has been consumed (seen >= j - 1 below)
      FNP_CUDA(cudaMemcpyAsync(hpin + 2 + 2 * j, resd + 2 * j, 2 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
      FNP_CUDA(cudaEventRecord(c.ev_iter[j & 1], c.stream));
      enq = j + 1;
      // the result of iteration j - 1 is there or about to be: look at it without draining the queue
      while (seen < enq - 1 && !cycle_over) cycle_over = consume();
      if (cycle_over) break;
      // wait for iteration j itself when the next one would probably be enqueued in vain, when the
      // timers ask for a clean attribution, or at the end of the cycle
      const double ratio = prev_res > 0.0 ? std::min(1.0, last_res / prev_res) : 1.0;
      const bool near = last_res * ratio * ratio <= 2.0 * tol || its + 2 >= c.max_it;
      if (near || j + 1 == m || c.timers_on || c.gmres_sync) cycle_over = consume();
    }
    while (!cycle_over && seen < enq) cycle_over = consume();
    if (cancelled) c.gmres_two_reductions = true;
    // y = H^-1 g (back substitution on the device) ; x += Z y  or  x += M^-1 (V y)
    double *ydev = c.red_out.p;   // reuse the h column slot
    backsolve_kernel<<<1, 32, 0, c.stream>>>(jdone, m, Hd, g, ydev);
    c.launches++;
    if (c.flexible) {
      multi_axpy_ptrs(c, n, Z.ptrs.p, jdone, ydev, x);
    } else {
      c.kr_b.ensure((size_t)n);
      vec_zero(c, n, c.kr_b.p);
      multi_axpy_ptrs(c, n, V.ptrs.p, jdone, ydev, c.kr_b.p);
      pc_apply_vec(c, c.kr_b.p, c.kr_x.p);
      ++napply;
      vec_axpy(c, n, 1.0, c.kr_x.p, x);
    }
    *its_out = its; *rnorm_out = res; *napply_out = napply;
    if (converged || (its >= c.max_it && !cancelled)) break;
    // restart: r = b - A x
    system_matvec(c, x, w);
    vec_axpby(c, n, 1.0, b, -1.0, w, w);
    dot(c, n, w, w, hdev);
    allreduce_sum(c, hdev, 1);
    FNP_CUDA(cudaMemcpyAsync(hpin, hdev, sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    vec_scale_inv_sqrt(c, n, hdev, w, V[0]);
    FNP_CUDA(cudaStreamSynchronize(c.stream));
    beta = std::sqrt(hpin[0]);
    res = beta;
    *rnorm_out = res;
    if (beta <= tol) break;
  }
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  // KSPConvergedReason as KSPConvergedDefault reports it: KSP_CONVERGED_ATOL 3, KSP_CONVERGED_RTOL 2, KSP_DIVERGED_ITS -3
  c.converged_reason = res <= tol ? (res <= c.atol ? 3 : 2) : -3;
}

void solve_fgmres(Ctx &c, const double *b, double *x, int32_t *its_out, double *rnorm_out, int32_t *napply_out) {
  StageTimer t(c, "FENaPack: PCDKSP solve");
  FNP_REQUIRE(c.is_setup, FNP_ERR_STATE, "fnp_solve before fnp_setup");
  solve_impl(c, b, x, its_out, rnorm_out, napply_out);
}

}  // namespace fnp
