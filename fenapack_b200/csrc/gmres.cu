// Right-preconditioned restarted (F)GMRES, device resident: KSPGMRES as PCDKSP
// configures it (fenapack/field_split.py:52-53; restart / rtol from
// demo_navier-stokes-pcd.py:146-148).  Classical Gram-Schmidt without
// refinement (PETSc's default): one fused multi-dot, one fused MAXPY+norm.
// Only the (j+2) Hessenberg entries of an iteration cross PCIe; the Givens
// recurrence and the convergence test (recurrence estimate of the residual norm
// against max(rtol*||b||, atol)) run on the host.
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
  // make sure vectors [0, upto] exist; allocate in chunks of 4 vectors
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

}  // namespace

static void allreduce_sum(Ctx &c, double *dev, int count) {
  if (c.nranks > 1) FNP_NCCL(nccl().AllReduce(dev, dev, count, ncclDouble, ncclSum, c.comm, c.stream));
}

void solve_fgmres(Ctx &c, const double *b, double *x, int32_t *its_out, double *rnorm_out, int32_t *napply_out) {
  StageTimer t(c, "FENaPack: PCDKSP solve");
  FNP_REQUIRE(c.is_setup, FNP_ERR_STATE, "fnp_solve before fnp_setup");
  const int64_t n = c.n_u + c.n_p;
  const int m = c.restart;
  FNP_REQUIRE(m >= 1 && m <= 990, FNP_ERR_OPTION, "ksp_gmres_restart must be in [1, 990]");
  c.kr_w.ensure((size_t)n);
  c.red_out.ensure(1024);
  {
    // block partials of a full restart cycle, sized before the apply is captured: nothing that a
    // captured kernel addresses may be reallocated later
    const size_t need = (size_t)c.num_sms * 4 * (size_t)(m + 1 + 8);
    if (c.red_partial.n < need) {
      c.drop_graph();
      c.red_partial.ensure(need);
    }
  }
  if (!c.pinned) {
    FNP_CUDA(cudaMallocHost(reinterpret_cast<void **>(&c.pinned), 1024 * sizeof(double)));
    c.pinned_n = 1024;
  }
  Basis V(c.V, n, m + 1), Z(c.Z, n, m + 1);
  double *w = c.kr_w.p;
  double *hdev = c.red_out.p;          // [0, m]: h column, [m+1]: norm^2
  double *hpin = c.pinned;

  std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), g(m + 1);
  auto Hat = [&](int i, int j) -> double & { return H[(size_t)j * (m + 1) + i]; };

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
  if (bnorm <= tol) {
    *its_out = 0; *rnorm_out = bnorm; *napply_out = 0;
    return;
  }
  // first cycle: r = b (zero initial guess)
  V.ensure(c, 0);
  vec_scale_inv_sqrt(c, n, hdev, b, V[0]);
  double beta = bnorm;
  bool done = false;
  while (!done) {
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = beta;
    int jdone = 0;
    bool converged = false;
    for (int j = 0; j < m; ++j) {
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
      // classical Gram-Schmidt
      StageTimer tgs(c, "FENaPack: GMRES orthogonalization");
      multi_dot_ptrs(c, n, V.ptrs.p, j + 1, w, hdev);
      allreduce_sum(c, hdev, j + 1);
      multi_axpy_norm_ptrs(c, n, V.ptrs.p, j + 1, hdev, w, hdev + j + 1);
      allreduce_sum(c, hdev + j + 1, 1);
      vec_scale_inv_sqrt(c, n, hdev + j + 1, w, V[j + 1]);
      FNP_CUDA(cudaMemcpyAsync(hpin, hdev, (j + 2) * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
      FNP_CUDA(cudaStreamSynchronize(c.stream));
      for (int i = 0; i <= j; ++i) Hat(i, j) = hpin[i];
      Hat(j + 1, j) = std::sqrt(hpin[j + 1]);
      FNP_REQUIRE(std::isfinite(Hat(j + 1, j)), FNP_ERR_NUMERIC, "GMRES breakdown: non-finite Hessenberg entry");
      // Givens
      for (int i = 0; i < j; ++i) {
        const double a = Hat(i, j), bb = Hat(i + 1, j);
        Hat(i, j) = cs[i] * a + sn[i] * bb;
        Hat(i + 1, j) = -sn[i] * a + cs[i] * bb;
      }
      {
        const double a = Hat(j, j), bb = Hat(j + 1, j);
        const double rho = std::hypot(a, bb);
        if (rho == 0.0) { cs[j] = 1.0; sn[j] = 0.0; } else { cs[j] = a / rho; sn[j] = bb / rho; }
        Hat(j, j) = rho;
        Hat(j + 1, j) = 0.0;
        g[j + 1] = -sn[j] * g[j];
        g[j] = cs[j] * g[j];
      }
      ++its;
      jdone = j + 1;
      res = std::fabs(g[j + 1]);
      c.res_hist.push_back(res);
      if (res <= tol || its >= c.max_it) {
        converged = res <= tol;
        break;
      }
    }
    // y = H^-1 g (back substitution) ; x += Z y  or  x += M^-1 (V y)
    std::vector<double> y(jdone);
    for (int i = jdone - 1; i >= 0; --i) {
      double s = g[i];
      for (int k = i + 1; k < jdone; ++k) s -= Hat(i, k) * y[k];
      y[i] = s / Hat(i, i);
    }
    double *ydev = c.red_out.p;   // reuse the h column slot
    FNP_CUDA(cudaMemcpyAsync(ydev, y.data(), jdone * sizeof(double), cudaMemcpyHostToDevice, c.stream));
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
    FNP_CUDA(cudaStreamSynchronize(c.stream));   // y (host) must outlive the copy
    if (converged || its >= c.max_it) break;
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
    if (beta <= tol) break;
  }
  *its_out = its;
  *rnorm_out = res;
  *napply_out = napply;
}

}  // namespace fnp
