// Device side of the inner solvers: Chebyshev-Jacobi sweeps (the Mp solve of the
// reference's "iterative" set-up, demo_navier-stokes-pcd.py:161-165, and the AMG
// smoother) and the smoothed-aggregation V-cycle that stands in for hypre
// BoomerAMG (demo_navier-stokes-pcd.py:153-160).  Every step is one SpMV-class
// kernel with a fused epilogue (kernels.cu).
#include <cmath>

#include "fnp_internal.cuh"

namespace fnp {

// KSPCHEBYSHEV + PCJACOBI recurrence (PETSc cheby.c, recalled -- SURVEY 8a row 9):
//   s = 2/(emax+emin), alpha = 1 - s emin, mu = 1/alpha, omegaprod = 2/alpha, c0 = 1, c1 = mu
//   p1 = s D^-1 b
//   pass: c2 = 2 mu c1 - c0; omega = omegaprod c1/c2; p2 = (1-omega) p0 + omega p1 + omega s D^-1 (b - A p1)
// `steps` = number of Jacobi applications.  Result: out = add + out_scale * p_last.
void cheb_jacobi(Ctx &c, const DevCsr &A, const double *b, double emin, double emax, int steps, double out_scale,
                 const double *add, double *out, double *w0, double *w1, bool p1_ready) {
  FNP_REQUIRE(A.has_dinv, FNP_ERR_STATE, "Chebyshev-Jacobi: operator has no Jacobi diagonal (fnp_setup missing)");
  FNP_REQUIRE(steps >= 1, FNP_ERR_ARG, "Chebyshev-Jacobi needs ksp_max_it >= 1");
  const int64_t n = A.vec_rows();
  const double s = 2.0 / (emax + emin);
  const double alpha = 1.0 - s * emin;
  const double mu = 1.0 / alpha;
  const double omegaprod = 2.0 / alpha;
  if (steps == 1) {
    vec_pointwise_scale(c, n, s * out_scale, A.dinv.p, b, add, out);
    return;
  }
  // rotating buffers; the last pass writes straight into `out`
  double *buf[2] = {w0, w1};
  double *p0 = nullptr;
  double *p1 = buf[0];
  if (!p1_ready) vec_pointwise_scale(c, n, s, A.dinv.p, b, nullptr, p1);     // else: written by the producer of b
  double c0 = 1.0, c1 = mu;
  for (int pass = 1; pass < steps; ++pass) {
    const double c2 = 2.0 * mu * c1 - c0;
    const double omega = omegaprod * c1 / c2;
    const bool last = pass == steps - 1;
    double *p2 = last ? out : (p0 ? p0 : buf[1]);
    const double sc = last ? out_scale : 1.0;
    EpiCheb e;
    e.out = p2;
    e.p0 = p0;
    e.p1 = p1;
    e.b = b;
    e.dinv = A.dinv.p;
    e.add = last ? add : nullptr;
    e.c0 = sc * (1.0 - omega);
    e.c1 = sc * omega;
    e.c2 = sc * omega * s;
    spmv_cheb(c, A, e);
    p0 = p1;
    p1 = p2;
    c0 = c1;
    c1 = c2;
  }
}

std::shared_ptr<HaloPlan> expand_plan(Ctx &c, const HaloPlan &p, int bs);

static void upload_csr(Ctx &c, const HostCsr &h, DevCsr &d, const std::string &tag, int bs,
                       std::shared_ptr<HaloPlan> halo = nullptr, int64_t n_own = -1) {
  csr_upload_pattern(c, d, h, tag, (halo && c.split_rows(h.nrows)) ? n_own : -1, bs, halo ? n_own : -1);
  csr_set_values(c, d, h, h.val.data(), false);
  if (halo) {
    d.halo = bs > 1 ? expand_plan(c, *halo, bs) : halo;
    d.ncols_own = (int32_t)n_own;
    d.nghost = halo->nghost;
  }
}

// `bs` > 1: the host hierarchy is the one of the scalar operator S; every level acts on
// bs interleaved right-hand sides (A_l (x) I_bs, P_l (x) I_bs, ...).
void amg_upload(Ctx &c, DevHierarchy &H, const std::string &name, DevCsr *level0, int bs) {
  const size_t L = H.host.levels.size();
  H.levels.clear();
  H.levels.resize(L);
  H.refresh_W.clear();          // plans of the device-side refresh belong to the old hierarchy
  H.refresh_row0.clear();
  H.refresh_diag.clear();
  H.refresh_built = false;
  for (size_t l = 0; l < L; ++l) {
    const HostLevel &hl = H.host.levels[l];
    DevLevel &dl = H.levels[l];
    if (l == 0 && level0) {
      dl.Ap = level0;
      FNP_REQUIRE(level0->has_dinv && level0->nrows == hl.A.nrows && level0->bs == bs, FNP_ERR_STATE,
                  "AMG level 0 operator mismatch");
    } else {
      dl.Ap = &dl.A_own;
      upload_csr(c, hl.A, dl.A_own, name + "/L" + std::to_string(l), bs, hl.halo, hl.n_own);
      std::vector<double> dinv((size_t)hl.dinv.size() * bs);
      for (size_t i = 0; i < hl.dinv.size(); ++i)
        for (int b = 0; b < bs; ++b) dinv[i * bs + b] = hl.dinv[i];
      dl.A_own.dinv.upload(dinv.data(), dinv.size(), c.stream);
      FNP_CUDA(cudaStreamSynchronize(c.stream));
      dl.A_own.has_dinv = true;
    }
    dl.rho = hl.rho;
    if (l + 1 < L || H.host.tail) {
      upload_csr(c, hl.P, dl.P, name + "/P" + std::to_string(l), bs);
      upload_csr(c, hl.R, dl.R, name + "/R" + std::to_string(l), bs);
    }
    const size_t n = (size_t)hl.A.nrows * bs;
    dl.x.alloc(n); dl.b.alloc(n); dl.r.alloc(n); dl.w0.alloc(n); dl.w1.alloc(n);
  }
  if (H.host.tail) {
    // replicated tail: a serial hierarchy on every rank, fed by one padded all-gather
    const std::vector<int64_t> &tb = H.host.tail_begins;
    const int R = c.nranks;
    int64_t maxloc = 0;
    for (int q = 0; q < R; ++q) maxloc = std::max(maxloc, tb[q + 1] - tb[q]);
    H.tail_n = tb[R] * bs;
    H.tail_nloc = (tb[c.rank + 1] - tb[c.rank]) * bs;
    H.tail_off = tb[c.rank] * bs;
    H.tail_maxloc = maxloc * bs;
    std::vector<int64_t> map((size_t)H.tail_n);
    for (int q = 0; q < R; ++q)
      for (int64_t t = 0; t < (tb[q + 1] - tb[q]) * bs; ++t) map[(size_t)(tb[q] * bs + t)] = (int64_t)q * H.tail_maxloc + t;
    H.tail_map.upload(map.data(), map.size(), c.stream);
    H.tail_gather.alloc((size_t)(R + 1) * H.tail_maxloc);
    H.tail_gather.zero(c.stream);
    H.tail_b.alloc((size_t)H.tail_n);
    H.tail_x.alloc((size_t)H.tail_n);
    H.tail_bloc.alloc((size_t)std::max<int64_t>(H.tail_nloc, 1));
    FNP_CUDA(cudaStreamSynchronize(c.stream));
    H.tail.reset(new DevHierarchy());
    H.tail->params = H.params;
    H.tail->serial = true;
    H.tail->host = *H.host.tail;
    const int saved_rank = c.rank, saved_n = c.nranks;
    c.rank = 0;
    c.nranks = 1;
    try {
      amg_upload(c, *H.tail, name + "/T", nullptr, bs);
    } catch (...) {
      c.rank = saved_rank;
      c.nranks = saved_n;
      throw;
    }
    c.rank = saved_rank;
    c.nranks = saved_n;
    H.built = true;
    return;
  }
  // coarsest level: dense inverse
  const int64_t nc = H.host.levels.back().A.nrows, cols = H.host.coarse_cols;
  H.coarse_n = (int)(nc * bs);
  H.coarse_cols = (int)(cols * bs);
  H.coarse_maxloc = (int)(H.host.coarse_maxloc * bs);
  if (c.nranks > 1) H.coarse_gather.alloc((size_t)(c.nranks + 1) * H.coarse_maxloc);
  // the scalar inverse serves the bs interleaved components (dense_gemv)
  H.coarse_bs = bs;
  H.coarse_inv.upload(H.host.coarse_inv.data(), H.host.coarse_inv.size(), c.stream);
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  H.built = true;
}

static void vcycle_level(Ctx &c, DevHierarchy &H, size_t l, const double *b, double *x, bool b_has_first_step = false);

// coarse part replicated on every rank: all-gather the restricted residual, run the serial
// tail V-cycle, keep the own slice of the correction (returned pointer)
static const double *tail_solve(Ctx &c, DevHierarchy &H, const double *b_loc) {
  double *slot = H.tail_gather.p + (size_t)c.nranks * H.tail_maxloc;
  if (H.tail_nloc) FNP_CUDA(cudaMemcpyAsync(slot, b_loc, H.tail_nloc * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
  FNP_NCCL(nccl().AllGather(slot, H.tail_gather.p, (size_t)H.tail_maxloc, ncclDouble, c.comm, c.stream));
  vec_gather(c, H.tail_n, H.tail_map.p, H.tail_gather.p, H.tail_b.p);
  vcycle_level(c, *H.tail, 0, H.tail_b.p, H.tail_x.p);
  return H.tail_x.p + H.tail_off;
}

static void vcycle_level(Ctx &c, DevHierarchy &H, size_t l, const double *b, double *x, bool b_has_first_step) {
  DevLevel &L = H.levels[l];
  const bool last = l + 1 == H.levels.size();
  if (last && !H.tail) {
    if (c.nranks == 1 || H.serial) {
      dense_gemv(c, H.coarse_n / H.coarse_bs, H.coarse_cols / H.coarse_bs, H.coarse_bs, H.coarse_inv.p, b, x);
    } else {
      // padded all-gather of the coarse right-hand side, then this rank's rows of the inverse
      double *slot = H.coarse_gather.p + (size_t)c.nranks * H.coarse_maxloc;
      FNP_CUDA(cudaMemsetAsync(slot, 0, H.coarse_maxloc * sizeof(double), c.stream));
      if (H.coarse_n) FNP_CUDA(cudaMemcpyAsync(slot, b, H.coarse_n * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
      FNP_NCCL(nccl().AllGather(slot, H.coarse_gather.p, (size_t)H.coarse_maxloc, ncclDouble, c.comm, c.stream));
      dense_gemv(c, H.coarse_n / H.coarse_bs, H.coarse_cols / H.coarse_bs, H.coarse_bs, H.coarse_inv.p, H.coarse_gather.p, x);
    }
    return;
  }
  const AmgParams &p = H.params;
  const double emax = L.rho, emin = L.rho / p.eig_ratio;
  // With >= 2 smoothing steps the first Chebyshev iterate s D^-1 b is a by-product of the kernel that
  // produces b: the restriction of the level above (pre-smoothing) and the residual (post-smoothing)
  // write it into the smoother's first buffer, which saves two launches per level.
  // (Not on large levels: there the epilogue's extra loads and stores slow the gather-bound SpMV
  // kernel by more than a streaming kernel of three vectors costs.)
  auto fuses = [&](DevLevel &lv) { return p.smooth_steps >= 2 && lv.A().vec_rows() < 1500000; };
  const bool fuse = fuses(L);
  auto first_step = [&](DevLevel &lv) {
    Jacobi1 j;
    if (fuses(lv)) {
      j.y2 = lv.w0.p;
      j.d2 = lv.A().dinv.p;
      j.s2 = cheb_first_step_scale(lv.rho / p.eig_ratio, lv.rho);
    }
    return j;
  };
  // pre-smoothing from the zero initial guess
  cheb_jacobi(c, L.A(), b, emin, emax, p.smooth_steps, 1.0, nullptr, x, L.w0.p, L.w1.p, fuse && b_has_first_step);
  // r = b - A x ; b_c = R r
  spmv_axpby(c, L.A(), x, -1.0, 1.0, b, L.r.p);
  const double *xc;
  if (last) {          // the next level lives in the replicated tail
    spmv_store(c, L.R, L.r.p, H.tail_bloc.p);
    xc = tail_solve(c, H, H.tail_bloc.p);
  } else {
    DevLevel &C = H.levels[l + 1];
    const bool c_smooths = !(l + 2 == H.levels.size() && !H.tail);      // the coarsest level is solved, not smoothed
    spmv_store(c, L.R, L.r.p, C.b.p, c_smooths ? first_step(C) : Jacobi1());
    vcycle_level(c, H, l + 1, C.b.p, C.x.p, c_smooths && fuses(C));
    xc = C.x.p;
  }
  // x += P x_c
  spmv_axpby(c, L.P, xc, 1.0, 1.0, x, x);
  // post-smoothing on the correction equation: x += cheb(A, b - A x)
  spmv_axpby(c, L.A(), x, -1.0, 1.0, b, L.r.p, first_step(L));
  cheb_jacobi(c, L.A(), L.r.p, emin, emax, p.smooth_steps, 1.0, x, x, L.w0.p, L.w1.p, fuse);
}

void amg_vcycle(Ctx &c, DevHierarchy &H, const double *b, double *x) {
  FNP_REQUIRE(H.built, FNP_ERR_STATE, "AMG hierarchy not built (fnp_setup missing)");
  vcycle_level(c, H, 0, b, x);
}

}  // namespace fnp
