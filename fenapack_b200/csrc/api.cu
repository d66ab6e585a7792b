// extern "C" surface of libfenapack_cuda (include/fenapack_cuda.h).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "fnp_internal.cuh"

namespace fnp {

std::vector<int64_t> comm_ranges(Ctx &c, int64_t n_local);
std::shared_ptr<HaloPlan> build_halo(Ctx &c, HostCsr &h, const std::vector<int64_t> &begins,
                                     std::vector<int64_t> *ghost_global_out);
std::shared_ptr<HaloPlan> expand_plan(Ctx &c, const HaloPlan &p, int bs);

static thread_local std::string g_last_error;
void set_last_error(const std::string &m) { g_last_error = m; }

Ctx::Ctx(int dev) : device(dev) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw Error(FNP_ERR_CUDA, std::string("no usable CUDA device (libfenapack_cuda has no CPU fallback): ") +
                                  cudaGetErrorString(e));
  FNP_REQUIRE(dev >= 0 && dev < count, FNP_ERR_ARG, "device index out of range");
  FNP_CUDA(cudaSetDevice(dev));
  cudaDeviceProp prop;
  FNP_CUDA(cudaGetDeviceProperties(&prop, dev));
  num_sms = prop.multiProcessorCount;
  FNP_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  own_stream = true;
  FNP_CUDA(cudaEventCreate(&ev_tic));
  FNP_CUDA(cudaEventCreate(&ev_toc));
  variant = 1;
  // defaults of the reference's "iterative" set-up (demo_navier-stokes-pcd.py:153-165)
  opt_u.ksp = KSP_RICHARDSON; opt_u.pc = PC_AMG; opt_u.max_it = 1;
  opt_ap.ksp = KSP_RICHARDSON; opt_ap.pc = PC_AMG; opt_ap.max_it = 2;
  opt_mp.ksp = KSP_CHEBYSHEV; opt_mp.pc = PC_JACOBI; opt_mp.max_it = 5; opt_mp.emin = 0.5; opt_mp.emax = 2.0;
  opt_rp.ksp = KSP_RICHARDSON; opt_rp.pc = PC_AMG; opt_rp.max_it = 1;     // demo_unsteady-navier-stokes-pcdr.py:167-170
}

Ctx::~Ctx() {
  cudaSetDevice(device);
  if (stream) cudaStreamSynchronize(stream);
  drop_graph();
  if (comm_stream) { cudaStreamSynchronize(comm_stream); cudaStreamDestroy(comm_stream); }
  if (ev_x) cudaEventDestroy(ev_x);
  if (ev_halo) cudaEventDestroy(ev_halo);
  if (comm_halo) nccl().CommDestroy(comm_halo);
  if (comm) nccl().CommDestroy(comm);
  if (pinned) cudaFreeHost(pinned);
  for (auto &p : pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
  for (auto e : event_pool) cudaEventDestroy(e);
  if (ev_tic) cudaEventDestroy(ev_tic);
  if (ev_toc) cudaEventDestroy(ev_toc);
  if (own_stream && stream) cudaStreamDestroy(stream);
}

cudaEvent_t Ctx::get_event() {
  if (!event_pool.empty()) {
    cudaEvent_t e = event_pool.back();
    event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  FNP_CUDA(cudaEventCreate(&e));
  return e;
}

void Ctx::resolve_timers() {
  if (pending.empty()) return;
  FNP_CUDA(cudaStreamSynchronize(stream));
  for (auto &p : pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      Timer &t = timers[p.name];
      t.ms += ms;
      t.calls += 1;
    }
    event_pool.push_back(p.a);
    event_pool.push_back(p.b);
  }
  pending.clear();
}

StageTimer::StageTimer(Ctx &ctx, const char *nm, int level) : c(ctx) {
  if (c.timers_on < level) return;
  name = nm;
  a = c.get_event();
  cudaEventRecord(a, c.stream);
}
StageTimer::StageTimer(Ctx &ctx, const std::string &nm, int level) : c(ctx) {
  if (c.timers_on < level) return;
  name = nm;
  a = c.get_event();
  cudaEventRecord(a, c.stream);
}
StageTimer::~StageTimer() {
  if (!a) return;
  cudaEvent_t b = c.get_event();
  cudaEventRecord(b, c.stream);
  c.pending.push_back({name, a, b});
  if (c.pending.size() > 200000) c.resolve_timers();
}

// ---------------------------------------------------------------------------
// options
// ---------------------------------------------------------------------------
static bool starts_with(const std::string &s, const std::string &p) { return s.compare(0, p.size(), p) == 0; }

static int parse_int(const std::string &name, const std::string &v) {
  char *end = nullptr;
  long r = std::strtol(v.c_str(), &end, 10);
  FNP_REQUIRE(end && *end == '\0' && !v.empty(), FNP_ERR_OPTION, "option " + name + ": not an integer: '" + v + "'");
  return (int)r;
}
static double parse_real(const std::string &name, const std::string &v) {
  char *end = nullptr;
  double r = std::strtod(v.c_str(), &end);
  FNP_REQUIRE(end && *end == '\0' && !v.empty(), FNP_ERR_OPTION, "option " + name + ": not a real: '" + v + "'");
  return r;
}

static void set_inner_option(InnerOpts &o, const std::string &full, const std::string &key, const std::string &v) {
  if (key == "ksp_type") {
    if (v == "preonly") o.ksp = KSP_PREONLY;
    else if (v == "richardson") o.ksp = KSP_RICHARDSON;
    else if (v == "chebyshev") o.ksp = KSP_CHEBYSHEV;
    else if (v == "cg") o.ksp = KSP_CG;
    else throw Error(FNP_ERR_OPTION, "option " + full + ": unsupported KSP type '" + v + "'");
  } else if (key == "pc_type") {
    if (v == "jacobi") o.pc = PC_JACOBI;
    else if (v == "none") o.pc = PC_NONE;
    else if (v == "amg" || v == "gamg" || v == "hypre" || v == "boomeramg") o.pc = PC_AMG;
    else throw Error(FNP_ERR_OPTION, "option " + full + ": unsupported PC type '" + v +
                                         "' (sparse direct solves are CPU-only in the reference and are not provided)");
  } else if (key == "pc_hypre_type") {
    FNP_REQUIRE(v == "boomeramg", FNP_ERR_OPTION, "option " + full + ": only boomeramg is mapped (to the SA-AMG V-cycle)");
    o.pc = PC_AMG;
  } else if (key == "ksp_max_it") {
    o.max_it = parse_int(full, v);
    FNP_REQUIRE(o.max_it >= 1, FNP_ERR_OPTION, "option " + full + ": must be >= 1");
  } else if (key == "ksp_rtol") {
    o.rtol = parse_real(full, v);
  } else if (key == "ksp_chebyshev_eigenvalues") {
    const size_t comma = v.find(',');
    FNP_REQUIRE(comma != std::string::npos, FNP_ERR_OPTION, "option " + full + ": expected 'emin, emax'");
    auto trim = [](std::string s) {
      const size_t a = s.find_first_not_of(" \t"), b = s.find_last_not_of(" \t");
      return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    };
    o.emin = parse_real(full, trim(v.substr(0, comma)));
    o.emax = parse_real(full, trim(v.substr(comma + 1)));
    FNP_REQUIRE(o.emin > 0 && o.emax > o.emin, FNP_ERR_OPTION, "option " + full + ": need 0 < emin < emax");
  } else if (key == "pc_amg_threshold") {
    o.amg.theta = parse_real(full, v);
  } else if (key == "pc_amg_levels") {
    o.amg.max_levels = parse_int(full, v);
  } else if (key == "pc_amg_refresh") {
    if (v == "rebuild") o.amg.refresh = 0;
    else if (v == "galerkin") o.amg.refresh = 1;
    else throw Error(FNP_ERR_OPTION, "pc_amg_refresh: rebuild | galerkin");
  } else if (key == "pc_amg_coarse_size") {
    o.amg.coarse_size = parse_int(full, v);
  } else if (key == "pc_amg_smooth_steps") {
    o.amg.smooth_steps = parse_int(full, v);
    FNP_REQUIRE(o.amg.smooth_steps >= 1, FNP_ERR_OPTION, "option " + full + ": must be >= 1");
  } else if (key == "pc_amg_prolongator_truncation") {
    o.amg.p_trunc = parse_real(full, v);
  } else if (key == "pc_amg_lag") {
    o.amg.lag = parse_int(full, v);
    FNP_REQUIRE(o.amg.lag >= 1, FNP_ERR_OPTION, "option " + full + ": must be >= 1");
  } else if (key == "pc_amg_replicate_size") {
    o.amg.replicate_size = parse_int(full, v);
  } else if (key == "pc_amg_coarse_drop") {
    o.amg.coarse_drop = parse_real(full, v);
  } else if (key == "pc_amg_eig_ratio") {
    o.amg.eig_ratio = parse_real(full, v);
  } else {
    throw Error(FNP_ERR_OPTION, "unknown option '" + full + "'");
  }
}

static void set_option(Ctx &c, const std::string &name, const std::string &v) {
  if (name.rfind("fnp_", 0) != 0) c.drop_graph();      // solver options are baked into the captured apply
  const std::string pu = "fieldsplit_u_", pap = "fieldsplit_p_PCD_Ap_", pmp = "fieldsplit_p_PCD_Mp_";
  const std::string prp = "fieldsplit_p_PCD_Rp_";
  if (starts_with(name, prp)) return set_inner_option(c.opt_rp, name, name.substr(prp.size()), v);
  if (starts_with(name, pap)) return set_inner_option(c.opt_ap, name, name.substr(pap.size()), v);
  if (starts_with(name, pmp)) return set_inner_option(c.opt_mp, name, name.substr(pmp.size()), v);
  if (starts_with(name, pu)) return set_inner_option(c.opt_u, name, name.substr(pu.size()), v);
  if (name == "fieldsplit_p_pc_python_type") {
    if (v == "fenapack.PCDPC_BRM1" || v == "BRM1") c.variant = 1;
    else if (v == "fenapack.PCDPC_BRM2" || v == "BRM2") c.variant = 2;
    else if (v == "fenapack.PCDRPC_BRM1" || v == "PCDR_BRM1") c.variant = 3;
    else if (v == "fenapack.PCDRPC_BRM2" || v == "PCDR_BRM2") c.variant = 4;
    else throw Error(FNP_ERR_OPTION, "option " + name + ": unsupported PCD class '" + v + "'");
  } else if (name == "ksp_type") {
    if (v == "gmres") c.flexible = false;
    else if (v == "fgmres") c.flexible = true;
    else throw Error(FNP_ERR_OPTION, "option ksp_type: only gmres and fgmres are provided");
  } else if (name == "ksp_gmres_restart") {
    c.restart = parse_int(name, v);
  } else if (name == "ksp_rtol") {
    c.rtol = parse_real(name, v);
  } else if (name == "ksp_atol") {
    c.atol = parse_real(name, v);
  } else if (name == "ksp_max_it") {
    c.max_it = parse_int(name, v);
  } else if (name == "ksp_pc_side") {
    FNP_REQUIRE(v == "right", FNP_ERR_OPTION, "PCDKSP uses right preconditioning only (field_split.py:53)");
  } else if (name == "fnp_spmv_kernel") {
    if (v == "auto") c.spmv_mode = 0;
    else if (v == "csr") c.spmv_mode = 1;
    else if (v == "sell") c.spmv_mode = 2;
    else throw Error(FNP_ERR_OPTION, "option fnp_spmv_kernel: auto | csr | sell (takes effect at fnp_set_pattern)");
  } else if (name == "fnp_halo_p2p") {
    c.p2p = parse_int(name, v);
  } else if (name == "fnp_reorder_nodes") {
    c.reorder = std::max(0, (int)parse_int(name, v)) / 6 * 6;   // windows hold whole nodes for 2 and 3 components
  } else if (name == "fnp_sell_gather") {
    c.sell_gather = (int)parse_int(name, v);
    c.drop_graph();
  } else if (name == "fnp_sell_sigma") {
    c.sell_sigma = std::max(32, (int)parse_int(name, v) / 32 * 32);
  } else if (name == "fnp_sell_max_mean_row") {
    c.sell_max_mean_row = parse_real(name, v);
  } else if (name == "fnp_prune_zeros") {
    c.prune = parse_int(name, v);
  } else if (name == "fnp_kronecker") {
    c.kron = parse_int(name, v);
  } else if (name == "fnp_halo_overlap") {
    c.overlap = parse_int(name, v);
  } else if (name == "fnp_cuda_graph") {
    c.use_graph = parse_int(name, v);
    c.drop_graph();
  } else if (name == "fnp_timers") {
    c.timers_on = parse_int(name, v);
  } else {
    throw Error(FNP_ERR_OPTION, "unknown option '" + name + "'");
  }
}

// ---------------------------------------------------------------------------
// operators
// ---------------------------------------------------------------------------
static void op_shape(const Ctx &c, int which, int64_t &nrows, int64_t &ncols) {
  switch (which) {
    case FNP_MAT_A00: case FNP_MAT_P00: nrows = c.n_u; ncols = c.n_u_global; break;
    case FNP_MAT_A01: nrows = c.n_u; ncols = c.n_p_global; break;
    case FNP_MAT_A10: nrows = c.n_p; ncols = c.n_u_global; break;
    default: nrows = c.n_p; ncols = c.n_p_global; break;
  }
}

double comm_allreduce(Ctx &c, double v, bool max_op);

// Is the (sorted) pattern that of S (x) I_bs with interleaved components?  Rows bs*i+comp
// must have equal length and columns bs*j+comp for the same nodes j.
static bool kron_pattern(const HostCsr &h, int bs, int64_t col_shift_ok) {
  if (bs < 2 || h.nrows == 0 || h.nrows % bs != 0 || h.ncols % bs != 0 || !col_shift_ok) return false;
  const int64_t nn = h.nrows / bs;
  bool ok = true;
#pragma omp parallel for schedule(static) reduction(&& : ok)
  for (int64_t i = 0; i < nn; ++i) {
    const int32_t b0 = h.rowptr[bs * i], len = h.rowptr[bs * i + 1] - b0;
    for (int comp = 0; comp < bs && ok; ++comp) {
      const int32_t b = h.rowptr[bs * i + comp];
      if (h.rowptr[bs * i + comp + 1] - b != len) { ok = false; break; }
      for (int32_t k = 0; k < len; ++k)
        if (h.col[b + k] % bs != comp || h.col[b + k] / bs != h.col[b0 + k] / bs) { ok = false; break; }
    }
  }
  return ok;
}

static void set_pattern(Ctx &c, int which, const int32_t *rowptr, const int32_t *colidx) {
  FNP_REQUIRE(c.have_layout, FNP_ERR_STATE, "fnp_set_layout must precede fnp_set_pattern");
  FNP_REQUIRE(which >= 0 && which < FNP_MAT_COUNT, FNP_ERR_ARG, "bad operator id");
  FNP_REQUIRE(rowptr && (colidx || rowptr[0] == 0), FNP_ERR_ARG, "null pattern");
  int64_t nrows, ncols;
  op_shape(c, which, nrows, ncols);
  HostCsr &h = c.hmat[which];
  h.nrows = nrows;
  h.ncols = ncols;
  FNP_REQUIRE(rowptr[0] == 0, FNP_ERR_ARG, "rowptr[0] must be 0");
  h.rowptr.assign(rowptr, rowptr + nrows + 1);
  const int64_t nnz = h.rowptr[nrows];
  FNP_REQUIRE(nnz >= 0, FNP_ERR_ARG, "negative nnz");
  h.col.assign(colidx, colidx + nnz);
  h.val.clear();
  // sort every row by column, remember the permutation if anything moved
  std::vector<int64_t> &perm = c.perm[which];
  perm.clear();
  bool sorted = true;
  for (int64_t i = 0; i < nrows && sorted; ++i) {
    FNP_REQUIRE(h.rowptr[i + 1] >= h.rowptr[i], FNP_ERR_ARG, "rowptr not monotone");
    for (int32_t k = h.rowptr[i] + 1; k < h.rowptr[i + 1]; ++k)
      if (h.col[k - 1] > h.col[k]) { sorted = false; break; }
  }
  if (!sorted) {
    perm.resize(nnz);
    std::iota(perm.begin(), perm.end(), (int64_t)0);
    for (int64_t i = 0; i < nrows; ++i)
      std::sort(perm.begin() + h.rowptr[i], perm.begin() + h.rowptr[i + 1],
                [&](int64_t a, int64_t b) { return colidx[a] < colidx[b]; });
    for (int64_t k = 0; k < nnz; ++k) h.col[k] = colidx[perm[k]];
  }
  for (int64_t i = 0; i < nrows; ++i) {
    for (int32_t k = h.rowptr[i]; k < h.rowptr[i + 1]; ++k)
      FNP_REQUIRE(h.col[k] >= 0 && h.col[k] < ncols, FNP_ERR_ARG, "column index out of range");
  }
  static const char *names[FNP_MAT_COUNT] = {"A00", "A01", "A10", "Ap", "Mp", "Kp", "P00"};
  DevCsr &d = c.dmat[which];
  // Kronecker detection for the velocity blocks (Picard/Oseen: the same scalar operator
  // for every component).  Every rank must take the same decision.
  int bs = 1;
  if (c.kron && (which == FNP_MAT_A00 || which == FNP_MAT_P00)) {
    for (int cand : {3, 2}) {
      const bool aligned = c.u_begin % cand == 0 && c.n_u % cand == 0 && c.n_u_global % cand == 0;
      double ok = kron_pattern(h, cand, aligned) ? 1.0 : 0.0;
      ok = -comm_allreduce(c, -ok, true);          // min over ranks
      if (ok > 0.5) { bs = cand; break; }
    }
  }
  c.kron_bs[which] = bs;
  c.kron_rowptr[which].clear();
  if (bs > 1) {
    // keep the scalar pattern (component 0 rows, node columns); remember the user's row
    // pointers to pull and verify the values on every refresh
    c.kron_rowptr[which] = h.rowptr;
    HostCsr hs;
    hs.nrows = h.nrows / bs;
    hs.ncols = h.ncols / bs;
    hs.rowptr.resize(hs.nrows + 1);
    hs.rowptr[0] = 0;
    for (int64_t i = 0; i < hs.nrows; ++i) hs.rowptr[i + 1] = hs.rowptr[i] + (h.rowptr[bs * i + 1] - h.rowptr[bs * i]);
    hs.col.resize(hs.rowptr[hs.nrows]);
    for (int64_t i = 0; i < hs.nrows; ++i)
      for (int32_t k = 0; k < hs.rowptr[i + 1] - hs.rowptr[i]; ++k) hs.col[hs.rowptr[i] + k] = h.col[h.rowptr[bs * i] + k] / bs;
    h = std::move(hs);
  }
  if (c.nranks == 1) {
    csr_upload_pattern(c, d, h, names[which], -1, bs);
  } else {
    // multi-rank: the device copy uses local column numbering [owned | ghost]; the host
    // copy keeps the global ids (the AMG set-up starts from them)
    const bool u_cols = which == FNP_MAT_A00 || which == FNP_MAT_P00 || which == FNP_MAT_A10;
    std::vector<int64_t> begins = u_cols ? c.u_begins : c.p_begins;
    for (auto &b : begins) b /= bs;               // scalar (node) ownership in Kronecker mode
    HostCsr loc = h;
    std::shared_ptr<HaloPlan> plan = build_halo(c, loc, begins, nullptr);
    const int64_t n_own_cols = begins[c.rank + 1] - begins[c.rank];
    csr_upload_pattern(c, d, loc, names[which], (c.overlap || c.p2p) ? n_own_cols : -1, bs, n_own_cols);
    d.halo = (plan && bs > 1) ? expand_plan(c, *plan, bs) : plan;
    d.ncols_own = (int32_t)n_own_cols;
    d.nghost = plan ? plan->nghost : 0;
    c.local_cols[which] = std::move(loc.col);
  }
  c.have_pattern[which] = true;
  c.have_values[which] = false;
}

static bool keeps_host_values(int which) {
  return which == FNP_MAT_A00 || which == FNP_MAT_P00 || which == FNP_MAT_AP || which == FNP_MAT_A01;   // A01: Rp of PCDR
}

static void set_values(Ctx &c, int which, const double *values) {
  FNP_REQUIRE(which >= 0 && which < FNP_MAT_COUNT, FNP_ERR_ARG, "bad operator id");
  FNP_REQUIRE(c.have_pattern[which], FNP_ERR_STATE, "fnp_set_values before fnp_set_pattern");
  FNP_REQUIRE(values != nullptr || c.dmat[which].nnz == 0, FNP_ERR_ARG, "null values");
  HostCsr &h = c.hmat[which];
  const int bs = c.kron_bs[which];
  const int64_t nnz_user = bs > 1 ? (int64_t)c.kron_rowptr[which].back() : h.nnz();
  const std::vector<int64_t> &perm = c.perm[which];
  const double *src = values;
  std::vector<double> tmp, tmps;
  if (!perm.empty()) {
    tmp.resize(nnz_user);
    for (int64_t k = 0; k < nnz_user; ++k) tmp[k] = values[perm[k]];
    src = tmp.data();
  }
  if (bs > 1) {
    // scalar values = component 0; the other components must carry the same numbers
    const std::vector<int32_t> &rp = c.kron_rowptr[which];
    tmps.resize(h.nnz());
    bool ok = true;
#pragma omp parallel for schedule(static) reduction(&& : ok)
    for (int64_t i = 0; i < h.nrows; ++i) {
      const int32_t len = h.rowptr[i + 1] - h.rowptr[i];
      for (int32_t k = 0; k < len; ++k) {
        const double v = src[rp[bs * i] + k];
        tmps[h.rowptr[i] + k] = v;
        for (int comp = 1; comp < bs; ++comp) {
          const double w = src[rp[bs * i + comp] + k];
          if (std::fabs(w - v) > 1e-13 * (std::fabs(v) + std::fabs(w))) ok = false;
        }
      }
    }
    FNP_REQUIRE(ok, FNP_ERR_STATE, "the velocity block has the pattern of S (x) I but its values differ between "
                                   "components (Newton coupling / component-wise BCs?): set option fnp_kronecker 0 "
                                   "before fnp_set_pattern");
    src = tmps.data();
  }
  const int64_t nnz = h.nnz();
  if (keeps_host_values(which)) {
    h.val.assign(src, src + nnz);
    src = h.val.data();
  }
  const bool want_dinv = which == FNP_MAT_MP || which == FNP_MAT_AP || which == FNP_MAT_A00 || which == FNP_MAT_P00;
  if (c.nranks == 1) {
    csr_set_values(c, c.dmat[which], h, src, want_dinv);
  } else {
    HostCsr view;                       // same rows, local column numbering (diagonal = row index)
    view.nrows = h.nrows;
    view.ncols = h.ncols;
    view.rowptr = h.rowptr;
    view.col = c.local_cols[which];
    csr_set_values(c, c.dmat[which], view, src, want_dinv);
  }
  c.have_values[which] = true;
  c.dirty[which] = true;
}

// opt-in internal numbering of the velocity dofs (Ctx::reorder) ------------------
static bool u_rows(int which) { return which == FNP_MAT_A00 || which == FNP_MAT_P00 || which == FNP_MAT_A01; }
static bool u_cols(int which) { return which == FNP_MAT_A00 || which == FNP_MAT_P00 || which == FNP_MAT_A10; }

// Permuted copy of a user pattern: internal row i = user row u_perm[i] (velocity rows), columns
// mapped through u_inv (velocity columns); c.rmap[which] maps internal entries to user entries.
// FNP_MAT_A00 defines the permutation: dofs sorted by row length inside windows of c.reorder dofs
// (stable, so the components of a node -- equal row lengths, adjacent -- stay interleaved).
static void reorder_pattern(Ctx &c, int which, const int32_t *rowptr, const int32_t *colidx, std::vector<int32_t> &rp,
                            std::vector<int32_t> &ci) {
  c.rmap[which].clear();
  if (!u_rows(which) && !u_cols(which)) return;
  FNP_REQUIRE(c.nranks == 1, FNP_ERR_OPTION, "fnp_reorder_nodes is available on single-rank contexts only");
  if (which == FNP_MAT_A00) {
    FNP_REQUIRE(!c.have_pattern[FNP_MAT_A01] && !c.have_pattern[FNP_MAT_A10] && !c.have_pattern[FNP_MAT_P00] && c.mu_diag.empty() &&
                    !c.have_is,
                FNP_ERR_STATE, "with fnp_reorder_nodes, FNP_MAT_A00 must be set before the other velocity data");
    const int64_t n = c.n_u;
    c.u_perm.resize(n);
    std::iota(c.u_perm.begin(), c.u_perm.end(), (int64_t)0);
    for (int64_t w0 = 0; w0 < n; w0 += c.reorder) {
      const int64_t w1 = std::min<int64_t>(n, w0 + c.reorder);
      std::stable_sort(c.u_perm.begin() + w0, c.u_perm.begin() + w1, [&](int64_t a, int64_t b) {
        return rowptr[a + 1] - rowptr[a] > rowptr[b + 1] - rowptr[b];
      });
    }
    c.u_inv.resize(n);
    for (int64_t i = 0; i < n; ++i) c.u_inv[c.u_perm[i]] = i;
    c.d_u_perm.upload(c.u_perm.data(), (size_t)n, c.stream);
    FNP_CUDA(cudaStreamSynchronize(c.stream));
  }
  FNP_REQUIRE(c.reordered(), FNP_ERR_STATE, "with fnp_reorder_nodes, FNP_MAT_A00 must be set before the other velocity data");
  int64_t nrows, ncols;
  op_shape(c, which, nrows, ncols);
  FNP_REQUIRE(rowptr[0] == 0, FNP_ERR_ARG, "rowptr[0] must be 0");
  for (int64_t i = 0; i < nrows; ++i) FNP_REQUIRE(rowptr[i + 1] >= rowptr[i], FNP_ERR_ARG, "rowptr not monotone");
  const int64_t nnz = rowptr[nrows];
  rp.assign(nrows + 1, 0);
  ci.resize(nnz);
  std::vector<int64_t> &map = c.rmap[which];
  map.resize(nnz);
  const bool pr = u_rows(which), pc = u_cols(which);
  for (int64_t i = 0; i < nrows; ++i) {
    const int64_t r = pr ? c.u_perm[i] : i;
    rp[i + 1] = rp[i] + (rowptr[r + 1] - rowptr[r]);
  }
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nrows; ++i) {
    const int64_t r = pr ? c.u_perm[i] : i;
    for (int32_t k = rowptr[r], q = rp[i]; k < rowptr[r + 1]; ++k, ++q) {
      int32_t col = colidx[k];
      if (pc && col >= 0 && col < ncols) col = (int32_t)c.u_inv[col];   // out-of-range ids are reported by set_pattern
      ci[q] = col;
      map[q] = k;
    }
  }
}

// user velocity vector (device) -> internal numbering, and back
static const double *u_to_internal(Ctx &c, const double *user, DevBuf<double> &buf) {
  if (!c.reordered()) return user;
  buf.ensure((size_t)c.n_u);
  vec_gather(c, c.n_u, c.d_u_perm.p, user, buf.p);
  return buf.p;
}
static double *u_internal_out(Ctx &c, double *user) {
  if (!c.reordered()) return user;
  c.ro_out.ensure((size_t)c.n_u);
  return c.ro_out.p;
}
static void u_to_user(Ctx &c, const double *internal, double *user) {
  if (c.reordered()) vec_scatter(c, c.n_u, c.d_u_perm.p, internal, user);
}

// staging of host-pointer calls ---------------------------------------------
struct Staged {
  Ctx &c;
  const double *dev_in[3] = {nullptr, nullptr, nullptr};
  double *dev_out[2] = {nullptr, nullptr};
  double *host_out[2] = {nullptr, nullptr};
  int64_t out_n[2] = {0, 0};
  bool on_device;
  int nin = 0, nout = 0;
  Staged(Ctx &ctx, bool dev) : c(ctx), on_device(dev) {}
  const double *in(const double *p, int64_t n) {
    if (on_device) return p;
    FNP_REQUIRE(nin < 2, FNP_ERR_ARG, "too many staged inputs");
    DevBuf<double> &b = c.io[nin++];
    b.ensure((size_t)n);
    FNP_CUDA(cudaMemcpyAsync(b.p, p, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    return b.p;
  }
  double *out(double *p, int64_t n) {
    if (on_device) return p;
    DevBuf<double> &b = c.io[2 + nout];
    b.ensure((size_t)n);
    host_out[nout] = p;
    dev_out[nout] = b.p;
    out_n[nout] = n;
    return dev_out[nout++];
  }
  void finish() {
    if (on_device) return;
    for (int i = 0; i < nout; ++i)
      FNP_CUDA(cudaMemcpyAsync(host_out[i], dev_out[i], out_n[i] * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    FNP_CUDA(cudaStreamSynchronize(c.stream));
  }
};

}  // namespace fnp

using namespace fnp;

struct fnp_context {
  Ctx c;
  explicit fnp_context(int dev) : c(dev) {}
};

#define FNP_API_BEGIN try {
#define FNP_API_END                                      \
  return FNP_OK;                                         \
  }                                                      \
  catch (const fnp::Error &e) {                          \
    fnp::set_last_error(e.what());                       \
    return e.code;                                       \
  }                                                      \
  catch (const std::exception &e) {                      \
    fnp::set_last_error(std::string("internal: ") + e.what()); \
    return FNP_ERR_ARG;                                  \
  }

#define CTX(ctx)                                                         \
  FNP_REQUIRE((ctx) != nullptr, FNP_ERR_ARG, "null context");           \
  Ctx &c = (ctx)->c;                                                     \
  FNP_CUDA(cudaSetDevice(c.device))

extern "C" {

const char *fnp_last_error(void) { return g_last_error.c_str(); }
const char *fnp_version(void) { return "libfenapack_cuda 0.1 (sm_100a)"; }

int fnp_create(fnp_context **out, int device) {
  FNP_API_BEGIN
  FNP_REQUIRE(out != nullptr, FNP_ERR_ARG, "null output pointer");
  *out = new fnp_context(device);
  FNP_API_END
}

int fnp_nccl_unique_id(void *out128) {
  FNP_API_BEGIN
  FNP_REQUIRE(out128 != nullptr, FNP_ERR_ARG, "null output pointer");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  FNP_NCCL(nccl().GetUniqueId(&id));
  std::memcpy(out128, &id, sizeof(id));
  FNP_API_END
}

int fnp_create_dist(fnp_context **out, int device, const void *nccl_id, int rank, int nranks) {
  FNP_API_BEGIN
  FNP_REQUIRE(out != nullptr && nccl_id != nullptr, FNP_ERR_ARG, "null pointer");
  FNP_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, FNP_ERR_ARG, "bad rank / nranks");
  std::unique_ptr<fnp_context> ctx(new fnp_context(device));
  ncclUniqueId id;
  std::memcpy(&id, nccl_id, sizeof(id));
  FNP_NCCL(nccl().CommInitRank(&ctx->c.comm, nranks, id, rank));
  ctx->c.rank = rank;
  ctx->c.nranks = nranks;
  if (nranks > 1) {
    // a second communicator for the overlapped halo exchanges: NCCL operations of ONE
    // communicator must not be in flight on two streams at once
    if (nccl().CommSplit) FNP_NCCL(nccl().CommSplit(ctx->c.comm, 0, rank, &ctx->c.comm_halo, nullptr));
    FNP_CUDA(cudaStreamCreateWithFlags(&ctx->c.comm_stream, cudaStreamNonBlocking));
    FNP_CUDA(cudaEventCreateWithFlags(&ctx->c.ev_x, cudaEventDisableTiming));
    FNP_CUDA(cudaEventCreateWithFlags(&ctx->c.ev_halo, cudaEventDisableTiming));
  }
  *out = ctx.release();
  FNP_API_END
}

int fnp_destroy(fnp_context *ctx) {
  FNP_API_BEGIN
  delete ctx;
  FNP_API_END
}

int fnp_set_stream(fnp_context *ctx, void *cuda_stream) {
  FNP_API_BEGIN
  CTX(ctx);
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  c.drop_graph();                       // the captured apply belongs to the old stream
  if (c.own_stream) cudaStreamDestroy(c.stream);
  c.stream = reinterpret_cast<cudaStream_t>(cuda_stream);
  c.own_stream = false;
  FNP_API_END
}

int fnp_synchronize(fnp_context *ctx) {
  FNP_API_BEGIN
  CTX(ctx);
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  FNP_API_END
}

int fnp_set_option(fnp_context *ctx, const char *name, const char *value) {
  FNP_API_BEGIN
  CTX(ctx);
  FNP_REQUIRE(name && value, FNP_ERR_ARG, "null option name/value");
  set_option(c, name, value);
  FNP_API_END
}

int fnp_set_layout(fnp_context *ctx, int64_t n_u_local, int64_t u_begin, int64_t n_u_global, int64_t n_p_local,
                   int64_t p_begin, int64_t n_p_global) {
  FNP_API_BEGIN
  CTX(ctx);
  FNP_REQUIRE(!c.have_layout, FNP_ERR_STATE, "layout already set (reinitialisation is not allowed, cf. field_split.py:60)");
  FNP_REQUIRE(n_u_local >= 0 && n_p_local >= 0 && u_begin >= 0 && p_begin >= 0, FNP_ERR_ARG, "negative size");
  FNP_REQUIRE(u_begin + n_u_local <= n_u_global && p_begin + n_p_local <= n_p_global, FNP_ERR_ARG,
              "ownership range exceeds the global size");
  FNP_REQUIRE(n_u_global < INT32_MAX && n_p_global < INT32_MAX, FNP_ERR_ARG, "global sizes must fit 32-bit indices");
  if (c.nranks == 1)
    FNP_REQUIRE(u_begin == 0 && p_begin == 0 && n_u_local == n_u_global && n_p_local == n_p_global, FNP_ERR_ARG,
                "single-rank context must own everything");
  c.n_u = n_u_local; c.u_begin = u_begin; c.n_u_global = n_u_global;
  c.n_p = n_p_local; c.p_begin = p_begin; c.n_p_global = n_p_global;
  c.u_begins = comm_ranges(c, n_u_local);
  c.p_begins = comm_ranges(c, n_p_local);
  FNP_REQUIRE(c.u_begins[c.rank] == u_begin && c.u_begins[c.nranks] == n_u_global &&
                  c.p_begins[c.rank] == p_begin && c.p_begins[c.nranks] == n_p_global,
              FNP_ERR_ARG, "ownership ranges of the ranks are not contiguous / do not add up to the global sizes");
  c.have_layout = true;
  FNP_API_END
}

int fnp_set_pattern(fnp_context *ctx, int which, const int32_t *rowptr, const int32_t *colidx) {
  FNP_API_BEGIN
  CTX(ctx);
  FNP_REQUIRE(c.have_layout, FNP_ERR_STATE, "fnp_set_layout must precede fnp_set_pattern");
  FNP_REQUIRE(which >= 0 && which < FNP_MAT_COUNT, FNP_ERR_ARG, "bad operator id");
  FNP_REQUIRE(rowptr != nullptr, FNP_ERR_ARG, "null pattern");
  std::vector<int32_t> re_rp, re_ci;
  if (c.reorder > 0) {
    reorder_pattern(c, which, rowptr, colidx, re_rp, re_ci);
    if (!re_rp.empty()) {
      rowptr = re_rp.data();
      colidx = re_ci.data();
    }
  } else {
    c.rmap[which].clear();
  }
  c.user_rowptr[which].clear();
  c.user_col[which].clear();
  c.pattern_pending[which] = false;
  if (c.prune && (which == FNP_MAT_A00 || which == FNP_MAT_P00)) {
    // DOLFIN stores the velocity block with the dense per-cell coupling of all components,
    // explicit zeros included: the pattern is finalised at the first
    // fnp_set_values, when the stored zeros are known and can be dropped
    int64_t nrows, ncols;
    op_shape(c, which, nrows, ncols);
    FNP_REQUIRE(rowptr[0] == 0, FNP_ERR_ARG, "rowptr[0] must be 0");
    for (int64_t i = 0; i < nrows; ++i) FNP_REQUIRE(rowptr[i + 1] >= rowptr[i], FNP_ERR_ARG, "rowptr not monotone");
    c.user_rowptr[which].assign(rowptr, rowptr + nrows + 1);
    c.user_col[which].assign(colidx, colidx + rowptr[nrows]);
    c.user_nnz[which] = rowptr[nrows];
    c.prune_mask[which].clear();
    c.pattern_pending[which] = true;
    c.have_pattern[which] = true;
    c.have_values[which] = false;
  } else {
    set_pattern(c, which, rowptr, colidx);
  }
  FNP_API_END
}

int fnp_set_values(fnp_context *ctx, int which, const double *values) {
  FNP_API_BEGIN
  CTX(ctx);
  FNP_REQUIRE(which >= 0 && which < FNP_MAT_COUNT, FNP_ERR_ARG, "bad operator id");
  FNP_REQUIRE(c.have_pattern[which], FNP_ERR_STATE, "fnp_set_values before fnp_set_pattern");
  std::vector<double> re_vals;
  if (!c.rmap[which].empty()) {
    FNP_REQUIRE(values != nullptr, FNP_ERR_ARG, "null values");
    const std::vector<int64_t> &map = c.rmap[which];
    re_vals.resize(map.size());
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < (int64_t)map.size(); ++k) re_vals[k] = values[map[k]];
    values = re_vals.data();
  }
  const bool pruning = !c.user_rowptr[which].empty();
  if (pruning) {
    FNP_REQUIRE(values != nullptr, FNP_ERR_ARG, "null values");
    const std::vector<int32_t> &rp = c.user_rowptr[which], &ci = c.user_col[which];
    const int64_t nrows = (int64_t)rp.size() - 1, nnz = c.user_nnz[which];
    const int64_t row0 = c.u_begin;                       // A00 / P00: square in the u numbering
    std::vector<char> &keepmask = c.prune_mask[which];
    // (re)build the pattern when it is still pending, or when an entry dropped earlier as a
    // stored zero carries a value now (e.g. the convection term after a zero initial guess);
    // the decision is collective because the pattern set-up is
    double rebuild = c.pattern_pending[which] ? 1.0 : 0.0;
    if (!c.pattern_pending[which])
      for (int64_t k = 0; k < nnz; ++k)
        if (!keepmask[k] && values[k] != 0.0) { rebuild = 1.0; break; }
    rebuild = comm_allreduce(c, rebuild, true);
    if (rebuild > 0.5) {
      if (keepmask.size() != (size_t)nnz) keepmask.assign((size_t)nnz, 0);
      std::vector<int32_t> prp(nrows + 1, 0), pci;
      pci.reserve(nnz);
      for (int64_t i = 0; i < nrows; ++i) {
        for (int32_t k = rp[i]; k < rp[i + 1]; ++k) {
          if (values[k] != 0.0 || ci[k] == row0 + i) keepmask[k] = 1;
          if (keepmask[k]) pci.push_back(ci[k]);
        }
        prp[i + 1] = (int32_t)pci.size();
      }
      set_pattern(c, which, prp.data(), pci.data());
      c.pattern_pending[which] = false;
      c.is_setup = false;                                 // hierarchies / work space follow the new pattern
    }
    std::vector<double> packed;
    packed.reserve(c.hmat[which].nnz() * (size_t)c.kron_bs[which]);
    for (int64_t k = 0; k < nnz; ++k)
      if (keepmask[k]) packed.push_back(values[k]);
    set_values(c, which, packed.data());
  } else {
    set_values(c, which, values);
  }
  FNP_API_END
}

int fnp_set_bc(fnp_context *ctx, const int32_t *idx_local, const double *values, int32_t n) {
  FNP_API_BEGIN
  CTX(ctx);
  FNP_REQUIRE(c.have_layout, FNP_ERR_STATE, "fnp_set_layout must precede fnp_set_bc");
  FNP_REQUIRE(n >= 0 && (n == 0 || (idx_local && values)), FNP_ERR_ARG, "bad BC arrays");
  for (int32_t i = 0; i < n; ++i)
    FNP_REQUIRE(idx_local[i] >= 0 && idx_local[i] < c.n_p, FNP_ERR_ARG, "BC index outside the local pressure range");
  c.nbc = n;
  c.bc_idx.upload(idx_local, (size_t)n, c.stream);
  c.bc_val.upload(values, (size_t)n, c.stream);
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  FNP_API_END
}

int fnp_set_mu_diag(fnp_context *ctx, const double *diag_local) {
  FNP_API_BEGIN
  CTX(ctx);
  FNP_REQUIRE(c.have_layout, FNP_ERR_STATE, "fnp_set_layout must precede fnp_set_mu_diag");
  FNP_REQUIRE(diag_local != nullptr || c.n_u == 0, FNP_ERR_ARG, "null diagonal");
  c.mu_diag.assign(diag_local, diag_local + c.n_u);
  if (c.reorder > 0) {
    FNP_REQUIRE(c.reordered(), FNP_ERR_STATE, "with fnp_reorder_nodes, FNP_MAT_A00 must be set before the other velocity data");
    for (int64_t i = 0; i < c.n_u; ++i) c.mu_diag[i] = diag_local[c.u_perm[i]];
  }
  c.mu_dirty = true;
  FNP_API_END
}

int fnp_rp_solve(fnp_context *ctx, const double *b, double *x, int on_device) {
  FNP_API_BEGIN
  CTX(ctx);
  FNP_REQUIRE(c.is_setup, FNP_ERR_STATE, "fnp_setup has not been called");
  FNP_REQUIRE(c.variant >= 3, FNP_ERR_STATE, "Rp exists for the PCDR variants only");
  Staged s(c, on_device != 0);
  const double *db = s.in(b, c.n_p);
  double *dx = s.out(x, c.n_p);
  rp_solve(c, db, dx);
  s.finish();
  FNP_API_END
}

int fnp_set_index_sets(fnp_context *ctx, const int64_t *is_u_local, const int64_t *is_p_local) {
  FNP_API_BEGIN
  CTX(ctx);
  FNP_REQUIRE(c.have_layout, FNP_ERR_STATE, "fnp_set_layout must precede fnp_set_index_sets");
  FNP_REQUIRE(is_u_local && is_p_local, FNP_ERR_ARG, "null index set");
  const int64_t n = c.n_u + c.n_p;
  std::vector<char> seen((size_t)n, 0);
  auto check = [&](const int64_t *is, int64_t m) {
    for (int64_t i = 0; i < m; ++i) {
      FNP_REQUIRE(is[i] >= 0 && is[i] < n && !seen[(size_t)is[i]], FNP_ERR_ARG, "index sets are not a partition of the local vector");
      seen[(size_t)is[i]] = 1;
    }
  };
  check(is_u_local, c.n_u);
  check(is_p_local, c.n_p);
  c.is_u.upload(is_u_local, (size_t)c.n_u, c.stream);
  c.is_p.upload(is_p_local, (size_t)c.n_p, c.stream);
  if (c.reorder > 0) {
    // internal velocity dof i sits at monolithic position is_u[u_perm[i]]
    FNP_REQUIRE(c.reordered(), FNP_ERR_STATE, "with fnp_reorder_nodes, FNP_MAT_A00 must be set before the other velocity data");
    std::vector<int64_t> re((size_t)c.n_u);
    for (int64_t i = 0; i < c.n_u; ++i) re[(size_t)i] = is_u_local[c.u_perm[i]];
    c.d_is_u_re.upload(re.data(), re.size(), c.stream);
  }
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  c.have_is = true;
  FNP_API_END
}

int fnp_setup(fnp_context *ctx) {
  FNP_API_BEGIN
  CTX(ctx);
  setup_all(c);
  FNP_API_END
}

int fnp_spmv(fnp_context *ctx, int which, const double *x, double *y, int on_device) {
  FNP_API_BEGIN
  CTX(ctx);
  FNP_REQUIRE(which == FNP_MAT_RP ? c.amg_rp.built || c.rp.nnz > 0 : (which >= 0 && which < FNP_MAT_COUNT && c.have_values[which]),
              FNP_ERR_STATE, "operator has no values");
  const DevCsr &A = which == FNP_MAT_RP ? c.rp : c.dmat[which];
  Staged s(c, on_device != 0);
  const double *dx = s.in(x, A.vec_cols());
  double *dy = s.out(y, A.vec_rows());
  const bool rx = c.reordered() && which != FNP_MAT_RP && u_cols(which), ry = c.reordered() && which != FNP_MAT_RP && u_rows(which);
  const double *ix = rx ? u_to_internal(c, dx, c.ro_in) : dx;
  double *iy = ry ? u_internal_out(c, dy) : dy;
  spmv_store(c, A, ix, iy);
  if (ry) u_to_user(c, iy, dy);
  s.finish();
  FNP_API_END
}

#define REQUIRE_SETUP() FNP_REQUIRE(c.is_setup, FNP_ERR_STATE, "fnp_setup has not been called")

// a peer-memory flag wait that timed out (a neighbour rank died or fell out of step)
static void check_p2p(Ctx &c) {
  if (c.nranks == 1 || !c.p2p_err.p) return;
  int e = 0;
  FNP_CUDA(cudaMemcpyAsync(&e, c.p2p_err.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  FNP_REQUIRE(e == 0, FNP_ERR_NCCL, "peer-memory halo exchange timed out waiting for a neighbour rank");
}

int fnp_mp_solve(fnp_context *ctx, const double *b, double *x, int on_device) {
  FNP_API_BEGIN
  CTX(ctx);
  REQUIRE_SETUP();
  Staged s(c, on_device != 0);
  const double *db = s.in(b, c.n_p);
  double *dx = s.out(x, c.n_p);
  mp_solve(c, db, 1.0, nullptr, dx);
  s.finish();
  FNP_API_END
}

int fnp_ap_solve(fnp_context *ctx, const double *b, double *x, int on_device) {
  FNP_API_BEGIN
  CTX(ctx);
  REQUIRE_SETUP();
  Staged s(c, on_device != 0);
  const double *db = s.in(b, c.n_p);
  double *dx = s.out(x, c.n_p);
  ap_solve(c, db, dx);
  s.finish();
  FNP_API_END
}

int fnp_u_solve(fnp_context *ctx, const double *b, double *x, int on_device) {
  FNP_API_BEGIN
  CTX(ctx);
  REQUIRE_SETUP();
  Staged s(c, on_device != 0);
  const double *db = s.in(b, c.n_u);
  double *dx = s.out(x, c.n_u);
  double *ix = u_internal_out(c, dx);
  u_solve(c, u_to_internal(c, db, c.ro_in), ix);
  u_to_user(c, ix, dx);
  s.finish();
  FNP_API_END
}

int fnp_schur_apply(fnp_context *ctx, const double *x_p, double *y_p, int on_device) {
  FNP_API_BEGIN
  CTX(ctx);
  REQUIRE_SETUP();
  FNP_REQUIRE(x_p != y_p, FNP_ERR_ARG, "x and y must not alias (PETSc PCApply contract)");
  Staged s(c, on_device != 0);
  const double *dx = s.in(x_p, c.n_p);
  double *dy = s.out(y_p, c.n_p);
  schur_apply(c, dx, dy);
  s.finish();
  FNP_API_END
}

int fnp_pc_apply(fnp_context *ctx, const double *x_u, const double *x_p, double *y_u, double *y_p, int on_device) {
  FNP_API_BEGIN
  CTX(ctx);
  REQUIRE_SETUP();
  Staged s(c, on_device != 0);
  const double *dxu = s.in(x_u, c.n_u);
  const double *dxp = s.in(x_p, c.n_p);
  double *dyu = s.out(y_u, c.n_u);
  double *dyp = s.out(y_p, c.n_p);
  double *iyu = u_internal_out(c, dyu);
  pc_apply(c, u_to_internal(c, dxu, c.ro_in), dxp, iyu, dyp);
  u_to_user(c, iyu, dyu);
  s.finish();
  check_p2p(c);
  FNP_API_END
}

int fnp_solve(fnp_context *ctx, const double *b_u, const double *b_p, double *x_u, double *x_p, int on_device,
              int32_t *iterations, double *residual_norm, int32_t *pc_applies) {
  FNP_API_BEGIN
  CTX(ctx);
  REQUIRE_SETUP();
  FNP_REQUIRE(b_u && b_p && x_u && x_p, FNP_ERR_ARG, "null vector");
  const int64_t n = c.n_u + c.n_p;
  // persistent staging of b and x in split layout [u;p] (no allocation per solve)
  DevBuf<double> &xs = c.sol_x, &bs = c.sol_b;
  xs.ensure((size_t)n);
  bs.ensure((size_t)n);
  const cudaMemcpyKind in_kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  const cudaMemcpyKind out_kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (c.reordered()) {
    c.ro_in2.ensure((size_t)c.n_u);
    FNP_CUDA(cudaMemcpyAsync(c.ro_in2.p, b_u, c.n_u * sizeof(double), in_kind, c.stream));
    vec_gather(c, c.n_u, c.d_u_perm.p, c.ro_in2.p, bs.p);
  } else {
    FNP_CUDA(cudaMemcpyAsync(bs.p, b_u, c.n_u * sizeof(double), in_kind, c.stream));
  }
  FNP_CUDA(cudaMemcpyAsync(bs.p + c.n_u, b_p, c.n_p * sizeof(double), in_kind, c.stream));
  int32_t its = 0, nap = 0;
  double rn = 0.0;
  solve_fgmres(c, bs.p, xs.p, &its, &rn, &nap);
  const double *xu_src = xs.p;
  if (c.reordered()) {
    c.ro_in2.ensure((size_t)c.n_u);
    vec_scatter(c, c.n_u, c.d_u_perm.p, xs.p, c.ro_in2.p);
    xu_src = c.ro_in2.p;
  }
  FNP_CUDA(cudaMemcpyAsync(x_u, xu_src, c.n_u * sizeof(double), out_kind, c.stream));
  FNP_CUDA(cudaMemcpyAsync(x_p, xs.p + c.n_u, c.n_p * sizeof(double), out_kind, c.stream));
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  if (iterations) *iterations = its;
  if (residual_norm) *residual_norm = rn;
  if (pc_applies) *pc_applies = nap;
  check_p2p(c);
  FNP_API_END
}

int fnp_solve_monolithic(fnp_context *ctx, const double *b, double *x, int on_device, int32_t *iterations,
                         double *residual_norm, int32_t *pc_applies) {
  FNP_API_BEGIN
  CTX(ctx);
  REQUIRE_SETUP();
  FNP_REQUIRE(c.have_is, FNP_ERR_STATE, "fnp_solve_monolithic needs fnp_set_index_sets");
  const int64_t n = c.n_u + c.n_p;
  DevBuf<double> &mono = c.sol_m, &bs = c.sol_b, &xs = c.sol_x;
  mono.ensure((size_t)n); bs.ensure((size_t)n); xs.ensure((size_t)n);
  const double *dmono = b;
  if (!on_device) {
    FNP_CUDA(cudaMemcpyAsync(mono.p, b, n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    dmono = mono.p;
  }
  const int64_t *isu = c.reordered() ? c.d_is_u_re.p : c.is_u.p;
  vec_gather(c, c.n_u, isu, dmono, bs.p);
  vec_gather(c, c.n_p, c.is_p.p, dmono, bs.p + c.n_u);
  int32_t its = 0, nap = 0;
  double rn = 0.0;
  solve_fgmres(c, bs.p, xs.p, &its, &rn, &nap);
  double *dout = on_device ? x : mono.p;
  vec_scatter(c, c.n_u, isu, xs.p, dout);
  vec_scatter(c, c.n_p, c.is_p.p, xs.p + c.n_u, dout);
  if (!on_device) FNP_CUDA(cudaMemcpyAsync(x, mono.p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  if (iterations) *iterations = its;
  if (residual_norm) *residual_norm = rn;
  if (pc_applies) *pc_applies = nap;
  FNP_API_END
}

int fnp_get_residual_history(fnp_context *ctx, double *out, int32_t capacity) {
  if (!ctx || !out) return FNP_ERR_ARG;
  const auto &h = ctx->c.res_hist;
  const int32_t n = std::min<int32_t>((int32_t)h.size(), capacity);
  for (int32_t i = 0; i < n; ++i) out[i] = h[i];
  return n;
}

// ---- introspection -------------------------------------------------------
static DevHierarchy &hier(Ctx &c, int which) {
  FNP_REQUIRE(which == FNP_MAT_AP || which == FNP_MAT_A00 || which == FNP_MAT_P00 || which == FNP_MAT_RP, FNP_ERR_ARG,
              "AMG hierarchies exist for FNP_MAT_AP, FNP_MAT_A00 and FNP_MAT_RP only");
  DevHierarchy &H = which == FNP_MAT_AP ? c.amg_ap : (which == FNP_MAT_RP ? c.amg_rp : c.amg_u);
  FNP_REQUIRE(H.built, FNP_ERR_STATE, "AMG hierarchy not built");
  return H;
}

static const HostCsr &hier_mat(DevHierarchy &H, int level, int kind) {
  FNP_REQUIRE(level >= 0 && level < (int)H.host.levels.size(), FNP_ERR_ARG, "bad AMG level");
  const HostLevel &L = H.host.levels[level];
  FNP_REQUIRE(kind >= 0 && kind <= 2, FNP_ERR_ARG, "bad kind");
  if (kind != 0) FNP_REQUIRE(level + 1 < (int)H.host.levels.size(), FNP_ERR_ARG, "coarsest level has no P/R");
  return kind == 0 ? L.A : (kind == 1 ? L.P : L.R);
}

int fnp_operator_block_size(fnp_context *ctx, int which, int32_t *bs) {
  FNP_API_BEGIN
  CTX(ctx);
  FNP_REQUIRE(((which >= 0 && which < FNP_MAT_COUNT) || which == FNP_MAT_RP) && bs, FNP_ERR_ARG, "bad argument");
  *bs = which == FNP_MAT_RP ? 1 : c.kron_bs[which];
  FNP_API_END
}

int fnp_amg_num_levels(fnp_context *ctx, int which, int32_t *levels) {
  FNP_API_BEGIN
  CTX(ctx);
  *levels = (int32_t)hier(c, which).host.levels.size();
  FNP_API_END
}

int fnp_amg_level_info(fnp_context *ctx, int which, int level, int kind, int64_t *nrows, int64_t *ncols, int64_t *nnz,
                       double *rho) {
  FNP_API_BEGIN
  CTX(ctx);
  DevHierarchy &H = hier(c, which);
  const HostCsr &M = hier_mat(H, level, kind);
  if (nrows) *nrows = M.nrows;
  if (ncols) *ncols = M.ncols;
  if (nnz) *nnz = M.nnz();
  if (rho) *rho = H.host.levels[level].rho;
  FNP_API_END
}

int fnp_amg_level_get(fnp_context *ctx, int which, int level, int kind, int32_t *rowptr, int32_t *colidx, double *values) {
  FNP_API_BEGIN
  CTX(ctx);
  const HostCsr &M = hier_mat(hier(c, which), level, kind);
  std::memcpy(rowptr, M.rowptr.data(), M.rowptr.size() * sizeof(int32_t));
  std::memcpy(colidx, M.col.data(), M.col.size() * sizeof(int32_t));
  std::memcpy(values, M.val.data(), M.val.size() * sizeof(double));
  FNP_API_END
}

int fnp_amg_coarse_inverse(fnp_context *ctx, int which, double *dense_row_major) {
  FNP_API_BEGIN
  CTX(ctx);
  DevHierarchy &H = hier(c, which);
  std::memcpy(dense_row_major, H.host.coarse_inv.data(), H.host.coarse_inv.size() * sizeof(double));
  FNP_API_END
}

int fnp_amg_vcycle(fnp_context *ctx, int which, const double *b, double *x, int on_device) {
  FNP_API_BEGIN
  CTX(ctx);
  DevHierarchy &H = hier(c, which);
  const int64_t n = H.levels[0].A().vec_rows();
  Staged s(c, on_device != 0);
  const double *db = s.in(b, n);
  double *dx = s.out(x, n);
  const bool ru = c.reordered() && (which == FNP_MAT_A00 || which == FNP_MAT_P00);
  double *ix = ru ? u_internal_out(c, dx) : dx;
  amg_vcycle(c, H, ru ? u_to_internal(c, db, c.ro_in) : db, ix);
  if (ru) u_to_user(c, ix, dx);
  s.finish();
  FNP_API_END
}

int fnp_get_timer(fnp_context *ctx, const char *name, double *ms, int64_t *calls) {
  FNP_API_BEGIN
  CTX(ctx);
  c.resolve_timers();
  auto it = c.timers.find(name ? name : "");
  if (ms) *ms = it == c.timers.end() ? 0.0 : it->second.ms;
  if (calls) *calls = it == c.timers.end() ? 0 : it->second.calls;
  FNP_API_END
}

int fnp_reset_timers(fnp_context *ctx) {
  FNP_API_BEGIN
  CTX(ctx);
  c.resolve_timers();
  c.timers.clear();
  FNP_API_END
}

int64_t fnp_kernel_launches(fnp_context *ctx) { return ctx ? ctx->c.launches : -1; }

int fnp_event_tic(fnp_context *ctx) {
  FNP_API_BEGIN
  CTX(ctx);
  FNP_CUDA(cudaEventRecord(c.ev_tic, c.stream));
  FNP_API_END
}

int fnp_event_toc(fnp_context *ctx, double *elapsed_ms) {
  FNP_API_BEGIN
  CTX(ctx);
  FNP_CUDA(cudaEventRecord(c.ev_toc, c.stream));
  FNP_CUDA(cudaEventSynchronize(c.ev_toc));
  float ms = 0.f;
  FNP_CUDA(cudaEventElapsedTime(&ms, c.ev_tic, c.ev_toc));
  if (elapsed_ms) *elapsed_ms = ms;
  FNP_API_END
}

}  // extern "C"
