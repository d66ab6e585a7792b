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
  for (auto e : ev_iter) cudaEventDestroy(e);
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
      t.bytes += p.bytes;
    }
    event_pool.push_back(p.a);
    event_pool.push_back(p.b);
  }
  pending.clear();
}

StageTimer::StageTimer(Ctx &ctx, const char *nm, int level, double nbytes) : c(ctx), bytes(nbytes) {
  if (c.timers_on < level) return;
  name = nm;
  a = c.get_event();
  cudaEventRecord(a, c.stream);
}
StageTimer::StageTimer(Ctx &ctx, const std::string &nm, int level, double nbytes) : c(ctx), bytes(nbytes) {
  if (c.timers_on < level) return;
  name = nm;
  a = c.get_event();
  cudaEventRecord(a, c.stream);
}
StageTimer::~StageTimer() {
  if (!a) return;
  cudaEvent_t b = c.get_event();
  cudaEventRecord(b, c.stream);
  c.pending.push_back({name, a, b, bytes});
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
  } else if (name == "fnp_sell_gather") {
    c.sell_gather = (int)parse_int(name, v);
    c.drop_graph();
  } else if (name == "fnp_refresh_chunk_terms") {
    c.refresh_chunk_terms = std::max<int64_t>(1, (int64_t)parse_real(name, v));
  } else if (name == "fnp_sell_warps_rows") {
    c.sell_warps_rows = (int64_t)parse_real(name, v);
  } else if (name == "fnp_halo_split_rows") {
    c.halo_split_rows = (int64_t)parse_real(name, v);
  } else if (name == "fnp_gmres_sync") {
    c.gmres_sync = parse_int(name, v);
  } else if (name == "fnp_sell_warps") {
    c.sell_warps = (int)parse_int(name, v);
    FNP_REQUIRE(c.sell_warps == 0 || c.sell_warps == 1 || c.sell_warps == 2 || c.sell_warps == 4 || c.sell_warps == 8,
                FNP_ERR_OPTION, "fnp_sell_warps: 0 (auto), 1, 2, 4 or 8 (takes effect at fnp_set_pattern / fnp_setup)");
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
  ingest_set_pattern(c, which, rowptr, colidx);
  FNP_API_END
}

int fnp_set_values(fnp_context *ctx, int which, const double *values) {
  FNP_API_BEGIN
  CTX(ctx);
  StageTimer t(c, "FENaPack: set_values");
  ingest_set_values(c, which, values);
  FNP_API_END
}

int fnp_set_bc(fnp_context *ctx, const int32_t *idx_local, const double *values, int32_t n) {
  FNP_API_BEGIN
  CTX(ctx);
  FNP_REQUIRE(c.have_layout, FNP_ERR_STATE, "fnp_set_layout must precede fnp_set_bc");
  FNP_REQUIRE(n >= 0 && (n == 0 || (idx_local && values)), FNP_ERR_ARG, "bad BC arrays");
  for (int32_t i = 0; i < n; ++i)
    FNP_REQUIRE(idx_local[i] >= 0 && idx_local[i] < c.n_p, FNP_ERR_ARG, "BC index outside the local pressure range");
  c.drop_graph();                       // the captured apply holds the BC buffers and their length
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
  spmv_store(c, A, dx, dy);
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
  u_solve(c, db, dx);
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
  pc_apply(c, dxu, dxp, dyu, dyp);
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
  FNP_CUDA(cudaMemcpyAsync(bs.p, b_u, c.n_u * sizeof(double), in_kind, c.stream));
  FNP_CUDA(cudaMemcpyAsync(bs.p + c.n_u, b_p, c.n_p * sizeof(double), in_kind, c.stream));
  int32_t its = 0, nap = 0;
  double rn = 0.0;
  solve_fgmres(c, bs.p, xs.p, &its, &rn, &nap);
  FNP_CUDA(cudaMemcpyAsync(x_u, xs.p, c.n_u * sizeof(double), out_kind, c.stream));
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
  const int64_t *isu = c.is_u.p;
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

int fnp_rp_info(fnp_context *ctx, int64_t *nrows_local, int64_t *nnz) {
  FNP_API_BEGIN
  CTX(ctx);
  FNP_REQUIRE(c.variant >= 3 && c.h_rp.nrows > 0, FNP_ERR_STATE, "Rp exists for the PCDR variants after fnp_setup only");
  if (nrows_local) *nrows_local = c.h_rp.nrows;
  if (nnz) *nnz = c.h_rp.nnz();
  FNP_API_END
}

int fnp_rp_get(fnp_context *ctx, int32_t *rowptr, int32_t *colidx_global, double *values) {
  FNP_API_BEGIN
  CTX(ctx);
  FNP_REQUIRE(c.variant >= 3 && c.h_rp.nrows > 0, FNP_ERR_STATE, "Rp exists for the PCDR variants after fnp_setup only");
  FNP_REQUIRE(rowptr && colidx_global && values, FNP_ERR_ARG, "null output");
  std::memcpy(rowptr, c.h_rp.rowptr.data(), c.h_rp.rowptr.size() * sizeof(int32_t));
  std::memcpy(colidx_global, c.h_rp.col.data(), c.h_rp.col.size() * sizeof(int32_t));
  std::memcpy(values, c.h_rp.val.data(), c.h_rp.val.size() * sizeof(double));
  FNP_API_END
}

int fnp_get_converged_reason(fnp_context *ctx, int32_t *reason) {
  if (!ctx || !reason) return FNP_ERR_ARG;
  *reason = ctx->c.converged_reason;
  return FNP_OK;
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
  amg_sync_host(c, H);                  // device-side refreshes leave the host mirror behind
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
  amg_vcycle(c, H, db, dx);
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

int fnp_get_timer_bytes(fnp_context *ctx, const char *name, double *bytes) {
  FNP_API_BEGIN
  CTX(ctx);
  c.resolve_timers();
  auto it = c.timers.find(name ? name : "");
  if (bytes) *bytes = it == c.timers.end() ? 0.0 : it->second.bytes;
  FNP_API_END
}

int fnp_timer_names(fnp_context *ctx, char *out, int64_t capacity) {
  if (!ctx) return FNP_ERR_ARG;
  try {
    ctx->c.resolve_timers();
  } catch (...) {
    return FNP_ERR_CUDA;
  }
  std::string all;
  for (const auto &kv : ctx->c.timers) {
    all += kv.first;
    all += '\n';
  }
  if (out && capacity > 0) {
    const size_t n = std::min<size_t>(all.size(), (size_t)capacity - 1);
    std::memcpy(out, all.data(), n);
    out[n] = '\0';
  }
  return (int)all.size();
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
