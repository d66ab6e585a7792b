// Operator ingestion: fnp_set_pattern / fnp_set_values.
//
// The pattern of an operator is analysed ONCE on the host (row sort, stored-zero pruning of the
// velocity blocks, Kronecker detection, SELL layout); what the analysis leaves behind is a map from
// every stored device entry back to its position in the caller's value array.  A value refresh --
// the per-Newton-step / per-time-step path of the reference, PCDInterface._assemble_operator_deep with
// MAT_REUSE_MATRIX (fenapack/field_split_backend.py:285-291, 331-334) driven by BasePCDPC.setUp
// (fenapack/preconditioners.py:71-85) -- is then device work only: the caller's array is copied to the
// device as it is (or used in place when it already is device memory) and ONE kernel scatters the
// values into the SELL / CSR storage, verifies the Kronecker structure, checks that the entries pruned
// as stored zeros are still zero and recomputes the Jacobi diagonal.  Host copies of the values are
// fetched lazily, only when a host-side hierarchy (re)build asks for them.
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#include "fnp_internal.cuh"

namespace fnp {

double comm_allreduce(Ctx &c, double v, bool max_op);
std::shared_ptr<HaloPlan> build_halo(Ctx &c, HostCsr &h, const std::vector<int64_t> &begins,
                                     std::vector<int64_t> *ghost_global_out);
std::shared_ptr<HaloPlan> expand_plan(Ctx &c, const HaloPlan &p, int bs);

static const char *kNames[FNP_MAT_COUNT] = {"A00", "A01", "A10", "Ap", "Mp", "Kp", "P00", "P01", "A11"};

static void op_shape(const Ctx &c, int which, int64_t &nrows, int64_t &ncols) {
  switch (which) {
    case FNP_MAT_A00: case FNP_MAT_P00: nrows = c.n_u; ncols = c.n_u_global; break;
    case FNP_MAT_A01: case FNP_MAT_P01: nrows = c.n_u; ncols = c.n_p_global; break;
    case FNP_MAT_A10: nrows = c.n_p; ncols = c.n_u_global; break;
    default: nrows = c.n_p; ncols = c.n_p_global; break;
  }
}

// Is the (sorted) pattern that of S (x) I_bs with interleaved components?  Rows bs*i+comp
// must have equal length and columns bs*j+comp for the same nodes j.
static bool kron_pattern(const HostCsr &h, int bs, bool aligned) {
  if (bs < 2 || h.nrows == 0 || h.nrows % bs != 0 || h.ncols % bs != 0 || !aligned) return false;
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

// exclusive scan of per-row counts into row pointers (parallel two-pass)
static void scan_rows(const std::vector<int32_t> &cnt, std::vector<int32_t> &rp) {
  const int64_t n = (int64_t)cnt.size();
  rp.assign(n + 1, 0);
  const int nt = std::max(1, omp_get_max_threads());
  std::vector<int64_t> part(nt + 1, 0);
#pragma omp parallel num_threads(nt)
  {
    const int t = omp_get_thread_num();
    const int64_t b = n * t / nt, e = n * (t + 1) / nt;
    int64_t s = 0;
    for (int64_t i = b; i < e; ++i) s += cnt[i];
    part[t + 1] = s;
#pragma omp barrier
#pragma omp single
    for (int q = 0; q < nt; ++q) part[q + 1] += part[q];
    int64_t o = part[t];
    for (int64_t i = b; i < e; ++i) {
      o += cnt[i];
      rp[i + 1] = (int32_t)o;
    }
  }
  FNP_REQUIRE(part[nt] < (int64_t)INT32_MAX, FNP_ERR_ARG, "operator exceeds 2^31 stored entries on one rank");
}

// Finalise the pattern of operator `which`.  (rowptr, colidx) is the caller's pattern; `keep`
// (optional, one byte per caller entry) selects the entries that are stored.  Leaves behind:
// hmat[which] (sorted pattern, scalar in Kronecker mode), the device operator, and the device maps
// of the value path.
static void finalize_pattern(Ctx &c, int which, const int32_t *rowptr, const int32_t *colidx, const char *keep) {
  FNP_REQUIRE(c.have_layout, FNP_ERR_STATE, "fnp_set_layout must precede fnp_set_pattern");
  FNP_REQUIRE(which >= 0 && which < FNP_MAT_COUNT, FNP_ERR_ARG, "bad operator id");
  int64_t nrows, ncols;
  op_shape(c, which, nrows, ncols);
  FNP_REQUIRE(rowptr && rowptr[0] == 0, FNP_ERR_ARG, "rowptr[0] must be 0");
  const int64_t nnz_user = rowptr[nrows];
  FNP_REQUIRE(nnz_user >= 0, FNP_ERR_ARG, "negative nnz");
  FNP_REQUIRE(colidx || nnz_user == 0, FNP_ERR_ARG, "null pattern");
  {
    bool mono = true, inrange = true;
#pragma omp parallel for schedule(static) reduction(&& : mono, inrange)
    for (int64_t i = 0; i < nrows; ++i) {
      if (rowptr[i + 1] < rowptr[i]) mono = false;
      else
        for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
          if (colidx[k] < 0 || colidx[k] >= ncols) inrange = false;
    }
    FNP_REQUIRE(mono, FNP_ERR_ARG, "rowptr not monotone");
    FNP_REQUIRE(inrange, FNP_ERR_ARG, "column index out of range");
  }
  // ---- expanded pattern: kept entries, every row sorted by column; src = caller entry ----------
  HostCsr &h = c.hmat[which];
  h.nrows = nrows;
  h.ncols = ncols;
  h.val.clear();
  std::vector<int32_t> cnt((size_t)nrows);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nrows; ++i) {
    int32_t n = 0;
    if (keep) {
      for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) n += keep[k] ? 1 : 0;
    } else {
      n = rowptr[i + 1] - rowptr[i];
    }
    cnt[i] = n;
  }
  scan_rows(cnt, h.rowptr);
  const int64_t nnz = h.rowptr[nrows];
  h.col.resize((size_t)nnz);
  std::vector<int32_t> src((size_t)nnz);
  bool identity = nnz == nnz_user;
#pragma omp parallel
  {
    std::vector<std::pair<int32_t, int32_t>> row;
    bool ident_local = true;
#pragma omp for schedule(static)
    for (int64_t i = 0; i < nrows; ++i) {
      row.clear();
      for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k)
        if (!keep || keep[k]) row.push_back({colidx[k], k});
      bool sorted = true;
      for (size_t t = 1; t < row.size(); ++t)
        if (row[t - 1].first > row[t].first) { sorted = false; break; }
      if (!sorted) std::stable_sort(row.begin(), row.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
      int32_t o = h.rowptr[i];
      for (const auto &e : row) {
        h.col[o] = e.first;
        src[o] = e.second;
        if (e.second != o) ident_local = false;
        ++o;
      }
    }
#pragma omp critical
    identity = identity && ident_local;
  }
  // ---- Kronecker detection for the velocity blocks (Picard/Oseen: the same scalar operator
  // for every component).  Every rank must take the same decision. -----------------------------
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
  c.d_exp_rowptr[which].release();
  if (bs > 1) {
    // keep the scalar pattern (component 0 rows, node columns); the expanded row pointers stay on
    // the device: the value kernel pulls component 0 and verifies the others on every refresh
    c.kron_rowptr[which] = h.rowptr;
    c.d_exp_rowptr[which].upload(h.rowptr.data(), h.rowptr.size(), c.stream);
    HostCsr hs;
    hs.nrows = h.nrows / bs;
    hs.ncols = h.ncols / bs;
    std::vector<int32_t> scnt((size_t)hs.nrows);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < hs.nrows; ++i) scnt[i] = h.rowptr[bs * i + 1] - h.rowptr[bs * i];
    scan_rows(scnt, hs.rowptr);
    hs.col.resize((size_t)hs.rowptr[hs.nrows]);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < hs.nrows; ++i)
      for (int32_t k = 0; k < hs.rowptr[i + 1] - hs.rowptr[i]; ++k) hs.col[hs.rowptr[i] + k] = h.col[h.rowptr[bs * i] + k] / bs;
    FNP_CUDA(cudaStreamSynchronize(c.stream));      // h.rowptr is about to be replaced
    h = std::move(hs);
  }
  if (identity) c.d_map[which].release();
  else c.d_map[which].upload(src.data(), src.size(), c.stream);
  // entries dropped as stored zeros: re-checked on every refresh
  c.n_dropped[which] = 0;
  c.d_dropped[which].release();
  if (keep && nnz != nnz_user) {
    std::vector<int32_t> dropped;
    dropped.reserve((size_t)(nnz_user - nnz));
    for (int64_t k = 0; k < nnz_user; ++k)
      if (!keep[k]) dropped.push_back((int32_t)k);
    c.n_dropped[which] = (int64_t)dropped.size();
    c.d_dropped[which].upload(dropped.data(), dropped.size(), c.stream);
    FNP_CUDA(cudaStreamSynchronize(c.stream));
  }
  DevCsr &d = c.dmat[which];
  if (c.nranks == 1) {
    csr_upload_pattern(c, d, h, kNames[which], -1, bs);
    c.local_cols[which].clear();
  } else {
    // multi-rank: the device copy uses local column numbering [owned | ghost]; the host
    // copy keeps the global ids (the AMG set-up starts from them)
    const bool u_cols = which == FNP_MAT_A00 || which == FNP_MAT_P00 || which == FNP_MAT_A10;
    std::vector<int64_t> begins = u_cols ? c.u_begins : c.p_begins;
    for (auto &b : begins) b /= bs;               // scalar (node) ownership in Kronecker mode
    HostCsr loc = h;
    std::shared_ptr<HaloPlan> plan = build_halo(c, loc, begins, nullptr);
    const int64_t n_own_cols = begins[c.rank + 1] - begins[c.rank];
    csr_upload_pattern(c, d, loc, kNames[which], c.split_rows(loc.nrows) ? n_own_cols : -1, bs, n_own_cols);
    d.halo = (plan && bs > 1) ? expand_plan(c, *plan, bs) : plan;
    d.ncols_own = (int32_t)n_own_cols;
    d.nghost = plan ? plan->nghost : 0;
    c.local_cols[which] = std::move(loc.col);
  }
  FNP_CUDA(cudaStreamSynchronize(c.stream));        // src / rowptr staging goes out of scope
  c.have_pattern[which] = true;
  c.have_values[which] = false;
  c.host_vals_valid[which] = false;
  c.pattern_gen[which]++;
  c.is_setup = false;                               // hierarchies / work space / graph follow the pattern
  c.drop_graph();
}

void ingest_set_pattern(Ctx &c, int which, const int32_t *rowptr, const int32_t *colidx) {
  FNP_REQUIRE(c.have_layout, FNP_ERR_STATE, "fnp_set_layout must precede fnp_set_pattern");
  FNP_REQUIRE(which >= 0 && which < FNP_MAT_COUNT, FNP_ERR_ARG, "bad operator id");
  FNP_REQUIRE(rowptr != nullptr, FNP_ERR_ARG, "null pattern");
  c.user_rowptr[which].clear();
  c.user_col[which].clear();
  c.prune_mask[which].clear();
  c.pattern_pending[which] = false;
  if (c.prune && (which == FNP_MAT_A00 || which == FNP_MAT_P00)) {
    // DOLFIN stores the velocity block with the dense per-cell coupling of all components,
    // explicit zeros included: the pattern is finalised at the first
    // fnp_set_values, when the stored zeros are known and can be dropped
    int64_t nrows, ncols;
    op_shape(c, which, nrows, ncols);
    FNP_REQUIRE(rowptr[0] == 0, FNP_ERR_ARG, "rowptr[0] must be 0");
    for (int64_t i = 0; i < nrows; ++i) FNP_REQUIRE(rowptr[i + 1] >= rowptr[i], FNP_ERR_ARG, "rowptr not monotone");
    FNP_REQUIRE(colidx || rowptr[nrows] == 0, FNP_ERR_ARG, "null pattern");
    c.user_rowptr[which].assign(rowptr, rowptr + nrows + 1);
    c.user_col[which].assign(colidx, colidx + rowptr[nrows]);
    c.user_nnz[which] = rowptr[nrows];
    c.pattern_pending[which] = true;
    c.have_pattern[which] = true;
    c.have_values[which] = false;
    c.dmat[which].nnz = rowptr[nrows];      // "has entries" until the pattern is finalised
  } else {
    finalize_pattern(c, which, rowptr, colidx, nullptr);
    int64_t nrows, ncols;
    op_shape(c, which, nrows, ncols);
    c.user_nnz[which] = rowptr[nrows];
  }
}

// ---------------------------------------------------------------------------------------------
// device value path
// ---------------------------------------------------------------------------------------------
constexpr int ING_BAD_KRON = 1, ING_DROPPED_NONZERO = 2;

__device__ __forceinline__ bool kron_differs(double v, double w) { return fabs(w - v) > 1e-13 * (fabs(v) + fabs(w)); }

// SELL storage: one thread per (slice, lane) walks its row.  rowptr_exp: row pointers of the EXPANDED
// pattern (BS rows per stored row); map: expanded entry -> caller entry (null: identity).
template <int BS>
__global__ void __launch_bounds__(256)
sell_values_kernel(int nslices, const int32_t *__restrict__ sl_ptr, const int32_t *__restrict__ sl_col,
                   double *__restrict__ sl_val, const int32_t *__restrict__ perm, const int32_t *__restrict__ rowptr_exp,
                   const int32_t *__restrict__ map, const double *__restrict__ user, double *__restrict__ dinv,
                   int *__restrict__ flag) {
  const int slice = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (slice >= nslices) return;
  const int base = __ldg(sl_ptr + slice) & ~31;
  const int len = ((__ldg(sl_ptr + slice + 1) & ~31) - base) >> 5;
  const int row = __ldg(perm + slice * 32 + lane);
  int beg[BS], rl = 0;
  if (row >= 0) {
#pragma unroll
    for (int b = 0; b < BS; ++b) beg[b] = __ldg(rowptr_exp + BS * row + b);
    rl = __ldg(rowptr_exp + BS * row + 1) - beg[0];
  }
  double diag = 0.0;
  bool bad = false;
  for (int k = 0; k < len; ++k) {
    double v = 0.0;
    if (k < rl) {
      const int e = beg[0] + k;
      v = user[map ? map[e] : e];
#pragma unroll
      for (int b = 1; b < BS; ++b) {
        const int e2 = beg[b] + k;
        bad = bad || kron_differs(v, user[map ? map[e2] : e2]);
      }
      if (dinv && __ldg(sl_col + base + k * 32 + lane) == row) diag = v;
    }
    sl_val[base + k * 32 + lane] = v;
  }
  if (row >= 0 && dinv) {
    const double r = diag != 0.0 ? 1.0 / diag : 0.0;
#pragma unroll
    for (int b = 0; b < BS; ++b) dinv[(int64_t)BS * row + b] = r;
  }
  if (bad) atomicOr(flag, ING_BAD_KRON);
}

// CSR storage: a sub-warp of 8 lanes per row
template <int BS>
__global__ void __launch_bounds__(256)
csr_values_kernel(int nrows, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ col, double *__restrict__ val,
                  const int32_t *__restrict__ rowptr_exp, const int32_t *__restrict__ map,
                  const double *__restrict__ user, double *__restrict__ dinv, int *__restrict__ flag) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = tid >> 3, lane = tid & 7;
  if (row >= nrows) return;
  const int b0 = rowptr[row], len = rowptr[row + 1] - b0;
  int beg[BS];
#pragma unroll
  for (int b = 0; b < BS; ++b) beg[b] = BS == 1 ? b0 : rowptr_exp[BS * row + b];
  bool bad = false;
  for (int k = lane; k < len; k += 8) {
    const int e = beg[0] + k;
    const double v = user[map ? map[e] : e];
#pragma unroll
    for (int b = 1; b < BS; ++b) {
      const int e2 = beg[b] + k;
      bad = bad || kron_differs(v, user[map ? map[e2] : e2]);
    }
    val[b0 + k] = v;
    if (dinv && col[b0 + k] == row) {
      const double r = v != 0.0 ? 1.0 / v : 0.0;
#pragma unroll
      for (int b = 0; b < BS; ++b) dinv[(int64_t)BS * row + b] = r;
    }
  }
  if (bad) atomicOr(flag, ING_BAD_KRON);
}

__global__ void dropped_check_kernel(int64_t n, const int32_t *__restrict__ dropped, const double *__restrict__ user,
                                     int *__restrict__ flag) {
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    bad = bad || user[dropped[i]] != 0.0;
  if (bad) atomicOr(flag, ING_DROPPED_NONZERO);
}

template <int BS>
static void launch_values(Ctx &c, DevCsr &A, const int32_t *rowptr_exp, const int32_t *map, const double *user, bool want_dinv) {
  double *dinv = nullptr;
  if (want_dinv) {
    A.dinv.ensure((size_t)A.nrows * BS);
    if (A.nrows > 0)      // (a rank may own no row of an operator)
      FNP_CUDA(cudaMemsetAsync(A.dinv.p, 0, (size_t)A.nrows * BS * sizeof(double), c.stream));   // rows without a stored diagonal
    dinv = A.dinv.p;
  }
  if (A.sell) {
    auto part = [&](int nsl, const int32_t *ptr, const int32_t *perm, int64_t off) {
      if (nsl <= 0) return;
      const int grid = (int)(((int64_t)nsl * 32 + 255) / 256);
      sell_values_kernel<BS><<<grid, 256, 0, c.stream>>>(nsl, ptr, A.sl_col.p + off, A.sl_val.p + off, perm, rowptr_exp, map,
                                                         user, dinv, c.d_flag.p);
      c.launches++;
    };
    part(A.nslices, A.sl_ptr.p, A.sl_perm.p, 0);
    part(A.nslices_b, A.sl_ptr_b.p, A.sl_perm_b.p, A.sell_entries_a);
  } else if (A.nrows > 0) {
    const int grid = (int)(((int64_t)A.nrows * 8 + 255) / 256);
    csr_values_kernel<BS><<<grid, 256, 0, c.stream>>>(A.nrows, A.rowptr.p, A.col.p, A.val.p, rowptr_exp, map, user, dinv,
                                                      c.d_flag.p);
    c.launches++;
  }
  FNP_CUDA(cudaPeekAtLastError());
  A.has_dinv = A.has_dinv || want_dinv;
}

static bool is_device_pointer(const void *p) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// caller's value array on the device: in place when it is device memory, staged otherwise
static const double *stage_user_values(Ctx &c, const double *values, int64_t n) {
  if (n == 0) return values;
  if (is_device_pointer(values)) return values;
  c.stage_vals.ensure((size_t)n);
  FNP_CUDA(cudaMemcpyAsync(c.stage_vals.p, values, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c.stream));
  return c.stage_vals.p;
}

static int read_flag(Ctx &c) {
  int f = 0;
  FNP_CUDA(cudaMemcpyAsync(&f, c.d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  FNP_CUDA(cudaStreamSynchronize(c.stream));
  return f;
}

void ingest_set_values(Ctx &c, int which, const double *values) {
  FNP_REQUIRE(which >= 0 && which < FNP_MAT_COUNT, FNP_ERR_ARG, "bad operator id");
  FNP_REQUIRE(c.have_pattern[which], FNP_ERR_STATE, "fnp_set_values before fnp_set_pattern");
  const int64_t nnz_user = c.user_nnz[which];
  FNP_REQUIRE(values != nullptr || nnz_user == 0, FNP_ERR_ARG, "null values");
  if (!c.d_flag.p) c.d_flag.alloc(1);
  FNP_CUDA(cudaMemsetAsync(c.d_flag.p, 0, sizeof(int), c.stream));
  const double *user = stage_user_values(c, values, nnz_user);
  const bool pruning = !c.user_rowptr[which].empty();
  if (pruning) {
    // (re)build the pattern when it is still pending, or when an entry dropped earlier as a stored
    // zero carries a value now (e.g. the convection term after a zero initial guess); the decision
    // is collective because the pattern set-up is
    double rebuild = c.pattern_pending[which] ? 1.0 : 0.0;
    if (!c.pattern_pending[which] && c.n_dropped[which] > 0) {
      const int64_t n = c.n_dropped[which];
      dropped_check_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8), 256, 0, c.stream>>>(
          n, c.d_dropped[which].p, user, c.d_flag.p);
      c.launches++;
      FNP_CUDA(cudaPeekAtLastError());
      if (read_flag(c) & ING_DROPPED_NONZERO) rebuild = 1.0;
      FNP_CUDA(cudaMemsetAsync(c.d_flag.p, 0, sizeof(int), c.stream));
    }
    rebuild = comm_allreduce(c, rebuild, true);
    if (rebuild > 0.5) {
      // host copy of the values decides what is kept (first upload, or a rare pattern growth)
      std::vector<double> hv;
      const double *hvals = values;
      if (is_device_pointer(values)) {
        hv.resize((size_t)nnz_user);
        FNP_CUDA(cudaMemcpyAsync(hv.data(), values, (size_t)nnz_user * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        FNP_CUDA(cudaStreamSynchronize(c.stream));
        hvals = hv.data();
      }
      const std::vector<int32_t> &rp = c.user_rowptr[which], &ci = c.user_col[which];
      const int64_t nrows = (int64_t)rp.size() - 1;
      const int64_t row0 = c.u_begin;                       // A00 / P00: square in the u numbering
      std::vector<char> &keepmask = c.prune_mask[which];
      if (keepmask.size() != (size_t)nnz_user) keepmask.assign((size_t)nnz_user, 0);
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < nrows; ++i)
        for (int32_t k = rp[i]; k < rp[i + 1]; ++k)
          if (hvals[k] != 0.0 || ci[k] == row0 + i) keepmask[k] = 1;      // a kept entry stays kept
      finalize_pattern(c, which, rp.data(), ci.data(), keepmask.data());
      c.pattern_pending[which] = false;
    }
  }
  DevCsr &A = c.dmat[which];
  const bool want_dinv = which == FNP_MAT_MP || which == FNP_MAT_AP || which == FNP_MAT_A00 || which == FNP_MAT_P00;
  const int bs = c.kron_bs[which];
  const int32_t *rowptr_exp = bs > 1 ? c.d_exp_rowptr[which].p : A.rowptr.p;
  const int32_t *map = c.d_map[which].p;
  switch (bs) {
    case 1: launch_values<1>(c, A, rowptr_exp, map, user, want_dinv); break;
    case 2: launch_values<2>(c, A, rowptr_exp, map, user, want_dinv); break;
    case 3: launch_values<3>(c, A, rowptr_exp, map, user, want_dinv); break;
    default: throw Error(FNP_ERR_ARG, "unsupported block size");
  }
  const int f = read_flag(c);       // also: the caller's array may be reused after this call returns
  FNP_REQUIRE(!(f & ING_BAD_KRON), FNP_ERR_STATE,
              "the velocity block has the pattern of S (x) I but its values differ between "
              "components (Newton coupling / component-wise BCs?): set option fnp_kronecker 0 "
              "before fnp_set_pattern");
  c.have_values[which] = true;
  c.host_vals_valid[which] = false;
  c.dirty[which] = true;
}

// hmat[which].val <- device values (stored order undone through the host position map)
void ensure_host_values(Ctx &c, int which) {
  if (c.host_vals_valid[which]) return;
  FNP_REQUIRE(c.have_values[which], FNP_ERR_STATE, "operator has no values");
  HostCsr &h = c.hmat[which];
  DevCsr &A = c.dmat[which];
  const int64_t nnz = h.nnz();
  h.val.resize((size_t)nnz);
  if (A.sell) {
    std::vector<double> buf((size_t)A.sell_entries);
    FNP_CUDA(cudaMemcpyAsync(buf.data(), A.sl_val.p, buf.size() * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    FNP_CUDA(cudaStreamSynchronize(c.stream));
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < nnz; ++k) h.val[(size_t)k] = buf[(size_t)A.sell_pos[(size_t)k]];
  } else {
    if (nnz) FNP_CUDA(cudaMemcpyAsync(h.val.data(), A.val.p, (size_t)nnz * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    FNP_CUDA(cudaStreamSynchronize(c.stream));
  }
  c.host_vals_valid[which] = true;
}

}  // namespace fnp
